#include "common.cuh"
#include <stdarg.h>

namespace fcl {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return FCL_ECUDA;
  }
  return FCL_OK;
}
}  // namespace fcl

extern "C" int fcl_abi_version(void) { return FCL_ABI_VERSION; }
extern "C" int fcl_operand_format(void) {
#ifdef FCL_OPERANDS_BF16
  return 1;
#else
  return 0;
#endif
}
extern "C" const char* fcl_last_error(void) { return fcl::g_err; }
extern "C" int fcl_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    fcl::set_error("fcl_sm_count: %s", cudaGetErrorString(cudaGetLastError()));
    return FCL_ECUDA;
  }
  return n;
}

// ABI self-check for foreign-language bindings: sizeof of every parameter struct.
extern "C" int fcl_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(FclLenRegParams);
    case 1: return (int)sizeof(FclFrameMapParams);
    case 2: return (int)sizeof(FclConvGemmParams);
    case 3: return (int)sizeof(FclLayerNormParams);
    case 4: return (int)sizeof(FclEmbedAddParams);
    case 5: return (int)sizeof(FclBiLstmParams);
    case 6: return (int)sizeof(FclDecoderParams);
    case 7: return (int)sizeof(FclConvGemmBf16Params);
    case 8: return (int)sizeof(FclDecoderBf16Params);
    case 9: return (int)sizeof(FclPackRowsParams);
    case 10: return (int)sizeof(FclBiLstmBf16Params);
    case 11: return (int)sizeof(FclConvTilesParams);
    case 12: return (int)sizeof(FclDecoderScheduleParams);
    case 13: return (int)sizeof(FclConvStackTilesParams);
    case 14: return (int)sizeof(FclConvStackParams);
    case 15: return (int)sizeof(FclPadRowsParams);
    case 16: return (int)sizeof(FclRowsToImageParams);
    case 17: return (int)sizeof(FclConvImgParams);
    case 18: return (int)sizeof(FclPrenet0TfParams);
    default: return -1;
  }
}
