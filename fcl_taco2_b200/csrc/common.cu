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

// L2 set-aside for persisting accesses on the current device (cudaLimitPersistingL2CacheSize, clamped to the device
// maximum). Kernels that mark a window persisting (the pair decoder's cell-state scratch) do so per launch and demote
// their lines before they exit, so other kernels see the whole L2; without a set-aside the windows have no effect.
namespace fcl { static size_t g_l2_persist[64] = {0}; size_t l2_persist_bytes() { int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0; return g_l2_persist[dev]; } }
extern "C" int fcl_l2_persist_limit(int64_t bytes, int64_t* set_bytes) {
  int dev = 0, maxb = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&maxb, cudaDevAttrMaxPersistingL2CacheSize, dev) != cudaSuccess) {
    fcl::set_error("fcl_l2_persist_limit: %s", cudaGetErrorString(cudaGetLastError()));
    return FCL_ECUDA;
  }
  size_t want = bytes < 0 ? 0 : (size_t)bytes;
  if (want > (size_t)maxb) want = (size_t)maxb;
  if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
    fcl::set_error("fcl_l2_persist_limit: %s", cudaGetErrorString(cudaGetLastError()));
    return FCL_ECUDA;
  }
  size_t got = 0;
  cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
  if (dev >= 0 && dev < 64) fcl::g_l2_persist[dev] = got;
  if (set_bytes) *set_bytes = (int64_t)got;
  return FCL_OK;
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
