// Shared helpers of the FCL-taco2 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/fcl_taco2.h"

namespace fcl {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define FCL_REQUIRE(cond, msg)                                                   \
  do {                                                                           \
    if (!(cond)) { ::fcl::set_error("%s: %s", __func__, msg); return FCL_EINVAL; } \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: remember what was set per (kernel, device)
// so that a second GPU in the same process (or a model on cuda:1 while cuda:0 is current elsewhere) gets it too.
template <typename Kernel>
static inline int ensure_dyn_smem(Kernel kernel, size_t bytes, const char* what) {
  static int have[64] = {0};                      // one table per kernel instantiation; benign race: worst case it is set twice
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess && (dev < 0 || dev >= 64 || have[dev] < (int)bytes)) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) have[dev] = (int)bytes;
  }
  if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return FCL_ECUDA; }
  return FCL_OK;
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- Philox4x32-10 (Salmon et al. SC'11); CPU twin: oracle/philox.py ------------------------
struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

// keep decisions of prenet units 8*oct .. 8*oct+7 for (utt, phoneme, step, layer): unit 8*oct+j uses the
// 16-bit lane  (j & 1) ? word[j >> 1] >> 16 : word[j >> 1] & 0xffff  and is kept iff lane >= dropout_threshold16(p)
__device__ __forceinline__ Philox4 dropout_words(uint64_t seed, uint32_t utt, uint32_t phoneme,
                                                  uint32_t step, uint32_t layer, uint32_t oct) {
  return philox4x32_10(oct, (step & 0xFFFFFFu) | (layer << 24), phoneme, utt,
                       (uint32_t)(seed & 0xFFFFFFFFull), (uint32_t)(seed >> 32));
}

__host__ __device__ __forceinline__ uint32_t dropout_threshold16(float p) {
  float t = p * 65536.0f + 0.5f;
  if (t >= 65535.0f) return 65535u;
  if (t <= 0.0f) return 0u;
  return (uint32_t)t;
}

// same for a kernel chosen at run time among template instantiations: keyed by (function pointer, device)
static inline int ensure_dyn_smem_fn(const void* kernel, size_t bytes, const char* what) {
  struct Entry { const void* fn; int dev; int bytes; };
  static Entry table[64] = {};
  static int n = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) {
    for (int i = 0; i < n; ++i)
      if (table[i].fn == kernel && table[i].dev == dev && table[i].bytes >= (int)bytes) return FCL_OK;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && n < 64) { table[n].fn = kernel; table[n].dev = dev; table[n].bytes = (int)bytes; ++n; }
  }
  if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return FCL_ECUDA; }
  return FCL_OK;
}

}  // namespace fcl
