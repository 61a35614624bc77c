// Peer-memory plumbing of the final mel gather: the root GPU exports receive buffers through CUDA
// IPC, the other ranks map them and push their finished mels with copy-engine DMA over NVLink (cudaMemcpyAsync on a
// side stream) -- no SM is taken from the persistent decoder of the next pass, and nothing on the hot path waits for
// it. Completion travels as stream-ordered 4-byte flag copies; the waiting side is a one-warp polling kernel.
// The reference has no counterpart (single-process, one utterance at a time: tts.py:655-674).
#include "common.cuh"
#include <string.h>

namespace fcl {

// lane r waits until flags[r] == expect[r] (volatile system-scope loads: the writer is another GPU's copy engine)
__global__ void __launch_bounds__(32, 1)
wait_flags_kernel(const volatile int32_t* flags, int n, int32_t expect) {
  const int lane = threadIdx.x;
  for (int i = lane; i < n; i += 32) {
    while (flags[i] != expect) __nanosleep(200);
  }
  __syncwarp();
  __threadfence_system();
}

__global__ void __launch_bounds__(32, 1)
write_flags_kernel(volatile int32_t* flags, int n, int32_t value) {
  __threadfence_system();
  for (int i = threadIdx.x; i < n; i += 32) flags[i] = value;
  __threadfence_system();
}

}  // namespace fcl

extern "C" int fcl_peer_alloc(int64_t bytes, void** dptr) {
  using namespace fcl;
  FCL_REQUIRE(dptr && bytes > 0, "bad arguments");
  cudaError_t e = cudaMalloc(dptr, (size_t)bytes);       // a dedicated allocation: IPC handles cover whole allocations
  if (e == cudaSuccess) e = cudaMemset(*dptr, 0, (size_t)bytes);
  if (e != cudaSuccess) { set_error("fcl_peer_alloc: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
  return FCL_OK;
}

extern "C" int fcl_peer_free(void* dptr) {
  using namespace fcl;
  cudaError_t e = cudaFree(dptr);
  if (e != cudaSuccess) { set_error("fcl_peer_free: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
  return FCL_OK;
}

extern "C" int fcl_ipc_export(void* dptr, uint8_t* handle64) {
  using namespace fcl;
  FCL_REQUIRE(dptr && handle64, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, dptr);
  if (e != cudaSuccess) { set_error("fcl_ipc_export: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
  memcpy(handle64, &h, 64);
  return FCL_OK;
}

extern "C" int fcl_ipc_open(const uint8_t* handle64, void** dptr) {
  using namespace fcl;
  FCL_REQUIRE(dptr && handle64, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess);   // mapped into the CURRENT device's context
  if (e != cudaSuccess) { set_error("fcl_ipc_open: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
  return FCL_OK;
}

extern "C" int fcl_ipc_close(void* dptr) {
  using namespace fcl;
  cudaError_t e = cudaIpcCloseMemHandle(dptr);
  if (e != cudaSuccess) { set_error("fcl_ipc_close: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
  return FCL_OK;
}

// device -> (peer) device copy on `stream`: executed by a copy engine, over NVLink when dst is a mapped peer buffer
extern "C" int fcl_copy_async(void* dst, const void* src, int64_t bytes, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(dst && src && bytes >= 0, "bad arguments");
  if (bytes == 0) return FCL_OK;
  cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, as_stream(stream));
  if (e != cudaSuccess) { set_error("fcl_copy_async: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
  return FCL_OK;
}

extern "C" int fcl_wait_flags(const int32_t* flags, int32_t n, int32_t expect, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(flags && n > 0, "bad arguments");
  wait_flags_kernel<<<1, 32, 0, as_stream(stream)>>>(flags, n, expect);
  return check_launch("fcl_wait_flags");
}

extern "C" int fcl_write_flags(int32_t* flags, int32_t n, int32_t value, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(flags && n > 0, "bad arguments");
  write_flags_kernel<<<1, 32, 0, as_stream(stream)>>>(flags, n, value);
  return check_launch("fcl_write_flags");
}
