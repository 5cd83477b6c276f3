// Multi-GPU exchange of the small per-image outputs over NVLink / NVSwitch peer memory — the part that is not inside
// block_forward_kernel (block_fused.cu stores the payload into every rank's gathered buffer and publishes the step's
// sequence number; see PeerOut there):
//   * the exchange buffers: plain cudaMalloc allocations exported / opened with CUDA IPC handles, so every rank (one
//     process per GPU) holds a device pointer to every other rank's gathered buffer and flag array.  These are the only
//     entry points of the library that allocate; they are called once at set-up, never on the data path;
//   * peer_wait_kernel: the consumer side.  One thread per source rank polls that rank's flag (system-scope acquire)
//     until it has reached the step count of this rank's own slot, bounded, then the stream continues.
// Protocol (per pipeline slot): step number s of the slot goes to parity half s & 1 of the peers' slices, flag = s + 1.
// A rank rewrites a parity half two of the slot's steps later, after its own wait for the step in between — which it can
// only pass once every peer has PUBLISHED that step, i.e. (stream order on the peer) after the peer's consumers of the
// older step were enqueued.  No acknowledgements are needed.
#include <string.h>

#include "common.cuh"

namespace mg {

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// thread r waits for source rank r's flag to reach *seq (this rank's own completed-step count of the slot: the block
// kernel of the same step, earlier on this stream, has already advanced it).  status[0] = 1 when the bound expires.
__global__ void peer_wait_kernel(const uint32_t* __restrict__ my_flags, long long first_flag, int world,
                                 const uint32_t* __restrict__ seq, int32_t* __restrict__ status, unsigned long long max_spins) {
  const int r = threadIdx.x;
  if (r >= world) return;
  const uint32_t need = *reinterpret_cast<volatile const uint32_t*>(seq);
  const uint32_t* flag = my_flags + first_flag + r;
  for (unsigned long long spin = 0; spin < max_spins; ++spin) {
    if ((int32_t)(ld_acquire_sys_u32(flag) - need) >= 0) return;       // wrap-safe >=
    __nanosleep(100);
  }
  if (status) atomicExch(status, 1);
}

}  // namespace mg

using namespace mg;

extern "C" {

int mg_peer_wait(const uint32_t* my_flags, int64_t first_flag, int world, const uint32_t* seq, int32_t* status,
                 mg_stream_t stream) {
  MG_REQUIRE(my_flags && seq && world > 0 && world <= 64 && first_flag >= 0, MG_ERR_INVALID, "mg_peer_wait: bad arguments");
  // ~4 M polls of (100 ns sleep + one system-scope load) = a few seconds: a missing peer shows up in status[0], not as a
  // hung GPU
  peer_wait_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(my_flags, first_flag, world, seq, status, 4ull * 1000 * 1000);
  return check_launch("peer_wait_kernel");
}

int mg_peer_mem_alloc(int64_t nbytes, void** ptr_host, unsigned char* handle64_host) {
  MG_REQUIRE(ptr_host && handle64_host && nbytes > 0, MG_ERR_INVALID, "mg_peer_mem_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)nbytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)nbytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("mg_peer_mem_alloc: %s", cudaGetErrorString(e));
    cudaGetLastError();
    if (p) cudaFree(p);
    return MG_ERR_CUDA;
  }
  memcpy(handle64_host, &h, 64);
  *ptr_host = p;
  return MG_OK;
}

int mg_peer_mem_open(const unsigned char* handle64_host, void** ptr_host) {
  MG_REQUIRE(ptr_host && handle64_host, MG_ERR_INVALID, "mg_peer_mem_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("mg_peer_mem_open: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return MG_ERR_CUDA;
  }
  *ptr_host = p;
  return MG_OK;
}

int mg_peer_mem_close(void* ptr) {
  MG_REQUIRE(ptr, MG_ERR_INVALID, "mg_peer_mem_close: null pointer");
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) {
    set_error("mg_peer_mem_close: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return MG_ERR_CUDA;
  }
  return MG_OK;
}

int mg_peer_mem_free(void* ptr) {
  MG_REQUIRE(ptr, MG_ERR_INVALID, "mg_peer_mem_free: null pointer");
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) {
    set_error("mg_peer_mem_free: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return MG_ERR_CUDA;
  }
  return MG_OK;
}

}  // extern "C"
