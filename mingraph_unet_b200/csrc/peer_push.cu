// Exchange of the small per-image outputs over NVLink / NVSwitch peer memory, without a collective library call.
//
// Every rank owns a "gathered" buffer in symmetric memory (same allocation on every GPU of the node, each mapped into
// every peer's address space; Python obtains the mappings from torch.distributed._symmetric_memory).  After the block
// kernel has written a step's packed payload (l_partition | region_features | hard_labels, 74 KB at cfg 2), ONE launch
// of peer_push_kernel stores that payload into slice `rank` of EVERY peer's gathered buffer (CTA p -> peer p, 16-byte
// stores that travel as NVLink writes), then publishes a per-(slot, source-rank) sequence number in the peer's signal
// pad with a system-scope release.  Nothing in it waits for another GPU, so — unlike an NCCL kernel — it cannot hold SMs
// while a slower rank catches up, and it is an ordinary kernel node of the step's CUDA graph.
// A consumer of the gathered data enqueues peer_wait_kernel, which spins (bounded) until every source rank's sequence
// number has reached the consumer's own count for that slot.
//
// STATUS: written at the end of round 1 after the GPU budget was spent — compiled, NOT yet run on hardware; opt-in only
// (bench.py --exchange p2p).  See DESIGN.md §5 / §7.
#include "common.cuh"

namespace mg {

constexpr int kPushThreads = 512;

#ifndef MG_HOST_EMULATION   // tests/emu provides host versions (tests/test_peer_emulation.py)
__device__ __forceinline__ void peer_st_release_sys(uint32_t* flag, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t peer_ld_acquire_sys(const uint32_t* flag) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
  return v;
}
#endif

__global__ void __launch_bounds__(kPushThreads) peer_push_kernel(const uint4* __restrict__ src, int64_t nvec,
                                                                 void* const* __restrict__ peer_bufs, int64_t dst_off_bytes,
                                                                 uint32_t* const* __restrict__ peer_signals, int64_t flag_index,
                                                                 uint32_t* __restrict__ seq) {
  const int p = blockIdx.x;                                    // destination rank (own rank included)
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(peer_bufs[p]) + dst_off_bytes);
  for (int64_t i = threadIdx.x; i < nvec; i += kPushThreads) dst[i] = __ldg(src + i);
  __threadfence_system();                                      // each thread's stores are ordered before the flag below
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t v = seq[p] + 1u;                            // CTA p owns seq[p]: launches of one slot are serialised
    seq[p] = v;
    peer_st_release_sys(peer_signals[p] + flag_index, v);
  }
}

// thread r waits for source rank r.  status[0] is set to 1 if a flag did not arrive within the spin bound.
__global__ void peer_wait_kernel(const uint32_t* __restrict__ my_signals, int64_t first_flag, int world, uint32_t* __restrict__ wseq,
                                 int32_t* __restrict__ status, unsigned long long max_spins) {
  const int r = threadIdx.x;
  if (r >= world) return;
  const uint32_t need = wseq[r] + 1u;
  wseq[r] = need;
  const uint32_t* flag = my_signals + first_flag + r;
  for (unsigned long long spin = 0; spin < max_spins; ++spin) {
    const uint32_t v = peer_ld_acquire_sys(flag);
    if ((int32_t)(v - need) >= 0) return;                      // wrap-safe v >= need
    __nanosleep(64);
  }
  if (status) atomicExch(status, 1);
}

}  // namespace mg

#ifndef MG_HOST_EMULATION
using namespace mg;

extern "C" {

int mg_peer_push(const void* src, int64_t nbytes, const void* const* peer_bufs_dev, int world, int64_t dst_offset_bytes,
                 const void* const* peer_signals_dev, int64_t flag_index, uint32_t* seq, mg_stream_t stream) {
  MG_REQUIRE(src && peer_bufs_dev && peer_signals_dev && seq && world > 0 && world <= 64 && nbytes > 0, MG_ERR_INVALID,
             "mg_peer_push: bad arguments");
  MG_REQUIRE(nbytes % 16 == 0 && dst_offset_bytes % 16 == 0 && (uintptr_t)src % 16 == 0, MG_ERR_INVALID,
             "mg_peer_push: payload size, offset and source must be multiples of 16 bytes");
  peer_push_kernel<<<world, kPushThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(src), nbytes / 16, const_cast<void* const*>(peer_bufs_dev), dst_offset_bytes,
      reinterpret_cast<uint32_t* const*>(const_cast<void* const*>(peer_signals_dev)), flag_index, seq);
  return check_launch("peer_push_kernel");
}

int mg_peer_wait(const uint32_t* my_signals, int64_t first_flag, int world, uint32_t* wseq, int32_t* status, mg_stream_t stream) {
  MG_REQUIRE(my_signals && wseq && world > 0 && world <= 64, MG_ERR_INVALID, "mg_peer_wait: bad arguments");
  // 2 M polls of (64 ns sleep + one system-scope load, ~1 us) = a few seconds: a missing peer shows up in status[0],
  // not as a hung GPU
  peer_wait_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(my_signals, first_flag, world, wseq, status, 2ull * 1000 * 1000);
  return check_launch("peer_wait_kernel");
}

}  // extern "C"
#endif  // MG_HOST_EMULATION
