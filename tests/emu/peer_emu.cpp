// TEST INFRASTRUCTURE — host execution of csrc/peer_push.cu's kernels (see cuda_warp_shim.h).  `world` emulated GPUs are
// plain host buffers; the "peer pointer arrays" are arrays of host pointers.  The harness replays what
// distributed.PeerGather does: per step and rank one push launch (grid = world CTAs of 512 threads) into slice `rank` of
// every peer's gathered buffer, then per rank one wait launch, and reports what the waits saw.
#include "cuda_warp_shim.h"

static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __nanosleep(unsigned) { sched_yield(); }
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
namespace mg {
static inline void peer_st_release_sys(uint32_t* flag, uint32_t v) { __atomic_store_n(flag, v, __ATOMIC_RELEASE); }
static inline uint32_t peer_ld_acquire_sys(const uint32_t* flag) { return __atomic_load_n(flag, __ATOMIC_ACQUIRE); }
}  // namespace mg

#include "../../mingraph_unet_b200/csrc/peer_push.cu"

using namespace mg;

// one step of rank `rank`: push `src` (nbytes) into [off, off + nbytes) of every peer buffer, flag index `flag_index`
extern "C" int emu_peer_push(const void* src, long long nbytes, void** peer_bufs, int world, long long off, uint32_t** peer_signals,
                             long long flag_index, uint32_t* seq) {
  emu_faults = 0;
  if (nbytes % 16 || off % 16 || (uintptr_t)src % 16) return -2;
  for (int p = 0; p < world; ++p)
    emu_run_block(p, world, kPushThreads, [&]() {
      peer_push_kernel(reinterpret_cast<const uint4*>(src), nbytes / 16, peer_bufs, off, peer_signals, flag_index, seq);
    });
  return emu_faults ? -1 : 0;
}

extern "C" int emu_peer_wait(const uint32_t* my_signals, long long first_flag, int world, uint32_t* wseq, int* status,
                             unsigned long long max_spins) {
  emu_faults = 0;
  emu_run_block(0, 1, 64, [&]() { peer_wait_kernel(my_signals, first_flag, world, wseq, status, max_spins); });
  return emu_faults ? -1 : 0;
}
