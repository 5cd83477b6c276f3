// TEST INFRASTRUCTURE — extends cuda_host_shim.h to kernels that use warp collectives, block barriers, mbarriers and
// bulk (TMA) copies: every CUDA thread of a CTA is a host thread; __syncwarp / __shfl* / ldmatrix / mma are lock-step
// exchanges over a per-warp barrier, __syncthreads a per-block barrier; an mbarrier is a small state machine
// (arrivals + transaction bytes -> phase flip) and a bulk copy is a bounds- and alignment-checked memcpy that completes
// on it.  Shared memory is one host array; 32-bit "shared addresses" are offsets into it.  CTAs run one after another.
// This checks a kernel's CONTROL FLOW and INDEXING (rings, cursors, strip order, output addresses); the fragment
// layouts of ldmatrix / mma are implemented from the PTX documentation, i.e. they share the kernel author's reading.
#pragma once
#define MG_EMU_WARP 1
#include "cuda_host_shim.h"

#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

struct EmuWarp {
  pthread_barrier_t bar;
  uint32_t xchg[32][4];
};
struct EmuBlock {
  pthread_barrier_t bar;
};
static thread_local EmuWarp* emu_warp = nullptr;
static thread_local EmuBlock* emu_block = nullptr;
static thread_local int emu_lane = 0;

static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu_warp->bar); }
static inline void __syncthreads() { pthread_barrier_wait(&emu_block->bar); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

template <typename T>
static inline T emu_exchange(T v, int src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  memcpy(&emu_warp->xchg[emu_lane][0], &v, 4);
  pthread_barrier_wait(&emu_warp->bar);
  T r;
  memcpy(&r, &emu_warp->xchg[src_lane][0], 4);
  pthread_barrier_wait(&emu_warp->bar);
  return r;
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  return emu_exchange(v, (emu_lane & ~(width - 1)) + (src & (width - 1)));
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
  const int partner = emu_lane ^ mask;
  return emu_exchange(v, (partner & ~(width - 1)) == (emu_lane & ~(width - 1)) ? partner : emu_lane);
}

// ---- shared memory, mbarriers, bulk copies ----------------------------------------------------------------------
constexpr int kEmuSmemBytes = 256 * 1024;
namespace mg {
alignas(1024) unsigned char pt_smem[kEmuSmemBytes];          // the kernels' `extern __shared__ ... pt_smem[]`
}
struct EmuMBar {
  int count = 0, pending = 0;
  long long tx = 0;
  std::atomic<int> phase{0};
  bool init = false;
};
static EmuMBar emu_mbar[kEmuSmemBytes / 8];
static std::mutex emu_mbar_mu;
static const char* emu_src_lo = nullptr;                      // registered global input range of bulk copies
static const char* emu_src_hi = nullptr;
static std::atomic<long long> emu_copied_bytes{0};

static inline void emu_fault(const char* what) {
  fprintf(stderr, "emu: %s\n", what);
  ++emu_faults;
}
static inline void emu_mbar_try_complete(EmuMBar& b) {
  if (b.pending == 0 && b.tx == 0) {
    b.pending = b.count;
    b.phase.store(b.phase.load() ^ 1);
  }
}

namespace mg {
static inline uint32_t pt_smem_u32(const void* p) {
  const long long off = (const unsigned char*)p - pt_smem;
  if (off < 0 || off >= kEmuSmemBytes) emu_fault("address outside shared memory");
  return (uint32_t)off;
}
static inline void pt_mbar_init(uint32_t bar, int count) {
  if (bar % 8) emu_fault("misaligned mbarrier");
  std::lock_guard<std::mutex> g(emu_mbar_mu);
  EmuMBar& b = emu_mbar[bar / 8];
  b.count = b.pending = count;
  b.tx = 0;
  b.phase.store(0);
  b.init = true;
}
static inline void pt_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  std::lock_guard<std::mutex> g(emu_mbar_mu);
  EmuMBar& b = emu_mbar[bar / 8];
  if (!b.init) emu_fault("expect_tx on an uninitialised mbarrier");
  if (b.pending <= 0) emu_fault("more arrivals than the mbarrier expects in this phase");
  b.tx += bytes;
  b.pending -= 1;
  emu_mbar_try_complete(b);
}
static inline void pt_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t) {
  if (dst % 16 || (uintptr_t)src % 16 || bytes % 16 || bytes == 0) emu_fault("bulk copy: size / addresses must be multiples of 16");
  if ((const char*)src < emu_src_lo || (const char*)src + bytes > emu_src_hi) {
    emu_fault("bulk copy reads outside the input tensor");
    return;
  }
  if ((long long)dst + bytes > kEmuSmemBytes) {
    emu_fault("bulk copy writes outside shared memory");
    return;
  }
  memcpy(pt_smem + dst, src, bytes);
  emu_copied_bytes += bytes;
  std::lock_guard<std::mutex> g(emu_mbar_mu);
  EmuMBar& b = emu_mbar[bar / 8];
  b.tx -= bytes;
  emu_mbar_try_complete(b);
}
static inline void pt_mbar_wait(uint32_t bar, uint32_t parity) {
  EmuMBar& b = emu_mbar[bar / 8];
  long long spins = 0;
  while ((uint32_t)b.phase.load() == parity) {               // try_wait.parity P succeeds once the phase of parity P is over
    sched_yield();
    if (++spins > 200000000LL) {
      emu_fault("mbarrier wait never satisfied (deadlock)");
      abort();
    }
  }
}
static inline void pt_fence_mbar_init() {}
static inline uint64_t pt_policy_evict_first() { return 0; }

// ldmatrix.sync.aligned.m8n8.x4.shared.b16: matrix m, row j is read from the address given by lane 8 m + j (16 bytes);
// lane l receives, from every matrix, the 32-bit pair at row l / 4, columns 2 (l % 4), 2 (l % 4) + 1
static inline void pt_ldmatrix_x4(uint32_t addr, uint32_t (&a)[4]) {
  if (addr % 16) emu_fault("ldmatrix row address not 16-byte aligned");
  emu_warp->xchg[emu_lane][0] = addr;
  pthread_barrier_wait(&emu_warp->bar);
  for (int m = 0; m < 4; ++m) {
    const uint32_t row = emu_warp->xchg[8 * m + emu_lane / 4][0];
    memcpy(&a[m], pt_smem + row + (emu_lane % 4) * 4, 4);
  }
  pthread_barrier_wait(&emu_warp->bar);
}
static inline float emu_bf16_pair_sum(uint32_t u) { return __uint_as_float(u << 16) + __uint_as_float(u & 0xffff0000u); }
// mma.m16n8k16 row.col f32 += bf16 x bf16 with B = ones: D[i][n] += sum_k A[i][k].  A fragment of lane (g = l / 4, t = l % 4):
// a0 = A[g][2t, 2t+1], a1 = A[g+8][2t, 2t+1], a2 = A[g][2t+8, 2t+9], a3 = A[g+8][2t+8, 2t+9];  D: d0, d1 = D[g][2t, 2t+1],
// d2, d3 = D[g+8][2t, 2t+1]
static inline void pt_mma_ones(float (&d)[4], const uint32_t (&a)[4]) {
  memcpy(emu_warp->xchg[emu_lane], a, 16);
  pthread_barrier_wait(&emu_warp->bar);
  const int g = emu_lane / 4;
  float lo = 0.f, hi = 0.f;
  for (int t = 0; t < 4; ++t) {
    const uint32_t* f = emu_warp->xchg[4 * g + t];
    lo += emu_bf16_pair_sum(f[0]) + emu_bf16_pair_sum(f[2]);
    hi += emu_bf16_pair_sum(f[1]) + emu_bf16_pair_sum(f[3]);
  }
  d[0] += lo; d[1] += lo; d[2] += hi; d[3] += hi;
  pthread_barrier_wait(&emu_warp->bar);
}
}  // namespace mg

// run one CTA of `threads` threads (multiple of 32): fn() is called by every thread with the index variables set
template <typename F>
static void emu_run_block(unsigned bx, unsigned grid_x, int threads, F fn) {
  const int nwarps = threads / 32;
  std::vector<EmuWarp> warps(nwarps);
  EmuBlock blk;
  pthread_barrier_init(&blk.bar, nullptr, threads);
  for (auto& w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
  for (auto& b : emu_mbar) b.init = false;
  std::vector<std::thread> ts;
  ts.reserve(threads);
  for (int t = 0; t < threads; ++t)
    ts.emplace_back([&, t]() {
      threadIdx = {(unsigned)t, 0, 0};
      blockIdx = {bx, 0, 0};
      blockDim = dim3(threads, 1, 1);
      gridDim = dim3(grid_x, 1, 1);
      emu_warp = &warps[t / 32];
      emu_block = &blk;
      emu_lane = t % 32;
      fn();
    });
  for (auto& th : ts) th.join();
  for (auto& w : warps) pthread_barrier_destroy(&w.bar);
  pthread_barrier_destroy(&blk.bar);
}
