// K1 patch mean pooling, region mean pooling (a11) and K7 nearest un-pool (a13).
// All three are HBM-streaming kernels: K1 reads the feature map once with 16-byte loads,
// K7 writes the dense (B,D,H,W) map once with 16-byte streaming stores.
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <type_traits>

#include "common.cuh"

namespace mg {

// ------------------------------------------------------------------------------------------
// K1: (B,C,Hf,Wf) -> (B, Hp*Wp, C), mean over ph x pw windows, zero padded right/bottom
// oracle: image_to_patches(x).mean((2,3))  (patch_graph_construction.py:27-47)
// ------------------------------------------------------------------------------------------
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ float sum(const float* p) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return (v.x + v.y) + (v.z + v.w);
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ float sum(const __nv_bfloat16* p) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    float s = 0.f;
    const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) s += __uint_as_float(w[i] << 16) + __uint_as_float(w[i] & 0xffff0000u);
    return s;
  }
};

// channels per block = warps per block, one channel strip per warp.  Measured at cfg 2 (B200, two shard branches):
// 32 / 16 / 8 / 4 / 2 channels per block -> 169.2 / 168.5 / 171.2 / 164.4 / 163.9 us per step: many short blocks ramp the
// HBM pipeline faster than few long ones.
static int pool_cc() {
  static const int v = getenv("MG_POOL_CC") ? atoi(getenv("MG_POOL_CC")) : 4;
  return v < 1 ? 1 : (v > 32 ? 32 : v);
}
constexpr int kPoolRows = 16; // rows whose 16-byte loads are issued together

// fast path: pw % VEC == 0, Wf % VEC == 0, lpp = pw/VEC a power of two <= 32.
// Block = (x chunk of 32 vectors, channel group) x patch row x image; warp w takes channels w, w+NW, ...
// of the group and, per channel, issues all ph row loads of its 512-byte strip before summing them.
template <typename TX, typename TO>
__global__ void __launch_bounds__(256) pool_patches_vec_kernel(const TX* __restrict__ x, int C, int Hf, int Wf, int ph,
                                                               int pw, int Hp, int Wp, int cg, int xchunks,
                                                               TO* __restrict__ out) {
  constexpr int VEC = Vec16<TX>::N;
  extern __shared__ float tile[];   // [32 / lpp][cg + 1]
  const int b = blockIdx.z, py = blockIdx.y;
  const int xc = blockIdx.x % xchunks, c0 = (blockIdx.x / xchunks) * cg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int lpp = pw / VEC;                               // lanes per patch
  const int pp = 32 / lpp;                                // patches per chunk
  const int y0 = py * ph, rows = min(Hf, y0 + ph) - y0;
  const float inv = 1.f / (float)(ph * pw);
  const int nvec = Wf / VEC;
  const int xv = xc * 32 + lane;
  const bool in_x = xv < nvec;
  const int ncc = min(cg, C - c0);
  for (int cc = warp; cc < ncc; cc += nw) {
    const TX* p = x + (((size_t)b * C + c0 + cc) * Hf + y0) * Wf + (size_t)xv * VEC;
    float acc = 0.f;
    for (int r0 = 0; r0 < rows; r0 += kPoolRows) {
      float part[kPoolRows];
#pragma unroll
      for (int r = 0; r < kPoolRows; ++r)
        part[r] = (in_x && r0 + r < rows) ? Vec16<TX>::sum(p + (size_t)(r0 + r) * Wf) : 0.f;
#pragma unroll
      for (int st = kPoolRows / 2; st > 0; st >>= 1)
#pragma unroll
        for (int r = 0; r < st; ++r) part[r] += part[r + st];
      acc += part[0];
    }
    for (int o = 1; o < lpp; o <<= 1) acc += __shfl_xor_sync(kFull, acc, o);
    if ((lane & (lpp - 1)) == 0) tile[(lane / lpp) * (cg + 1) + cc] = acc * inv;
  }
  __syncthreads();
  const int px0 = xc * pp;
  for (int idx = threadIdx.x; idx < pp * ncc; idx += blockDim.x) {
    const int pl = idx / ncc, cc = idx - pl * ncc;
    if (px0 + pl < Wp)
      out[((size_t)b * Hp * Wp + (size_t)py * Wp + px0 + pl) * C + c0 + cc] = from_f32<TO>(tile[pl * (cg + 1) + cc]);
  }
}

// ------------------------------------------------------------------------------------------
// TMA path (default when it applies): a (image, channel, patch-row) strip is ph FULL rows of the map = one contiguous
// run of ph*Wf elements, so the feature map is a stream of contiguous strips.  Persistent CTAs (one per SM) take strips
// round-robin; inside a CTA warp w owns strips w, w + kPtWarps, ... and runs its OWN ring of `stages` shared-memory
// buffers: lane 0 moves whole rows global -> shared with cp.async.bulk (chunks of <= stage_bytes) completing on the
// ring's mbarriers (expect_tx / complete_tx), always `stages` chunks ahead of the chunk being summed, so ~200 KB per SM
// are in flight with no global load instructions or address arithmetic in the summing code.  The warp reads a landed
// chunk with conflict-free 16-byte shared loads (kPtUnroll rows per batch), keeps fp32 column sums in registers, folds
// lanes into patch columns by shuffle and stores one (N,C) element per patch.  A ring is private to its warp, so the
// buffer hand-back needs no second barrier: after __syncwarp() lane 0 re-arms the buffer it has just drained.  Strips
// are ordered channel-fastest so that concurrently processed strips fill the same output sectors.
// ------------------------------------------------------------------------------------------
constexpr int kPtWarps = 8;
constexpr int kPtUnroll = 8;        // rows whose 16-byte shared loads are issued together
constexpr int kPtMaxPasses = 8;
constexpr int kPtSmemBytes = 208 * 1024;

__device__ __forceinline__ uint32_t pt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pt_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pt_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pt_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PT_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PT_WAIT_DONE;\n"
      "bra PT_WAIT_LOOP;\n"
      "PT_WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// read-once stream: L2 evict-first policy so the map does not displace the block's working set
__device__ __forceinline__ void pt_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

// sum of 16 bytes as fp32
template <typename T>
__device__ __forceinline__ float pt_sum16(const uint4& u);
template <>
__device__ __forceinline__ float pt_sum16<float>(const uint4& u) {
  return (__uint_as_float(u.x) + __uint_as_float(u.y)) + (__uint_as_float(u.z) + __uint_as_float(u.w));
}
template <>
__device__ __forceinline__ float pt_sum16<__nv_bfloat16>(const uint4& u) {
  const unsigned w[4] = {u.x, u.y, u.z, u.w};
  float lo = 0.f, hi = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lo += __uint_as_float(w[i] << 16);
    hi += __uint_as_float(w[i] & 0xffff0000u);
  }
  return lo + hi;
}

// Dynamic strip scheduling (MG_POOL_DYNAMIC=1; off by default: written after the round-1 GPU budget was spent, not yet
// measured).  Counter pairs live in a small per-device array; the kernel's last CTA re-arms the pair it used.  A launch
// recorded into a CUDA graph OWNS its pair for the life of the process (its replays are serialised by the graph's
// stream; pairs 32..63, static scheduling once they are used up), eager launches take pairs 0..31 round-robin — so an
// eager launch can never share a pair with a concurrently replaying graph.
__device__ int g_pool_counters[64][2];
static int* pool_counters(cudaStream_t st) {
  static const int enabled = getenv("MG_POOL_DYNAMIC") ? atoi(getenv("MG_POOL_DYNAMIC")) : 0;
  if (!enabled) return nullptr;
  constexpr int kMaxDev = 64;
  static int* base[kMaxDev] = {};                            // the symbol has one instance per device
  static std::atomic<unsigned> next_eager[kMaxDev], next_captured[kMaxDev];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
  if (!base[dev] && cudaGetSymbolAddress(reinterpret_cast<void**>(&base[dev]), g_pool_counters) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (cs == cudaStreamCaptureStatusActive) {
    const unsigned i = next_captured[dev].fetch_add(1);
    return i < 32 ? base[dev] + 2 * (32 + i) : nullptr;
  }
  return base[dev] + 2 * (next_eager[dev].fetch_add(1) % 32);
}

struct PoolTmaArgs {
  const void* x;
  void* out;
  int C, Hf, Wf, ph, pw, Hp, Wp;
  int rpc;          // rows per chunk (rpc * Wf * esz <= stage_bytes)
  int nstrips;      // B * Hp * C, strip s = (b * Hp + py) * C + c
  int stages;       // ring depth per warp
  int stage_bytes;  // multiple of 128
  int* counters;    // null: static round-robin strips.  else {next strip, finished CTAs}: dynamic strip scheduling
};

constexpr int kPtFifo = 16;         // per-warp FIFO of fetched strip ids (dynamic scheduling), > stages + 1

// walks the chunks of one warp's strips in order
template <typename TX>
struct PtCursor {
  int s, step, r0, rows, c;
  size_t orow;
  const TX* base;
  __device__ __forceinline__ void load(const PoolTmaArgs& A) {
    if (s >= A.nstrips) return;
    c = s % A.C;
    const int bp = s / A.C;
    const int py = bp % A.Hp, b = bp / A.Hp;
    const int y0 = py * A.ph;
    rows = min(A.Hf, y0 + A.ph) - y0;
    r0 = 0;
    orow = ((size_t)b * A.Hp + py) * A.Wp;
    base = reinterpret_cast<const TX*>(A.x) + (((size_t)b * A.C + c) * A.Hf + y0) * A.Wf;
  }
  __device__ __forceinline__ bool valid(const PoolTmaArgs& A) const { return s < A.nstrips; }
  // advance by one chunk; returns true when the strip is exhausted (the caller then sets s and calls load)
  __device__ __forceinline__ bool advance(const PoolTmaArgs& A) {
    r0 += A.rpc;
    return r0 >= rows;
  }
};

template <typename TX, typename TO, bool DYN>
__global__ void __launch_bounds__(kPtWarps * 32, 1) pool_patches_tma_kernel(const PoolTmaArgs A) {
  constexpr int VEC = Vec16<TX>::N;
  extern __shared__ __align__(128) unsigned char pt_smem[];
  const int q = A.stages;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ring = pt_smem + (size_t)warp * q * A.stage_bytes;
  const uint32_t ring0 = pt_smem_u32(ring);
  const uint32_t full0 = pt_smem_u32(pt_smem) + (uint32_t)(kPtWarps * q) * (uint32_t)A.stage_bytes + 8u * (uint32_t)(warp * q);
  if (lane == 0) {
    for (int i = 0; i < q; ++i) pt_mbar_init(full0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  const int row_bytes = A.Wf * (int)sizeof(TX);
  const int rv = row_bytes >> 4;                              // 16-byte vectors per row
  const int nvec = A.Wf / VEC;
  const int passes = ceil_div(nvec, 32);
  const int lpp = A.pw / VEC;
  const float inv = 1.f / (float)(A.ph * A.pw);
  TO* out = reinterpret_cast<TO*>(A.out);

  // Strip order.  Static: warp w of CTA b takes strips b + (w + 8 i) * grid.  Dynamic (A.counters): every warp draws
  // its next strip from one global counter, so CTAs that become resident late (SMs held by a neighbouring step's
  // cluster kernel) simply draw fewer strips; the issue cursor draws, the consume cursor follows through a small
  // per-warp FIFO in shared memory (an id >= nstrips terminates both).
  constexpr bool dynamic = DYN;                              // the static instantiation carries none of the dynamic code
  int* fifo = reinterpret_cast<int*>(pt_smem + (size_t)kPtWarps * q * (A.stage_bytes + 8)) + warp * kPtFifo;
  int tail = 0, head = 0;
  auto draw = [&]() -> int {                                    // issue side: next strip id
    int s = 0;
    if (lane == 0) {
      s = atomicAdd(A.counters, 1);
      fifo[tail & (kPtFifo - 1)] = s;
    }
    ++tail;
    __syncwarp();
    return __shfl_sync(kFull, s, 0);
  };
  auto follow = [&]() -> int {                                  // consume side: same sequence, later
    const int s = fifo[head & (kPtFifo - 1)];
    ++head;
    return s;
  };
  PtCursor<TX> ic, cc;                                        // issue cursor (q chunks ahead), consume cursor
  ic.step = cc.step = kPtWarps * gridDim.x;
  if constexpr (dynamic) {
    ic.s = draw();
    cc.s = follow();
  } else {
    ic.s = cc.s = blockIdx.x + warp * gridDim.x;
  }
  ic.load(A);
  cc.load(A);
  int issued = 0;
  auto issue = [&]() {
    if (!ic.valid(A)) return;
    if (lane == 0) {
      const int st = issued % q;
      const uint32_t bytes = (uint32_t)(min(A.rpc, ic.rows - ic.r0) * row_bytes);
      pt_mbar_expect_tx(full0 + 8 * st, bytes);
      pt_bulk_g2s(ring0 + (uint32_t)st * (uint32_t)A.stage_bytes, ic.base + (size_t)ic.r0 * A.Wf, bytes, full0 + 8 * st, policy);
    }
    ++issued;
    if (ic.advance(A)) {
      if constexpr (dynamic) ic.s = draw();
      else ic.s += ic.step;
      ic.load(A);
    }
  };
  for (int t = 0; t < q; ++t) issue();

  int consumed = 0;
  float acc[kPtMaxPasses];
  while (cc.valid(A)) {
    if (cc.r0 == 0) {
#pragma unroll
      for (int p = 0; p < kPtMaxPasses; ++p) acc[p] = 0.f;
    }
    const int st = consumed % q;
    pt_mbar_wait(full0 + 8 * st, (uint32_t)((consumed / q) & 1));
    const int nr = min(A.rpc, cc.rows - cc.r0);
    const uint4* sp = reinterpret_cast<const uint4*>(ring + (size_t)st * A.stage_bytes) + lane;
    // runtime loop over the 512-byte column passes (ONE copy of the summing code: fully unrolling it per pass made the
    // kernel instruction-cache bound); the pass's register accumulator is selected by predicated adds
    for (int p = 0; p < passes; ++p) {
      float part = 0.f;
      if (p * 32 + lane < nvec) {
        const uint4* a0 = sp + p * 32;
        int r = 0;
        for (; r + kPtUnroll <= nr; r += kPtUnroll) {
          uint4 u[kPtUnroll];
#pragma unroll
          for (int j = 0; j < kPtUnroll; ++j) u[j] = a0[(r + j) * rv];
          float t[kPtUnroll];
#pragma unroll
          for (int j = 0; j < kPtUnroll; ++j) t[j] = pt_sum16<TX>(u[j]);
#pragma unroll
          for (int h = kPtUnroll / 2; h > 0; h >>= 1)
#pragma unroll
            for (int j = 0; j < h; ++j) t[j] += t[j + h];
          part += t[0];
        }
        for (; r < nr; ++r) part += pt_sum16<TX>(a0[r * rv]);
      }
#pragma unroll
      for (int pp = 0; pp < kPtMaxPasses; ++pp)
        if (pp == p) acc[pp] += part;
    }
    __syncwarp();                                             // every lane has read the buffer: lane 0 may re-arm it
    issue();
    ++consumed;
    const bool last = cc.r0 + A.rpc >= cc.rows;
    if (last) {
#pragma unroll
      for (int p = 0; p < kPtMaxPasses; ++p) {
        if (p < passes) {
          float v = acc[p];
          for (int o = 1; o < lpp; o <<= 1) v += __shfl_xor_sync(kFull, v, o);
          const int px = (p * 32 + lane) / lpp;
          if ((lane & (lpp - 1)) == 0 && p * 32 + lane < nvec && px < A.Wp)
            out[(cc.orow + px) * A.C + cc.c] = from_f32<TO>(v * inv);
        }
      }
    }
    if (cc.advance(A)) {
      if constexpr (dynamic) cc.s = follow();
      else cc.s += cc.step;
      cc.load(A);
    }
  }
  if constexpr (dynamic) {                                    // the last CTA to finish re-arms the counters
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(A.counters + 1, 1) == (int)gridDim.x - 1) {
        A.counters[0] = 0;
        A.counters[1] = 0;
        __threadfence();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Tensor-core summation (opt-in, MG_POOL_MMA=1; bf16 maps, pw == 16, Wf % 256 == 0).  STATUS: written after the
// round-1 GPU budget was spent — compiles, NOT yet run on hardware, off by default.
// Why: the bulk-copy kernel above always finds its data landed (long-scoreboard stalls 2 % of samples) and spends its
// time ISSUING the sum — ~21 warp instructions per 512 bytes (shift / mask / add per bf16 pair) with two warps per
// scheduler (profiles/r1_pool_tma.md).  A patch row of 16 bf16 pixels is 32 contiguous bytes, so 512 contiguous bytes
// of an image row are a 16 x 16 row-major matrix A (row = patch, column = pixel); with B = ones,
// mma.m16n8k16 (bf16 x bf16 -> fp32) adds that image row's 16 patch-row sums into D, and accumulating D over the ph
// image rows of the strip gives the 16 patch sums: ONE ldmatrix.x4 + ONE mma per 512 bytes.  Products with 1.0 are
// exact and the accumulation is fp32, like the scalar code (different summation order).
// Same persistent per-warp TMA rings and static strip order as pool_patches_tma_kernel; only the consumer differs.
// ldmatrix lane addresses: matrix m = lane / 8 covers A rows (lane % 8) + 8 (m & 1); the two 16-byte halves of an A row
// may go to either k half (a sum does not care about k order), so rows 4..7 of every 8-row phase take the other half:
// the 8 addresses of a phase then fall into 8 different 16-byte bank groups (rows are only 32 bytes apart).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pt_ldmatrix_x4(uint32_t addr, uint32_t (&a)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void pt_mma_ones(float (&d)[4], const uint32_t (&a)[4]) {
  const uint32_t ones = 0x3F803F80u;                          // bf16 (1.0, 1.0)
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(ones), "r"(ones));
}

template <typename TO>
__global__ void __launch_bounds__(kPtWarps * 32, 1) pool_patches_mma_kernel(const PoolTmaArgs A) {
  using TX = __nv_bfloat16;
  extern __shared__ __align__(128) unsigned char pt_smem[];
  const int q = A.stages;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ring = pt_smem + (size_t)warp * q * A.stage_bytes;
  const uint32_t ring0 = pt_smem_u32(ring);
  const uint32_t full0 = pt_smem_u32(pt_smem) + (uint32_t)(kPtWarps * q) * (uint32_t)A.stage_bytes + 8u * (uint32_t)(warp * q);
  if (lane == 0) {
    for (int i = 0; i < q; ++i) pt_mbar_init(full0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  const int row_bytes = A.Wf * (int)sizeof(TX);
  const int tiles = A.Wf >> 8;                               // 16 patches x 16 pixels = 512 bytes per tile and image row
  const float inv = 1.f / (float)(A.ph * A.pw);
  TO* out = reinterpret_cast<TO*>(A.out);
  const uint32_t lane_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * 32 + ((((lane >> 4) & 1) ^ ((lane >> 2) & 1)) * 16));

  PtCursor<TX> ic, cc;                                        // issue cursor (q chunks ahead), consume cursor
  ic.step = cc.step = kPtWarps * gridDim.x;
  ic.s = cc.s = blockIdx.x + warp * gridDim.x;
  ic.load(A);
  cc.load(A);
  int issued = 0;
  auto issue = [&]() {
    if (!ic.valid(A)) return;
    if (lane == 0) {
      const int st = issued % q;
      const uint32_t bytes = (uint32_t)(min(A.rpc, ic.rows - ic.r0) * row_bytes);
      pt_mbar_expect_tx(full0 + 8 * st, bytes);
      pt_bulk_g2s(ring0 + (uint32_t)st * (uint32_t)A.stage_bytes, ic.base + (size_t)ic.r0 * A.Wf, bytes, full0 + 8 * st, policy);
    }
    ++issued;
    if (ic.advance(A)) {
      ic.s += ic.step;
      ic.load(A);
    }
  };
  for (int t = 0; t < q; ++t) issue();

  int consumed = 0;
  float d[kPtMaxPasses][4];
  while (cc.valid(A)) {
    if (cc.r0 == 0) {
#pragma unroll
      for (int p = 0; p < kPtMaxPasses; ++p) d[p][0] = d[p][1] = d[p][2] = d[p][3] = 0.f;
    }
    const int st = consumed % q;
    pt_mbar_wait(full0 + 8 * st, (uint32_t)((consumed / q) & 1));
    const int nr = min(A.rpc, cc.rows - cc.r0);
    const uint32_t buf = ring0 + (uint32_t)st * (uint32_t)A.stage_bytes + lane_off;
    for (int r = 0; r < nr; ++r) {
#pragma unroll
      for (int p = 0; p < kPtMaxPasses; ++p) {
        if (p < tiles) {
          uint32_t a[4];
          pt_ldmatrix_x4(buf + (uint32_t)(r * row_bytes + p * 512), a);
          pt_mma_ones(d[p], a);
        }
      }
    }
    __syncwarp();                                             // every lane has read the buffer: lane 0 may re-arm it
    issue();
    ++consumed;
    if (cc.r0 + A.rpc >= cc.rows) {                           // strip complete: D rows g and g + 8 of lanes with t == 0
      if ((lane & 3) == 0) {
        const int g = lane >> 2;
#pragma unroll
        for (int p = 0; p < kPtMaxPasses; ++p) {
          if (p < tiles) {
            const int px = p * 16 + g;
            if (px < A.Wp) out[(cc.orow + px) * A.C + cc.c] = from_f32<TO>(d[p][0] * inv);
            if (px + 8 < A.Wp) out[(cc.orow + px + 8) * A.C + cc.c] = from_f32<TO>(d[p][2] * inv);
          }
        }
      }
    }
    if (cc.advance(A)) {
      cc.s += cc.step;
      cc.load(A);
    }
  }
}

// generic path: any window; one thread per output element (px fastest so window reads share lines)
template <typename TX, typename TO>
__global__ void pool_patches_generic_kernel(const TX* __restrict__ x, int B, int C, int Hf, int Wf, int ph, int pw,
                                            int Hp, int Wp, TO* __restrict__ out) {
  const int64_t total = (int64_t)B * C * Hp * Wp;
  const float inv = 1.f / (float)(ph * pw);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int px = (int)(t % Wp);
    const int py = (int)((t / Wp) % Hp);
    const int c = (int)((t / ((int64_t)Wp * Hp)) % C);
    const int b = (int)(t / ((int64_t)Wp * Hp * C));
    const TX* plane = x + ((size_t)b * C + c) * Hf * Wf;
    const int y1 = min(Hf, (py + 1) * ph), x1 = min(Wf, (px + 1) * pw);
    float acc = 0.f;
    for (int y = py * ph; y < y1; ++y)
      for (int xx = px * pw; xx < x1; ++xx) acc += to_f32<TX>(plane[(size_t)y * Wf + xx]);
    out[((size_t)b * Hp * Wp + (size_t)py * Wp + px) * C + c] = from_f32<TO>(acc * inv);
  }
}

// ------------------------------------------------------------------------------------------
// a11: region mean pool.  Two deterministic stages (no atomics, fixed reduction order):
//   partial: block (chunk, image) reduces its chunk of nodes with warp-private shared accumulators;
//   finalize: thread per (image, k, d) sums the chunk partials in chunk order and divides by the count.
// ------------------------------------------------------------------------------------------
constexpr int kSegChunkNodes = 64;

__global__ void __launch_bounds__(256) segment_partial_kernel(const float* __restrict__ h, const int32_t* __restrict__ labels,
                                                              int N, int D, int K, float* __restrict__ part,
                                                              int32_t* __restrict__ part_cnt) {
  extern __shared__ float acc[];                       // [8][K][D] then int cnt[8][K]
  int* cnt = reinterpret_cast<int*>(acc + (size_t)8 * K * D);
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * K * D; i += blockDim.x) acc[i] = 0.f;
  for (int i = threadIdx.x; i < 8 * K; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  float* mine = acc + (size_t)warp * K * D;
  const int n0 = chunk * kSegChunkNodes, n1 = min(N, n0 + kSegChunkNodes);
  for (int n = n0 + warp; n < n1; n += 8) {
    const int k = labels ? __ldg(labels + (size_t)b * N + n) : n;
    if (k < 0 || k >= K) continue;
    const float* row = h + ((size_t)b * N + n) * D;
    for (int d = lane; d < D; d += 32) mine[k * D + d] += __ldg(row + d);
    if (lane == 0) cnt[warp * K + k] += 1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += acc[(size_t)w * K * D + i];
    part[((size_t)b * nchunks + chunk) * K * D + i] = s;
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    int c = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) c += cnt[w * K + k];
    part_cnt[((size_t)b * nchunks + chunk) * K + k] = c;
  }
}

// mean != 0: out = sum / count (0 for empty regions, train_end_to_end.py:372-373); else out = scale * sum
__global__ void segment_finalize_kernel(const float* __restrict__ part, const int32_t* __restrict__ part_cnt, int B, int nchunks,
                                        int K, int D, int mean, float scale, float* __restrict__ out,
                                        int32_t* __restrict__ counts) {
  const int total = B * K * D;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int d = t % D, k = (t / D) % K, b = t / (D * K);
    float s = 0.f;
    int c = 0;
    for (int ch = 0; ch < nchunks; ++ch) {
      s += part[((size_t)b * nchunks + ch) * K * D + k * D + d];
      c += part_cnt[((size_t)b * nchunks + ch) * K + k];
    }
    out[t] = mean ? (c > 0 ? s / (float)c : 0.f) : s * scale;
    if (counts && d == 0) counts[b * K + k] = c;
  }
}

// shared by mg_segment_mean and the un-pool backward (block_backward.cu).  work: segment_work_bytes().
int64_t segment_work_bytes(int B, int N, int D, int K) {
  const int nchunks = ceil_div(N, kSegChunkNodes);
  return (int64_t)B * nchunks * K * (D + 1) * 4 + 256;
}
int segment_reduce_launch(const float* h, const int32_t* labels, int B, int N, int D, int K, int mean, float scale, float* out,
                          int32_t* counts, void* work, cudaStream_t st) {
  const size_t smem = (size_t)8 * K * D * 4 + (size_t)8 * K * 4;
  MG_REQUIRE(smem <= 200 * 1024, MG_ERR_UNSUPPORTED, "segment reduce: K*D=%d too large for shared accumulators", K * D);
  const int nchunks = ceil_div(N, kSegChunkNodes);
  float* part = reinterpret_cast<float*>(work);
  int32_t* part_cnt = reinterpret_cast<int32_t*>(part + (size_t)B * nchunks * K * D);
  if (smem > 48 * 1024) cudaFuncSetAttribute(segment_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  dim3 grid(nchunks, B);
  MG_REQUIRE(B <= 65535, MG_ERR_INVALID, "segment reduce: batch too large");
  segment_partial_kernel<<<grid, 256, smem, st>>>(h, labels, N, D, K, part, part_cnt);
  int rc = check_launch("segment_partial_kernel");
  if (rc) return rc;
  segment_finalize_kernel<<<std::min(ceil_div(B * K * D, 256), num_sms() * 4), 256, 0, st>>>(part, part_cnt, B, nchunks, K, D, mean,
                                                                                       scale, out, counts);
  return check_launch("segment_finalize_kernel");
}

// ------------------------------------------------------------------------------------------
// K7: nearest un-pool.  out[b,d,y,x] = table[b, label[b, iy(y)*Wp + ix(x)], d]
// index rule = torch CPU upsample_nearest ('nearest'): identity, >>1, else
// min(floor(dst * float(in/out)), in-1) in fp32.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size, float scale) {
  if (out_size == in_size) return dst;
  if (out_size == 2 * in_size) return dst >> 1;
  const int s = (int)floorf(__fmul_rn((float)dst, scale));
  return min(s, in_size - 1);
}

template <typename TO>
struct Pack;
template <>
struct Pack<float> {
  static constexpr int VEC = 4;
  static __device__ __forceinline__ uint4 make(const float* v) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
};
template <>
struct Pack<__nv_bfloat16> {
  static constexpr int VEC = 8;
  static __device__ __forceinline__ uint4 make(const float* v) {
    uint4 r;
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]),
                   c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    r.x = *reinterpret_cast<unsigned*>(&a); r.y = *reinterpret_cast<unsigned*>(&b);
    r.z = *reinterpret_cast<unsigned*>(&c); r.w = *reinterpret_cast<unsigned*>(&d);
    return r;
  }
};

constexpr int kUnTX = 64, kUnTY = 4, kUnRY = 16, kUnDC = 8;

// vector path (W % VEC == 0, 16-byte aligned rows): grid (xblocks, yblocks, B * dchunks)
template <typename TO>
__global__ void __launch_bounds__(kUnTX* kUnTY) unpool_vec_kernel(const float* __restrict__ table,
                                                                 const int32_t* __restrict__ labels, int K, int D,
                                                                 int Hp, int Wp, int H, int W, TO* __restrict__ out,
                                                                 int64_t out_batch_stride, float sy, float sx) {
  constexpr int VEC = Pack<TO>::VEC;
  const int dchunks = ceil_div(D, kUnDC);
  const int b = blockIdx.z / dchunks, d0 = (blockIdx.z - b * dchunks) * kUnDC;
  const int nd = min(kUnDC, D - d0);
  const int xv = blockIdx.x * kUnTX + threadIdx.x;
  if (xv * VEC >= W) return;
  int px[VEC];
  bool same = true;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    px[v] = nearest_src(xv * VEC + v, Wp, W, sx);
    same = same && (px[v] == px[0]);
  }
  const int yend = min(H, (int)(blockIdx.y + 1) * kUnRY);
  const float* tb = table + (size_t)b * K * D;
  const int32_t* lb = labels ? labels + (size_t)b * Hp * Wp : nullptr;
  TO* ob = out + (size_t)b * out_batch_stride;
  int prev_py = -1;
  int lab[VEC];
  for (int y = blockIdx.y * kUnRY + threadIdx.y; y < yend; y += kUnTY) {
    const int py = nearest_src(y, Hp, H, sy);
    if (py != prev_py) {
      prev_py = py;
      if (same) {
        const int n = py * Wp + px[0];
        const int l = lb ? __ldg(lb + n) : n;
#pragma unroll
        for (int v = 0; v < VEC; ++v) lab[v] = l;
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const int n = py * Wp + px[v];
          lab[v] = lb ? __ldg(lb + n) : n;
        }
      }
    }
    TO* orow = ob + ((size_t)d0 * H + y) * W + (size_t)xv * VEC;
    if (same) {
      const float* tp = tb + (size_t)lab[0] * D + d0;
#pragma unroll
      for (int d = 0; d < kUnDC; ++d) {
        if (d < nd) {
          const float t = __ldg(tp + d);
          float vals[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) vals[v] = t;
          st_cs_v4(orow + (size_t)d * H * W, Pack<TO>::make(vals));
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < kUnDC; ++d) {
        if (d < nd) {
          float vals[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) vals[v] = __ldg(tb + (size_t)lab[v] * D + d0 + d);
          st_cs_v4(orow + (size_t)d * H * W, Pack<TO>::make(vals));
        }
      }
    }
  }
}

// scalar path: any W / alignment
template <typename TO>
__global__ void unpool_scalar_kernel(const float* __restrict__ table, const int32_t* __restrict__ labels, int B, int K,
                                     int D, int Hp, int Wp, int H, int W, TO* __restrict__ out,
                                     int64_t out_batch_stride, float sy, float sx) {
  const int64_t total = (int64_t)B * D * H * W;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(t % W);
    const int y = (int)((t / W) % H);
    const int d = (int)((t / ((int64_t)W * H)) % D);
    const int b = (int)(t / ((int64_t)W * H * D));
    const int n = nearest_src(y, Hp, H, sy) * Wp + nearest_src(xx, Wp, W, sx);
    const int l = labels ? __ldg(labels + (size_t)b * Hp * Wp + n) : n;
    out[(size_t)b * out_batch_stride + ((size_t)d * H + y) * W + xx] = from_f32<TO>(__ldg(table + ((size_t)b * K + l) * D + d));
  }
}

template <typename TX, typename TO>
static int launch_pool(const void* x, int B, int C, int Hf, int Wf, int ph, int pw, void* out, cudaStream_t st) {
  constexpr int VEC = Vec16<TX>::N;
  const int Hp = ceil_div(Hf, ph), Wp = ceil_div(Wf, pw);
  const int lpp = pw / VEC;
  const bool fast = (pw % VEC == 0) && (Wf % VEC == 0) && lpp >= 1 && lpp <= 32 && (lpp & (lpp - 1)) == 0 &&
                    ((uintptr_t)x % 16 == 0);
  // MG_POOL_VARIANT=-1 forces the LDG kernel below (A/B runs); default 0 = bulk-copy staged kernel where it applies
  static const int variant = getenv("MG_POOL_VARIANT") ? atoi(getenv("MG_POOL_VARIANT")) : 0;
  // default: TMA-staged persistent kernel (at least one whole row per ring buffer, <= 8 column passes per warp)
  const int row_bytes = Wf * (int)sizeof(TX);
  static const int stages_env = getenv("MG_POOL_STAGES") ? atoi(getenv("MG_POOL_STAGES")) : 3;
  static const int chunk_env = getenv("MG_POOL_CHUNK") ? atoi(getenv("MG_POOL_CHUNK")) : 8192;
  {
    int stages = std::max(2, std::min(8, stages_env));
    int stage_bytes = std::max(row_bytes, std::max(1024, chunk_env)) / 128 * 128;
    stage_bytes = std::max(stage_bytes, (row_bytes + 127) / 128 * 128);
    while (stages > 2 && (size_t)kPtWarps * stages * (stage_bytes + 8) + kPtWarps * kPtFifo * 4 > (size_t)kPtSmemBytes) --stages;
    const size_t smem = (size_t)kPtWarps * stages * (stage_bytes + 8) + kPtWarps * kPtFifo * 4;
    if (fast && variant == 0 && smem <= (size_t)kPtSmemBytes && ceil_div(Wf / VEC, 32) <= kPtMaxPasses && lpp <= 32) {
      PoolTmaArgs A;
      A.x = x; A.out = out; A.C = C; A.Hf = Hf; A.Wf = Wf; A.ph = ph; A.pw = pw; A.Hp = Hp; A.Wp = Wp;
      A.rpc = std::max(1, std::min(ph, stage_bytes / row_bytes));
      A.nstrips = B * Hp * C;
      A.stages = stages;
      A.stage_bytes = stage_bytes;
      // MG_POOL_MMA=1 (opt-in, not yet run on hardware): tensor-core summation for bf16 maps with 16-pixel-wide patches
      static const int use_mma = getenv("MG_POOL_MMA") ? atoi(getenv("MG_POOL_MMA")) : 0;
      if constexpr (std::is_same<TX, __nv_bfloat16>::value) {
        if (use_mma && pw == 16 && Wf % 256 == 0 && Wf / 256 <= kPtMaxPasses) {
          A.counters = nullptr;
          auto mk = pool_patches_mma_kernel<TO>;
          if (cudaFuncSetAttribute(mk, cudaFuncAttributeMaxDynamicSharedMemorySize, kPtSmemBytes) != cudaSuccess) {
            set_error("mg_pool_patches: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
            return MG_ERR_CUDA;
          }
          mk<<<std::min(num_sms(), ceil_div(A.nstrips, kPtWarps)), kPtWarps * 32, smem, st>>>(A);
          return check_launch("pool_patches_mma_kernel");
        }
      }
      A.counters = pool_counters(st);
      auto kern = A.counters ? pool_patches_tma_kernel<TX, TO, true> : pool_patches_tma_kernel<TX, TO, false>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPtSmemBytes) != cudaSuccess) {
        set_error("mg_pool_patches: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
        return MG_ERR_CUDA;
      }
      const int grid = std::min(num_sms(), ceil_div(A.nstrips, kPtWarps));
      kern<<<grid, kPtWarps * 32, smem, st>>>(A);
      return check_launch("pool_patches_tma_kernel");
    }
  }
  if (fast) {
    const int cg = std::min(C, pool_cc());
    // warps per block: the count in 4..8 that wastes the fewest channel slots (ties -> more warps)
    int nw = std::min(cg, 8), best_waste = 1 << 30;
    for (int w = std::min(cg, 8); w >= std::min(cg, 4); --w) {
      const int waste = ceil_div(cg, w) * w - cg;
      if (waste < best_waste) { best_waste = waste; nw = w; }
    }
    const int xchunks = ceil_div(Wf / VEC, 32);
    const size_t smem = (size_t)(32 / lpp) * (cg + 1) * 4;
    dim3 grid(xchunks * ceil_div(C, cg), Hp, B);
    pool_patches_vec_kernel<TX, TO><<<grid, nw * 32, smem, st>>>(reinterpret_cast<const TX*>(x), C, Hf, Wf, ph, pw, Hp, Wp,
                                                               cg, xchunks, reinterpret_cast<TO*>(out));
    return check_launch("pool_patches_vec_kernel");
  }
  const int64_t total = (int64_t)B * C * Hp * Wp;
  const int grid = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 16);
  pool_patches_generic_kernel<TX, TO><<<grid, 256, 0, st>>>(reinterpret_cast<const TX*>(x), B, C, Hf, Wf, ph, pw, Hp, Wp,
                                                           reinterpret_cast<TO*>(out));
  return check_launch("pool_patches_generic_kernel");
}

template <typename TO>
static int launch_unpool(const float* table, const int32_t* labels, int B, int K, int D, int Hp, int Wp, int H, int W,
                         void* out, int64_t stride, cudaStream_t st) {
  constexpr int VEC = Pack<TO>::VEC;
  // torch: scale = float(in) / float(out)
  const float sy = (float)Hp / (float)H, sx = (float)Wp / (float)W;
  const bool vec_ok = (W % VEC == 0) && ((uintptr_t)out % 16 == 0) && ((stride * (int64_t)sizeof(TO)) % 16 == 0);
  if (vec_ok) {
    dim3 block(kUnTX, kUnTY);
    dim3 grid(ceil_div(W / VEC, kUnTX), ceil_div(H, kUnRY), B * ceil_div(D, kUnDC));
    MG_REQUIRE(grid.z <= 65535 && grid.y <= 65535, MG_ERR_INVALID, "mg_unpool_nearest: grid too large");
    unpool_vec_kernel<TO><<<grid, block, 0, st>>>(table, labels, K, D, Hp, Wp, H, W, reinterpret_cast<TO*>(out), stride, sy, sx);
    return check_launch("unpool_vec_kernel");
  }
  const int64_t total = (int64_t)B * D * H * W;
  const int grid = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 32);
  unpool_scalar_kernel<TO><<<grid, 256, 0, st>>>(table, labels, B, K, D, Hp, Wp, H, W, reinterpret_cast<TO*>(out), stride, sy, sx);
  return check_launch("unpool_scalar_kernel");
}

}  // namespace mg

using namespace mg;

extern "C" {

int mg_pool_patches(const void* x, int x_dtype, int B, int C, int Hf, int Wf, int ph, int pw, void* out, int out_dtype,
                    mg_stream_t stream) {
  MG_REQUIRE(x && out && B > 0 && C > 0 && Hf > 0 && Wf > 0 && ph > 0 && pw > 0, MG_ERR_INVALID,
             "mg_pool_patches: bad arguments");
  MG_REQUIRE(B <= 65535 && ceil_div(Hf, ph) <= 65535, MG_ERR_INVALID, "mg_pool_patches: grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == MG_F32 && out_dtype == MG_F32) return launch_pool<float, float>(x, B, C, Hf, Wf, ph, pw, out, st);
  if (x_dtype == MG_BF16 && out_dtype == MG_BF16)
    return launch_pool<__nv_bfloat16, __nv_bfloat16>(x, B, C, Hf, Wf, ph, pw, out, st);
  if (x_dtype == MG_BF16 && out_dtype == MG_F32) return launch_pool<__nv_bfloat16, float>(x, B, C, Hf, Wf, ph, pw, out, st);
  if (x_dtype == MG_F32 && out_dtype == MG_BF16) return launch_pool<float, __nv_bfloat16>(x, B, C, Hf, Wf, ph, pw, out, st);
  set_error("mg_pool_patches: unsupported dtypes %d -> %d", x_dtype, out_dtype);
  return MG_ERR_INVALID;
}

int64_t mg_segment_work_bytes(int B, int N, int D, int K) {
  if (B <= 0 || N <= 0 || D <= 0 || K <= 0) return 0;
  return segment_work_bytes(B, N, D, K);
}

int mg_segment_mean(const float* h, const int32_t* labels, int B, int N, int D, int K, float* out, int32_t* counts, void* work,
                    mg_stream_t stream) {
  MG_REQUIRE(h && labels && out && work && B > 0 && N > 0 && D > 0 && K > 0, MG_ERR_INVALID, "mg_segment_mean: bad arguments");
  return segment_reduce_launch(h, labels, B, N, D, K, 1, 1.f, out, counts, work, (cudaStream_t)stream);
}

int mg_unpool_nearest(const float* table, const int32_t* labels, int B, int K, int D, int Hp, int Wp, int H, int W,
                      void* out, int out_dtype, int64_t out_batch_stride, mg_stream_t stream) {
  MG_REQUIRE(table && out && B > 0 && K > 0 && D > 0 && Hp > 0 && Wp > 0 && H > 0 && W > 0, MG_ERR_INVALID,
             "mg_unpool_nearest: bad arguments");
  MG_REQUIRE(labels || K == Hp * Wp, MG_ERR_INVALID, "mg_unpool_nearest: labels==NULL requires K == Hp*Wp");
  MG_REQUIRE(out_batch_stride >= (int64_t)D * H * W, MG_ERR_INVALID, "mg_unpool_nearest: batch stride smaller than D*H*W");
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == MG_F32) return launch_unpool<float>(table, labels, B, K, D, Hp, Wp, H, W, out, out_batch_stride, st);
  if (out_dtype == MG_BF16)
    return launch_unpool<__nv_bfloat16>(table, labels, B, K, D, Hp, Wp, H, W, out, out_batch_stride, st);
  set_error("mg_unpool_nearest: unsupported out dtype %d", out_dtype);
  return MG_ERR_INVALID;
}

}  // extern "C"
