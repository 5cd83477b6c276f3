// Device code of the bulk-copy (TMA) staged patch-pooling kernels — included by pool_unpool.cu (launch code there) and,
// with MG_HOST_EMULATION, by tests/emu (the kernels compiled for the host: warps as lock-step host threads, mbarriers and
// bulk copies emulated), so their control flow can be checked without a GPU.
#pragma once
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

#ifndef MG_HOST_EMULATION   // tests/emu provides host versions of these (mbarrier / bulk copy / address helpers)
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
__device__ __forceinline__ void pt_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ uint64_t pt_policy_evict_first() {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
#endif

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

// Packed accumulation (sm_100 add.f32x2): four fp32 pair accumulators per lane take one 16-byte vector per call.
// bf16: a 32-bit word holds two values whose fp32 images are (w << 16) and (w & 0xffff0000): shift + mask + ONE packed
// add per word instead of two scalar adds; fp32: the vector is two pairs.  The summing code of this kernel is what
// limits it (two warps per scheduler), so instructions per byte matter.
#ifndef MG_HOST_EMULATION
__device__ __forceinline__ void pt_add2(unsigned long long& acc, uint32_t lo, uint32_t hi) {
  unsigned long long v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(v));
}
__device__ __forceinline__ float pt_fold2(unsigned long long a, unsigned long long b, unsigned long long c, unsigned long long d) {
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(c) : "l"(d));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(c));
  uint32_t lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(a));
  return __uint_as_float(lo) + __uint_as_float(hi);
}
template <typename T>
__device__ __forceinline__ void pt_acc16(unsigned long long (&q)[4], const uint4& u);
template <>
__device__ __forceinline__ void pt_acc16<__nv_bfloat16>(unsigned long long (&q)[4], const uint4& u) {
  pt_add2(q[0], u.x << 16, u.x & 0xffff0000u);
  pt_add2(q[1], u.y << 16, u.y & 0xffff0000u);
  pt_add2(q[2], u.z << 16, u.z & 0xffff0000u);
  pt_add2(q[3], u.w << 16, u.w & 0xffff0000u);
}
template <>
__device__ __forceinline__ void pt_acc16<float>(unsigned long long (&q)[4], const uint4& u) {
  pt_add2(q[0], u.x, u.y);
  pt_add2(q[1], u.z, u.w);
}
#endif

struct PoolTmaArgs {
  const void* x;
  void* out;
  int C, Hf, Wf, ph, pw, Hp, Wp;
  int rpc;          // rows per chunk (rpc * Wf * esz <= stage_bytes)
  int nstrips;      // B * Hp * C, strip s = (b * Hp + py) * C + c
  int stages;       // ring depth per warp
  int stage_bytes;  // multiple of 128
};

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

template <typename TX, typename TO>
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
    pt_fence_mbar_init();
  }
  __syncwarp();
  const uint64_t policy = pt_policy_evict_first();
  const int row_bytes = A.Wf * (int)sizeof(TX);
  const int rv = row_bytes >> 4;                              // 16-byte vectors per row
  const int nvec = A.Wf / VEC;
  const int passes = ceil_div(nvec, 32);
  const int lpp = A.pw / VEC;
  const float inv = 1.f / (float)(A.ph * A.pw);
  TO* out = reinterpret_cast<TO*>(A.out);

  // Strip order: warp w of CTA b takes strips b + (w + 8 i) * grid.  (A global strip counter and a tensor-core summation
  // were measured in round 2 and lost: 42.7 us and 50.2 us against 40.6 us, profiles/r2_pool_variants.md.)
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
#ifndef MG_HOST_EMULATION
        unsigned long long q[4] = {0ull, 0ull, 0ull, 0ull};     // packed fp32 pair accumulators (independent chains)
        int r = 0;
        for (; r + kPtUnroll <= nr; r += kPtUnroll) {
          uint4 u[kPtUnroll];
#pragma unroll
          for (int j = 0; j < kPtUnroll; ++j) u[j] = a0[(r + j) * rv];
#pragma unroll
          for (int j = 0; j < kPtUnroll; ++j) pt_acc16<TX>(q, u[j]);
        }
        if (r + 4 <= nr) {                                      // chunks of 4 rows (2 KB rows: 1024-wide bf16 maps)
          uint4 u[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) u[j] = a0[(r + j) * rv];
#pragma unroll
          for (int j = 0; j < 4; ++j) pt_acc16<TX>(q, u[j]);
          r += 4;
        }
        for (; r < nr; ++r) pt_acc16<TX>(q, a0[r * rv]);
        part = pt_fold2(q[0], q[1], q[2], q[3]);
#else
        for (int r = 0; r < nr; ++r) part += pt_sum16<TX>(a0[r * rv]);
#endif
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
      cc.s += cc.step;
      cc.load(A);
    }
  }
}

}  // namespace mg
