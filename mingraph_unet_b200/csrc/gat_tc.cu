// K4 on the tensor pipe — fused multi-head GAT layer forward for bf16 node features
// (MultiHeadGATLayer.forward, eval mode, model/gat/graph_attention.py:40-118,150-160).
//
// Same math as gat_fused_kernel (gat_kernels.cuh): one CSR gather of x_i serves every head,
//   z_j^h = sum_i p_ij^h x_i / (den_j^h + 1e-10),  out_j = mean_h / concat_h ELU(W_h z_j^h),
// but the per-head transform W_h z runs on the 5th-generation tensor cores:
//   * a persistent CTA owns tiles of 128 destination nodes;
//   * 16 gather warps (a group of in/8 lanes per destination, 16-byte bf16 row chunks, 4 rows in
//     flight per group) accumulate z in registers and write it to shared memory as the UMMA A
//     operand: K-major, 128-byte swizzle, one 16 KB block per (head, 32-wide K slice);
//   * W_h sits in shared memory as the K-major B operand (staged once per CTA);
//   * one elected thread issues tcgen05.mma (kind::tf32, M=128, N=F, K=8) into TMEM, one F-column
//     accumulator block per head, double-buffered over tiles (2*heads*F <= 512 columns);
//   * 4 epilogue warps read TMEM (tcgen05.ld 32x32b), apply ELU and the head mean / concat and
//     store the tile through a swizzled staging buffer with fully coalesced 16-byte writes,
//     overlapping the next tile's gather.
// Precision: x is bf16 (exact in tf32); z and W are read as tf32 (10-bit mantissa), accumulation
// is fp32 in TMEM.  Used only for bf16 storage (tolerance 2e-2, north_star); the fp32 path stays
// on the FP32 pipe to hold 1e-5.
#include <stdlib.h>

#include "gat_kernels.cuh"

namespace mg {

constexpr int kTcTile = 128;
constexpr int kTcGatherWarps = 16;
constexpr int kTcEpiWarp0 = kTcGatherWarps;                 // 4 epilogue warps; the first one also issues the MMAs
constexpr int kTcMmaWarp = kTcEpiWarp0;
constexpr int kTcThreads = (kTcGatherWarps + 4) * 32;       // 640 -> 96 registers per thread
constexpr int kTcHeader = 1024;                             // barriers + TMEM base pointer

struct GatTcArgs {
  const __nv_bfloat16* x;
  const int32_t* rowptr;
  const int32_t* col;
  const float* s;          // (N, 2*heads)
  const float* gmax;       // (G, heads)
  const float* W;          // (heads, F, in)
  void* out;
  int N, F, concat, out_bf16, nodes_per_graph, tmem_cols;
  int64_t E;
  float slope;
};

// ---- programmatic dependent launch: the four kernels of a layer are chained on one stream; each lets its successor
// start (prologue + loads of the layer's INPUTS) while it is still running, and the successor waits for its
// predecessor's outputs right before it first reads them -------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// RULE: whatever the predecessor grid wrote (u, s, gmax, gsrc) is read with COHERENT loads (ld_pre below), never
// with __ldg.  ptxas treats ld.global.nc as a load of immutable data and schedules it above griddepcontrol.wait (ACQBULK) —
// seen in the SASS of tc_scores_kernel<4,32>, whose loads of u sat right behind PREEXIT and raced with tc_u_kernel (wrong
// scores whenever tc_u_kernel ran long enough); a data dependency through the inline asm does not help, the wait has no operands.
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// (asm volatile: ordered behind the wait by the front end; ld.global.ca: an ordinary weak load, which ptxas keeps behind ACQBULK.
//  No line of these buffers is in this SM's L1 before the wait, so the L1-cached form is as safe as .cg and schedules like __ldg.)
__device__ __forceinline__ float ld_pre(const float* p) {
  float v;
  asm volatile("ld.global.ca.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ld_pre(const float2* p) {
  float2 v;
  asm volatile("ld.global.ca.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_pre(const float4* p) {
  float4 v;
  asm volatile("ld.global.ca.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// ---- PTX helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TC_WAIT_DONE;\n"
      "bra TC_WAIT_LOOP;\n"
      "TC_WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// same, with a short back-off between polls (roles whose wait is long: they must not steal issue slots)
__device__ __forceinline__ void tc_mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(64);
  }
}
__device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// K-major operand, 128-byte swizzle: 8-row atoms of 1024 B (SBO), LBO unused for swizzled K-major (encoded 1)
__device__ __forceinline__ uint64_t tc_desc_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset (16 B units)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tc_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tc_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed fp32x2 FMA (sm_100 FFMA2): d = a * b + d on two fp32 lanes held in one 64-bit register
__device__ __forceinline__ unsigned long long tc_pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
  return r;
}
__device__ __forceinline__ void tc_unpack2(unsigned long long v, float& lo, float& hi) {
  unsigned a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
  lo = __uint_as_float(a); hi = __uint_as_float(b);
}
__device__ __forceinline__ void tc_fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

// ELU for the bf16 path: exp(v) - 1 through the hardware ex2 (abs. error ~1e-7, far inside the 2e-2 budget); the
// accurate expm1f costs ~30 dependent instructions per element and made the 4 epilogue warps the bottleneck
__device__ __forceinline__ float tc_elu(float v) {
  return fmaxf(v, 0.f) + (tc_ex2(fminf(v, 0.f) * 1.4426950408889634f) - 1.f);      // branch-free
}

template <int NH>
__device__ __forceinline__ void load_scores(const float* p, float (&o)[NH]);
template <>
__device__ __forceinline__ void load_scores<1>(const float* p, float (&o)[1]) { o[0] = ld_pre(p); }
template <>
__device__ __forceinline__ void load_scores<2>(const float* p, float (&o)[2]) {
  const float2 v = ld_pre(reinterpret_cast<const float2*>(p));
  o[0] = v.x; o[1] = v.y;
}
template <>
__device__ __forceinline__ void load_scores<4>(const float* p, float (&o)[4]) {
  const float4 v = ld_pre(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}

// shared-memory carve-up (byte offsets from the 1024-aligned base)
struct TcSmem {
  int a_off, b_off, stage_off, total;
  int a_block, b_block, kblocks;    // bytes per (head, k-slice) block of A / B; k-slices per head
};
__host__ __device__ inline TcSmem tc_smem_layout(int heads, int in_dim, int F, bool staged) {
  TcSmem L;
  L.kblocks = in_dim / 32;
  L.a_block = kTcTile * 128;
  L.b_block = F * 128;
  int o = kTcHeader;
  L.a_off = o;     o += heads * L.kblocks * L.a_block;
  L.b_off = o;     o += heads * L.kblocks * L.b_block;
  o = (o + 1023) & ~1023;
  L.stage_off = o; o += staged ? kTcTile * F * 2 : 0;
  L.total = o + 1024;               // slack for the manual 1024-byte alignment of the dynamic base
  return L;
}

// ---- pre-pass 1: attention scalars s[n] = (x_n . u_src[h], x_n . u_tgt[h]) -------------------------------------
// u_src[h] = W_h^T a_h[:F], u_tgt[h] = W_h^T a_h[F:] come from a one-block kernel (which also resets the per-graph maxima).
template <int NH, int LPN>
__global__ void __launch_bounds__(256) tc_u_kernel(const float* __restrict__ W, const float* __restrict__ a, int F,
                                                   int num_graphs, float* __restrict__ u, float* __restrict__ gmax,
                                                   float* __restrict__ gsrc /* (NH) or null: max_i s_src[i][h] */) {
  // one block per head: thread = (f-partition, input column); all of a thread's loads are independent
  constexpr int IN = LPN * 8;
  constexpr int FP = 256 / IN;
  __shared__ float u_part[FP][2 * IN];
  const int h = blockIdx.x;
  pdl_launch_dependents();
  {
    const int i = threadIdx.x % IN, fp = threadIdx.x / IN;
    const float* Wh = W + (size_t)h * F * IN + i;
    const float* ah = a + (size_t)h * 2 * F;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 16
    for (int f = fp; f < F; f += FP) {
      const float w = __ldg(Wh + (size_t)f * IN);
      acc0 = fmaf(__ldg(ah + f), w, acc0);
      acc1 = fmaf(__ldg(ah + F + f), w, acc1);
    }
    u_part[fp][i] = acc0;
    u_part[fp][IN + i] = acc1;
  }
  if (h == 0) {
    for (int i = threadIdx.x; i < num_graphs * NH; i += blockDim.x) gmax[i] = -INFINITY;
    if (gsrc && threadIdx.x < NH) gsrc[threadIdx.x] = -INFINITY;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * IN; idx += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int fp = 0; fp < FP; ++fp) v += u_part[fp][idx];
    const int half = idx / IN, i = idx - half * IN;
    u[(size_t)(half * NH + h) * IN + i] = v;                    // rows: u_src[0..NH) then u_tgt[0..NH)
  }
}

template <int NH, int LPN>
__global__ void __launch_bounds__(256) tc_scores_kernel(const __nv_bfloat16* __restrict__ x, int N,
                                                        const float* __restrict__ u, float* __restrict__ s,
                                                        float* __restrict__ gsrc /* nullable */) {
  constexpr int IN = LPN * 8, NQ = 2 * NH;
  const float* u_s = u;                                        // (2*NH, IN) fp32, 2 KB at most: L1/L2 resident (read after the wait)
  // s = X U^T on mma.sync m16n8k16 (bf16 x bf16 -> f32): a warp takes 16 nodes per step.  x is bf16 already; u is split
  // into bf16 hi + lo parts (two MMAs), so the products carry ~16 mantissa bits of u and the sums are fp32.  The k index
  // of the MMA is a permutation of the feature index chosen so that every lane feeds whole 16-byte row chunks.
  constexpr int CH = IN / 32;                                  // 16-byte chunks per lane per row
  constexpr int STEPS = IN / 16;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  pdl_launch_dependents();
  // the first step's rows of x (an INPUT of the layer) are requested before this kernel waits for u (its predecessor's
  // output): the DRAM round trip overlaps the tail of tc_u_kernel and the fragment set-up below
  uint4 x0[CH], x1[CH];
  {
    const int base = wid * 16;
    if (base < N) {
      const int r0 = min(base + g, N - 1), r1 = min(base + g + 8, N - 1);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        x0[c] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)r0 * IN) + c * 4 + t);
        x1[c] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)r1 * IN) + c * 4 + t);
      }
    }
  }
  pdl_wait_prior_grid();                                       // u comes from tc_u_kernel: read it with ld_pre
  uint32_t bhi[STEPS][2], blo[STEPS][2];
#pragma unroll
  for (int m = 0; m < STEPS; ++m)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int jj = 2 * m + r;                                // pair register jj of this lane: chunk (jj/4)*4 + t, pair jj%4
      const int k = ((jj >> 2) * 4 + t) * 8 + 2 * (jj & 3);
      const float u0 = g < NQ ? ld_pre(u_s + g * IN + k) : 0.f, u1 = g < NQ ? ld_pre(u_s + g * IN + k + 1) : 0.f;
      const __nv_bfloat16 h0 = __float2bfloat16_rn(u0), h1 = __float2bfloat16_rn(u1);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(u0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(u1 - __bfloat162float(h1));
      bhi[m][r] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      blo[m][r] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
  float ms0 = -INFINITY, ms1 = -INFINITY;                       // running max of this lane's two score columns (rows it owns)
  for (int base = wid * 16; base < N; base += nw * 16) {
    if (base != wid * 16) {                                    // (the first step's rows are already in flight)
      const int r0 = min(base + g, N - 1), r1 = min(base + g + 8, N - 1);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        x0[c] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)r0 * IN) + c * 4 + t);
        x1[c] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)r1 * IN) + c * 4 + t);
      }
    }
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
    for (int m = 0; m < STEPS; ++m) {
      const int c = m >> 1;
      const uint32_t a0 = (m & 1) ? x0[c].z : x0[c].x, a2 = (m & 1) ? x0[c].w : x0[c].y;
      const uint32_t a1 = (m & 1) ? x1[c].z : x1[c].x, a3 = (m & 1) ? x1[c].w : x1[c].y;
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bhi[m][0]), "r"(bhi[m][1]));
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(blo[m][0]), "r"(blo[m][1]));
    }
    if (2 * t < NQ) {
      if (base + g < N) {
        *reinterpret_cast<float2*>(s + (size_t)(base + g) * NQ + 2 * t) = make_float2(c0, c1);
        ms0 = fmaxf(ms0, c0); ms1 = fmaxf(ms1, c1);
      }
      if (base + g + 8 < N) {
        *reinterpret_cast<float2*>(s + (size_t)(base + g + 8) * NQ + 2 * t) = make_float2(c2, c3);
        ms0 = fmaxf(ms0, c2); ms1 = fmaxf(ms1, c3);
      }
    }
  }
  // max_i s_src[i][h] (columns 0 .. NH-1 of s): the bound the edge-max pre-pass prunes with.  Lanes with equal t hold
  // the same two columns; one atomic per column and BLOCK (atomics on four addresses serialise in L2).
  if (gsrc) {
    __shared__ float red[8][NQ];
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      ms0 = fmaxf(ms0, __shfl_xor_sync(kFull, ms0, o));
      ms1 = fmaxf(ms1, __shfl_xor_sync(kFull, ms1, o));
    }
    const int warp = threadIdx.x >> 5;
    if (g == 0 && 2 * t < NQ) { red[warp][2 * t] = ms0; red[warp][2 * t + 1] = ms1; }
    __syncthreads();
    if (threadIdx.x < NH) {
      float v = red[0][threadIdx.x];
#pragma unroll
      for (int w = 1; w < 8; ++w) v = fmaxf(v, red[w][threadIdx.x]);
      if (v > -INFINITY) atomic_max_f32(gsrc + threadIdx.x, v);
    }
  }
}

// ---- pre-pass 2a (single graph): a LOWER bound of the edge maximum from every destination's first in-edge.  Together
// with gsrc it lets the full scan below skip every destination whose best possible edge, s_tgt[j] + max_i s_src[i],
// cannot exceed it: fp32 addition is monotone, so the bound holds bit for bit and the result stays the exact maximum.
template <int NH>
__global__ void __launch_bounds__(256) tc_edge_first_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                         const float* s_arg, int N, float* __restrict__ gmax) {
  constexpr int NQ = 2 * NH;
  const float* s = s_arg;                                      // (laundered by the wait below)
  __shared__ float red[8][NH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_launch_dependents();
  float m[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) m[h] = -INFINITY;
  const int stride = gridDim.x * blockDim.x;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int beg = 0, end = 0, c0 = -1;
  if (j < N) {                                                 // inputs of the layer: requested before the wait
    beg = __ldg(rowptr + j); end = __ldg(rowptr + j + 1);
    if (beg < end) c0 = __ldg(col + beg);
  }
  pdl_wait_prior_grid();                                       // s comes from tc_scores_kernel
  for (; j < N; j += stride) {
    if (c0 >= 0) {
      float sv[NH], st[NH];
      load_scores<NH>(s + (size_t)c0 * NQ, sv);
      load_scores<NH>(s + (size_t)j * NQ + NH, st);
#pragma unroll
      for (int h = 0; h < NH; ++h) m[h] = fmaxf(m[h], sv[h] + st[h]);
    }
    const int jn = j + stride;
    c0 = -1;
    if (jn < N) {
      beg = __ldg(rowptr + jn); end = __ldg(rowptr + jn + 1);
      if (beg < end) c0 = __ldg(col + beg);
    }
  }
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float v = warp_max(m[h]);
    if (lane == 0) red[warp][h] = v;
  }
  __syncthreads();
  if (threadIdx.x < NH) {
    float v = red[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < 8; ++w) v = fmaxf(v, red[w][threadIdx.x]);
    if (v > -INFINITY) atomic_max_f32(gmax + threadIdx.x, v);
  }
}

// ---- pre-pass 2: exact per-graph maximum of s_src[i] + s_tgt[j] over the edges (graph_attention.py:86) ----------
// Thread per destination, 32 consecutive destinations per warp step; a thread walks its in-edges four at a time with
// the four column loads and then the four 16-byte score gathers all in flight (the round-1 kernel put 8 lanes on a
// destination: one edge per lane, three dependent round trips per destination and a full-mask shuffle inside a loop
// that not every group of the last warp entered — a hang for N % 4 != 0).  Maxima are order-independent, so the warp
// reduction + one atomic max per (warp, graph, head) is deterministic.
template <int NH, int UN>
__global__ void __launch_bounds__(256) tc_edge_max_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                       const float* s_arg, int N, int nodes_per_graph,
                                                       float* __restrict__ gmax, const float* gsrc_arg /* nullable */) {
  constexpr int NQ = 2 * NH;
  const float* s = s_arg;                                      // (both laundered by the wait below)
  const float* gsrc = gsrc_arg;
  __shared__ float red[8][NH];
  __shared__ int red_g[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  pdl_launch_dependents();
  pdl_wait_prior_grid();                                       // s, gsrc come from tc_scores_kernel
  // single graph: what a destination must beat to matter, and the best any source can contribute (see
  // tc_edge_first_kernel).  gmax only grows while this kernel runs; a stale read merely prunes less.
  float lowb[NH], gs[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    lowb[h] = gsrc ? __ldcg(gmax + h) : -INFINITY;
    gs[h] = gsrc ? __ldcg(gsrc + h) : INFINITY;
  }
  // running maxima of the warp for graph g_run (flushed with one atomic per head when the graph changes / at the end:
  // atomics on the same four addresses serialise in L2, so there must be few of them)
  float m_run[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) m_run[h] = -INFINITY;
  int g_run = -1;
  for (int base = wid * 32; base < N; base += nw * 32) {       // warp-uniform trip count: the shuffles below are safe
    const int j = base + lane;
    const bool ok = j < N;
    float m[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) m[h] = -INFINITY;
    float st[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) st[h] = 0.f;
    bool scan = ok;
    if (ok) {
      load_scores<NH>(s + (size_t)j * NQ + NH, st);
      if (gsrc) {                                              // can any edge into j exceed the lower bound?
        bool may = false;
#pragma unroll
        for (int h = 0; h < NH; ++h) may = may || (st[h] + gs[h] > lowb[h]);
        scan = may;
      }
    }
    if (scan) {
      const int beg = __ldg(rowptr + j), end = __ldg(rowptr + j + 1);
      for (int k = beg; k < end; k += UN) {                    // UN column loads, then UN score gathers, all in flight
        int c[UN];
        // whole 16-byte groups of a row when it is aligned (fixed in-degree 4 / 8 / ...): a thread's 4 column ids in one load, and a
        // warp's 32 loads cover 8 lines of 128 bytes instead of 32 (the pre-pass is bound by L1 wavefronts, not by bytes)
        if (((reinterpret_cast<uintptr_t>(col + k) & 15) == 0) && k + UN <= end) {
#pragma unroll
          for (int e = 0; e < UN; e += 4) {
            const int4 v = __ldg(reinterpret_cast<const int4*>(col + k + e));
            c[e] = v.x; c[e + 1] = v.y; c[e + 2] = v.z; c[e + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < UN; ++e) c[e] = k + e < end ? __ldg(col + k + e) : -1;
        }
        float sv[UN][NH];
#pragma unroll
        for (int e = 0; e < UN; ++e) {
#pragma unroll
          for (int h = 0; h < NH; ++h) sv[e][h] = -INFINITY;
          if (c[e] >= 0) load_scores<NH>(s + (size_t)c[e] * NQ, sv[e]);
        }
#pragma unroll
        for (int e = 0; e < UN; ++e)
#pragma unroll
          for (int h = 0; h < NH; ++h) m[h] = fmaxf(m[h], sv[e][h]);
      }
#pragma unroll
      for (int h = 0; h < NH; ++h) m[h] += st[h];                                 // -inf stays -inf for isolated nodes
    }
    const int g = (nodes_per_graph > 0 && ok) ? j / nodes_per_graph : 0;
    const int g_first = __shfl_sync(kFull, g, 0);
    const bool uniform = __all_sync(kFull, !ok || g == g_first);
    if (uniform) {
      if (g_first != g_run) {                                  // warp-uniform: flush the previous graph's maxima
        if (g_run >= 0 && lane == 0) {
#pragma unroll
          for (int h = 0; h < NH; ++h)
            if (m_run[h] > -INFINITY) atomic_max_f32(gmax + (size_t)g_run * NH + h, m_run[h]);
        }
#pragma unroll
        for (int h = 0; h < NH; ++h) m_run[h] = -INFINITY;
        g_run = g_first;
      }
#pragma unroll
      for (int h = 0; h < NH; ++h) m_run[h] = fmaxf(m_run[h], warp_max(m[h]));
    } else if (ok) {                                           // a graph boundary inside the warp's 32 nodes (rare)
#pragma unroll
      for (int h = 0; h < NH; ++h)
        if (m[h] > -INFINITY) atomic_max_f32(gmax + (size_t)g * NH + h, m[h]);
    }
  }
  // block level: when every warp ended on the same graph, one atomic per head for the whole block
  if (lane == 0) {
    red_g[warp] = g_run;
#pragma unroll
    for (int h = 0; h < NH; ++h) red[warp][h] = m_run[h];
  }
  __syncthreads();
  if (warp == 0) {
    const int gw = lane < 8 ? red_g[lane] : -1;
    const int g0 = __shfl_sync(kFull, gw, 0);
    const bool same = __all_sync(kFull, lane >= 8 || gw == g0 || gw < 0);
    if (same) {
      if (lane < NH && g0 >= 0) {
        float v = red[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) v = fmaxf(v, red_g[w] >= 0 ? red[w][lane] : -INFINITY);
        if (v > -INFINITY) atomic_max_f32(gmax + (size_t)g0 * NH + lane, v);
      }
    } else if (lane < 8 && gw >= 0) {
#pragma unroll
      for (int h = 0; h < NH; ++h)
        if (red[lane][h] > -INFINITY) atomic_max_f32(gmax + (size_t)gw * NH + h, red[lane][h]);
    }
  }
}

// ---- pre-pass 2b (single graph, pruned): only destinations that can still raise the maximum are scanned, and they are
// scanned by whole warps.  A block takes 256 consecutive destinations: every thread tests its destination against the
// bound (s_tgt[j] + max_i s_src[i] > lower bound, any head), the survivors are compacted into a shared-memory list, and
// each warp then walks the in-edges of one survivor at a time, 32 edges per pass (coalesced column load, 32 score gathers
// in flight).  A thread-per-destination scan gains nothing from pruning: one surviving lane keeps its warp in the loop.
template <int NH>
__global__ void __launch_bounds__(256) tc_edge_max_pruned_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                              const float* s_arg, int N, float* __restrict__ gmax,
                                                              const float* gsrc_arg) {
  constexpr int NQ = 2 * NH;
  const float* s = s_arg;                                      // (both laundered by the wait below)
  const float* gsrc = gsrc_arg;
  __shared__ int list[256];
  __shared__ int wcount[8];
  __shared__ float red[8][NH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_launch_dependents();
  pdl_wait_prior_grid();                                       // s, gsrc, the first-edge lower bound in gmax
  float lowb[NH], gs[NH], run[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    lowb[h] = __ldcg(gmax + h);                                // gmax only grows while this kernel runs: a stale read prunes less
    gs[h] = __ldcg(gsrc + h);
    run[h] = -INFINITY;
  }
  for (int base = blockIdx.x * 256; base < N; base += gridDim.x * 256) {
    const int j = base + threadIdx.x;
    bool may = false;
    if (j < N) {
      float st[NH];
      load_scores<NH>(s + (size_t)j * NQ + NH, st);
#pragma unroll
      for (int h = 0; h < NH; ++h) may = may || (st[h] + gs[h] > lowb[h]);
    }
    const unsigned bal = __ballot_sync(kFull, may);
    if (lane == 0) wcount[warp] = __popc(bal);
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < warp) off += wcount[w];
      total += wcount[w];
    }
    if (may) list[off + __popc(bal & ((1u << lane) - 1u))] = j;
    __syncthreads();
    for (int idx = warp; idx < total; idx += 8) {              // warp-uniform trip count
      const int jj = list[idx];
      const int beg = __ldg(rowptr + jj), end = __ldg(rowptr + jj + 1);
      float m[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) m[h] = -INFINITY;
      for (int k = beg + lane; k < end; k += 32) {
        float sv[NH];
        load_scores<NH>(s + (size_t)__ldg(col + k) * NQ, sv);
#pragma unroll
        for (int h = 0; h < NH; ++h) m[h] = fmaxf(m[h], sv[h]);
      }
      float st[NH];
      load_scores<NH>(s + (size_t)jj * NQ + NH, st);
#pragma unroll
      for (int h = 0; h < NH; ++h) run[h] = fmaxf(run[h], warp_max(m[h]) + st[h]);    // -inf stays -inf for isolated nodes
    }
    __syncthreads();                                           // the list is reused by the next chunk
  }
  if (lane == 0) {
#pragma unroll
    for (int h = 0; h < NH; ++h) red[warp][h] = run[h];
  }
  __syncthreads();
  if (threadIdx.x < NH) {
    float v = red[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < 8; ++w) v = fmaxf(v, red[w][threadIdx.x]);
    if (v > -INFINITY) atomic_max_f32(gmax + threadIdx.x, v);
  }
}

// NH heads (1, 2, 4), LPN lanes per destination node (in_dim = 8 * LPN bf16 = 16 B per lane)
template <int NH, int LPN>
__global__ void __launch_bounds__(kTcThreads, 1) gat_tc_kernel(const GatTcArgs A) {
  constexpr int IN = LPN * 8;
  constexpr int NPW = 32 / LPN;                              // destination nodes per warp per pass
  constexpr int NODES_PER_PASS = kTcGatherWarps * NPW;
  constexpr int PASSES = kTcTile / NODES_PER_PASS;
  constexpr int EPI = LPN < 8 ? LPN : 8;                     // edges (= source rows in flight) per group iteration
  constexpr int KB = IN / 32;                                // 32-wide K slices per head
  static_assert(PASSES >= 1 && NH * IN <= 256, "tile does not fit");

  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  const bool staged = A.out_bf16 && !A.concat;
  const TcSmem L = tc_smem_layout(NH, IN, A.F, staged);
  const uint32_t bar_a_empty = tc_smem_u32(sm + 8);
  const uint32_t bar_t_full0 = tc_smem_u32(sm + 16), bar_t_empty0 = tc_smem_u32(sm + 32);   // [2] each, 8 B apart
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(sm + 64);
  int* arrive_cnt = reinterpret_cast<int*>(sm + 72);        // gather warps that finished their share of a tile
  unsigned char* As = sm + L.a_off;
  unsigned char* Bs = sm + L.b_off;
  unsigned char* Ss = sm + L.stage_off;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F = A.F;
  const int ntiles = ceil_div(A.N, kTcTile);

  if (tid == 0) {
    *arrive_cnt = 0;
    tc_mbar_init(bar_a_empty, 1);
    tc_mbar_init(bar_t_full0, 1);
    tc_mbar_init(bar_t_full0 + 8, 1);
    tc_mbar_init(bar_t_empty0, 4);
    tc_mbar_init(bar_t_empty0 + 8, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kTcMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_ptr_s)),
                 "r"(A.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // W -> B operand blocks: element (h, f, i) -> block (h, i/32), row f, 16-byte chunk ((i%32)/4) ^ (f&7)
  for (int idx = tid; idx < NH * F * (IN / 4); idx += kTcThreads) {
    const int i4 = idx % (IN / 4);
    const int f = (idx / (IN / 4)) % F;
    const int h = idx / ((IN / 4) * F);
    const float4 w = __ldg(reinterpret_cast<const float4*>(A.W + ((size_t)h * F + f) * IN) + i4);
    const int kb = i4 >> 3, chunk = i4 & 7;
    *reinterpret_cast<float4*>(Bs + (size_t)(h * KB + kb) * L.b_block + f * 128 + ((chunk ^ (f & 7)) << 4)) = w;
  }
  tc_fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const float* s_in = A.s;
  const float* gmax_in = A.gmax;
  pdl_wait_prior_grid();                                       // s, gmax of the pre-pass are read from here on (ld_pre)

  // instruction descriptor: D=f32, A=B=tf32, both K-major, N=F, M=128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(F >> 3) << 17) | ((uint32_t)(kTcTile >> 4) << 24);

  auto issue_mma = [&](int it) {                            // one thread: tile `it` of this CTA, A operand complete
    const int stg = it & 1, n = it >> 1;
    if (n > 0) tc_mbar_wait(bar_t_empty0 + 8 * stg, (uint32_t)((n - 1) & 1));
    tc_fence_after();
#pragma unroll 1
    for (int h = 0; h < NH; ++h) {
      const uint32_t d_tmem = tmem_base + (uint32_t)(stg * NH * F + h * F);
#pragma unroll 1
      for (int kb = 0; kb < KB; ++kb) {
        const uint64_t ad = tc_desc_sw128(tc_smem_u32(As + (size_t)(h * KB + kb) * L.a_block));
        const uint64_t bd = tc_desc_sw128(tc_smem_u32(Bs + (size_t)(h * KB + kb) * L.b_block));
#pragma unroll
        for (int k = 0; k < 4; ++k)                         // UMMA_K = 8 tf32 = 32 bytes = 2 descriptor units
          tc_mma_tf32(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
      }
    }
    tc_commit(bar_a_empty);                                 // A may be overwritten once these MMAs retire
    tc_commit(bar_t_full0 + 8 * stg);                       // accumulators of this tile are complete
  };


  if (warp < kTcGatherWarps) {
    // =========================== gather warps: z tile -> A operand ===========================
    // A group of LPN lanes owns one destination node; lane gl holds dims [8*gl, 8*gl+8) of every head's z.
    // Score step: lane gl < EPI computes the attention numerators of edge slot gl for all heads; the group
    // then streams the EPI source rows (16 bytes per lane per row, all in flight together).
    const int gl = lane % LPN, grp = lane / LPN;
    constexpr int two_h = 2 * NH;
    constexpr float kLog2e = 1.4426950408889634f;
    // passes are numbered linearly over this CTA's tiles: pc = it * PASSES + pass
    auto node_at = [&](int pc) {
      const int t = (int)blockIdx.x + (pc / PASSES) * (int)gridDim.x;
      const int j = t * kTcTile + (pc % PASSES) * NODES_PER_PASS + warp * NPW + grp;
      return (t < ntiles && j < A.N) ? j : -1;
    };
    // software pipeline: row pointers two passes ahead, the first column chunk one pass ahead, so that a pass
    // starts with its source-row loads instead of a rowptr -> col -> row chain of three dependent round trips
    int beg = 0, end = 0, src0 = 0, nbeg = 0, nend = 0;
    {
      const int j0 = node_at(0), j1 = node_at(1);
      if (j0 >= 0) { beg = __ldg(A.rowptr + j0); end = __ldg(A.rowptr + j0 + 1); }
      if (j1 >= 0) { nbeg = __ldg(A.rowptr + j1); nend = __ldg(A.rowptr + j1 + 1); }
      if (gl < EPI && beg + gl < end) src0 = __ldg(A.col + beg + gl);
    }
    int pc = 0;
    for (int it = 0, tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int tile_base = tile * kTcTile;
#pragma unroll 1
      for (int pass = 0; pass < PASSES; ++pass, ++pc) {
        const int q = pass * NODES_PER_PASS + warp * NPW + grp;
        const int j = tile_base + q;
        const bool node_ok = j < A.N;
        // prefetch: first column chunk of pass pc+1, row pointers of pass pc+2
        int nsrc = 0, n2beg = 0, n2end = 0;
        if (gl < EPI && nbeg + gl < nend) nsrc = __ldg(A.col + nbeg + gl);
        {
          const int j2 = node_at(pc + 2);
          if (j2 >= 0) { n2beg = __ldg(A.rowptr + j2); n2end = __ldg(A.rowptr + j2 + 1); }
        }
        const int g = (A.nodes_per_graph > 0 && node_ok) ? j / A.nodes_per_graph : 0;
        float stgt[NH], Mh[NH], den[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          stgt[h] = 0.f; Mh[h] = 0.f; den[h] = 0.f;
        }
        if (node_ok && gl < EPI) {
          load_scores<NH>(s_in + (size_t)j * two_h + NH, stgt);
          load_scores<NH>(gmax_in + (size_t)g * NH, Mh);
#pragma unroll
          for (int h = 0; h < NH; ++h) Mh[h] = leaky_relu(Mh[h], A.slope);
        }
        unsigned long long z2[NH][4];                          // z[h][2i], z[h][2i+1] packed: accumulated with FFMA2
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
          for (int d = 0; d < 4; ++d) z2[h][d] = 0ull;
        const int iters = (end - beg + EPI - 1) / EPI;
        const int max_iters = __reduce_max_sync(kFull, iters);
        for (int itr = 0; itr < max_iters; ++itr) {
          const int k0 = beg + itr * EPI;
          const int ke = k0 + gl;
          const bool ev = gl < EPI && ke < end;
          const int srcn = itr == 0 ? src0 : (ev ? __ldg(A.col + ke) : 0);
          // source rows first (their address only needs col), then the scores
          uint4 xv[EPI];
#pragma unroll
          for (int e = 0; e < EPI; ++e) {
            const int sn = __shfl_sync(kFull, srcn, e, LPN);
            xv[e] = make_uint4(0u, 0u, 0u, 0u);
            if (k0 + e < end) xv[e] = __ldg(reinterpret_cast<const uint4*>(A.x + (size_t)sn * IN) + gl);
          }
          float pv[NH];
          {
            float ssrc[NH];
            load_scores<NH>(s_in + (size_t)srcn * two_h, ssrc);
#pragma unroll
            for (int h = 0; h < NH; ++h) {
              const float e = leaky_relu(ssrc[h] + stgt[h], A.slope);
              pv[h] = ev ? tc_ex2((e - Mh[h]) * kLog2e) : 0.f;
              den[h] += pv[h];
            }
          }
#pragma unroll
          for (int e = 0; e < EPI; ++e) {
            unsigned long long xf2[4];                          // bf16 pairs -> fp32 pairs (a shift and a mask per word)
            xf2[0] = tc_pack2(__uint_as_float(xv[e].x << 16), __uint_as_float(xv[e].x & 0xffff0000u));
            xf2[1] = tc_pack2(__uint_as_float(xv[e].y << 16), __uint_as_float(xv[e].y & 0xffff0000u));
            xf2[2] = tc_pack2(__uint_as_float(xv[e].z << 16), __uint_as_float(xv[e].z & 0xffff0000u));
            xf2[3] = tc_pack2(__uint_as_float(xv[e].w << 16), __uint_as_float(xv[e].w & 0xffff0000u));
#pragma unroll
            for (int h = 0; h < NH; ++h) {
              const float ph = __shfl_sync(kFull, pv[h], e, LPN);
              const unsigned long long ph2 = tc_pack2(ph, ph);
#pragma unroll
              for (int d = 0; d < 4; ++d) tc_fma2(z2[h][d], ph2, xf2[d]);
            }
          }
        }
        float z[NH][8];
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
          for (int d = 0; d < 4; ++d) tc_unpack2(z2[h][d], z[h][2 * d], z[h][2 * d + 1]);
        // softmax denominators: sum over the EPI edge-slot lanes (lanes >= EPI hold 0), broadcast from lane 0
        float inv[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
#pragma unroll
          for (int o = 1; o < EPI; o <<= 1) den[h] += __shfl_xor_sync(kFull, den[h], o, LPN);
          inv[h] = 1.f / (__shfl_sync(kFull, den[h], 0, LPN) + 1e-10f);                        // graph_attention.py:96
        }
        if (it > 0 && pass == 0) tc_mbar_wait(bar_a_empty, (uint32_t)((it - 1) & 1));          // previous tile's MMAs done
        // row q of the A operand: head h, K columns [gl*8, gl*8+8) = 2 chunks of slice gl/4
        const int kb = gl >> 2, c0 = (gl & 3) * 2, sw = q & 7;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          unsigned char* rowp = As + (size_t)(h * KB + kb) * L.a_block + q * 128;
          const float iv = inv[h];
          *reinterpret_cast<float4*>(rowp + (((c0) ^ sw) << 4)) =
              make_float4(z[h][0] * iv, z[h][1] * iv, z[h][2] * iv, z[h][3] * iv);
          *reinterpret_cast<float4*>(rowp + (((c0 + 1) ^ sw) << 4)) =
              make_float4(z[h][4] * iv, z[h][5] * iv, z[h][6] * iv, z[h][7] * iv);
        }
        beg = nbeg; end = nend; src0 = nsrc; nbeg = n2beg; nend = n2end;
      }
      tc_fence_async_smem();                                  // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      // the LAST gather warp to finish the tile issues its MMAs (no hand-off to another warp on the critical path)
      int last = 0;
      if (lane == 0) {
        __threadfence_block();
        int prev;                                             // (shared-space atomic: atomicAdd on the generic pointer compiles to ATOM.E)
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(prev) : "r"(tc_smem_u32(arrive_cnt)) : "memory");
        last = prev == kTcGatherWarps * (it + 1) - 1;
        __threadfence_block();
        if (last) issue_mma(it);
      }
      __syncwarp();
    }
  } else {
    // ============ epilogue warps (TMEM -> ELU -> mean/concat -> global); the first one also issues the MMAs ============
    const int ew = warp & 3;                                  // TMEM lane quarter this warp may access
    const int row = ew * 32 + lane;
    const int et = (warp - kTcEpiWarp0) * 32 + lane;          // 0..127 linear id among epilogue threads
    const float inv_h = 1.f / (float)NH;
    const int out_w = A.concat ? NH * F : F;
    const int cpr = F / 8;                                    // 16-byte chunks per staged bf16 row
    const int swz_mask = ((cpr & (cpr - 1)) == 0) ? (min(cpr, 8) - 1) : 0;
    auto epilogue_tile = [&](int it, int tile) {
      const int stg = it & 1, n = it >> 1;
      const int tile_base = tile * kTcTile;
      const int node = tile_base + row;
      tc_mbar_wait_relaxed(bar_t_full0 + 8 * stg, (uint32_t)(n & 1));
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(stg * NH * F);
      // 16 accumulator columns at a time (F % 16 == 0): fully unrolled, no per-element guards, two heads' loads in flight
      for (int c0 = 0; c0 < F; c0 += 16) {
        float oacc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) oacc[i] = 0.f;
#pragma unroll
        for (int h0 = 0; h0 < NH; h0 += 2) {
          uint32_t v[2][16];
          tc_tmem_ld16(t_row + (uint32_t)(h0 * F + c0), v[0]);
          if (h0 + 1 < NH) tc_tmem_ld16(t_row + (uint32_t)((h0 + 1) * F + c0), v[1]);
          tc_tmem_wait_ld();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (h0 + hh >= NH) break;
            float e16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) e16[i] = tc_elu(__uint_as_float(v[hh][i]));          // ELU per head (:118)
            if (A.concat) {
              if (node < A.N) {
                const size_t o = (size_t)node * out_w + (size_t)(h0 + hh) * F + c0;
                if (A.out_bf16) {
                  uint4 pk[2];
                  unsigned* pw = reinterpret_cast<unsigned*>(pk);
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    __nv_bfloat162 b2 = __floats2bfloat162_rn(e16[2 * i], e16[2 * i + 1]);
                    pw[i] = *reinterpret_cast<unsigned*>(&b2);
                  }
                  uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(A.out) + o);
                  op[0] = pk[0]; op[1] = pk[1];
                } else {
                  float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(A.out) + o);
#pragma unroll
                  for (int i = 0; i < 4; ++i) op[i] = make_float4(e16[4 * i], e16[4 * i + 1], e16[4 * i + 2], e16[4 * i + 3]);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) oacc[i] += e16[i];                                  // then the head mean (:158)
            }
          }
        }
        if (!A.concat) {
          if (staged) {
            uint4 pk[2];
            unsigned* pw = reinterpret_cast<unsigned*>(pk);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 b2 = __floats2bfloat162_rn(oacc[2 * i] * inv_h, oacc[2 * i + 1] * inv_h);
              pw[i] = *reinterpret_cast<unsigned*>(&b2);
            }
            const int chunk = c0 >> 3;
            *reinterpret_cast<uint4*>(Ss + (size_t)row * F * 2 + ((chunk ^ (row & swz_mask)) << 4)) = pk[0];
            *reinterpret_cast<uint4*>(Ss + (size_t)row * F * 2 + (((chunk + 1) ^ (row & swz_mask)) << 4)) = pk[1];
          } else if (node < A.N) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(A.out) + (size_t)node * out_w + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              op[i] = make_float4(oacc[4 * i] * inv_h, oacc[4 * i + 1] * inv_h, oacc[4 * i + 2] * inv_h, oacc[4 * i + 3] * inv_h);
          }
        }
      }
      // TMEM stage drained: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(bar_t_empty0 + 8 * stg);
      if (staged) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // the tile's rows are contiguous in the output: coalesced 16-byte stores
        const int valid_rows = min(kTcTile, A.N - tile_base);
        const int nchunks = valid_rows * cpr;
        uint4* gdst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(A.out) + (size_t)tile_base * F);
        for (int idx = et; idx < nchunks; idx += 128) {
          const int r = idx / cpr, c = idx - r * cpr;
          gdst[idx] = *reinterpret_cast<const uint4*>(Ss + (size_t)r * F * 2 + ((c ^ (r & swz_mask)) << 4));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    };

    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) epilogue_tile(it, tile);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kTcMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(A.tmem_cols) : "memory");
  }
}


// =====================================================================================================================
// Aggregation ON THE TENSOR CORES (heads = 4, in = 64; round 2).  The FFMA2 gather warps above spend ~280 warp
// instructions per destination (bf16 -> fp32 conversion, one FFMA2 per (edge, head, feature pair), a shuffle per (edge,
// head)); at N = 262 144, k = 8 that alone is 63 us of issue slots.  Here the weighted sum itself is an MMA:
//   rows    (destination d, head h) of 4 destinations x 4 heads = 16          (M)
//   columns edge slots: 8 per destination, two destinations per k-step        (K = 16)
//   A       block-diagonal attention numerators p[d][h][e] (bf16 hi + lo parts: 16 mantissa bits, two MMAs)
//   B       the 32 gathered source rows (bf16, exact), staged in shared memory by cp.async (zero-filled where a slot
//           has no edge) and read back with ldmatrix.trans
//   C       z[(d,h)][0..64) in fp32 fragments: 32 mma.m16n8k16 per 4 destinations and 8 edges each
// so a destination costs ~45 warp instructions.  A warp owns "steps" of 4 consecutive destinations (step g of the CTA
// -> warp g % 12) and walks each step's edges in chunks of 8 per destination; source rows are double-buffered per warp
// and software-pipelined three deep (column indices of item n+2, rows + attention scalars of item n+1 in flight while
// item n runs on the tensor cores).  The normalised z goes to shared memory as the bf16 K-major UMMA A operand
// (128-byte swizzle), the transform W_h z runs on tcgen05 (kind::f16, M128 x N=F x K16, fp32 accumulators in TMEM,
// double-buffered over tiles) and the 4 epilogue warps are those of gat_tc_kernel.
// Precision: p carries 16 mantissa bits, x is exact, sums are fp32; z and W are rounded to bf16 for the transform
// (fp32 accumulate) — measured against the fp32 oracle in tests/test_gpu_tc.py (budget 2e-2).
// =====================================================================================================================
constexpr int kAgWarps = 12;
constexpr int kAgThreads = (kAgWarps + 4) * 32;            // 512 -> up to 128 registers per thread
constexpr int kAgIn = 64, kAgNH = 4;
constexpr int kAgRowBytes = kAgIn * 2;                      // 128
constexpr int kAgStageBytes = 32 * kAgRowBytes;             // 4 KB: 4 destinations x 8 edge slots
constexpr int kAgStages = 2;
constexpr int kAgTile = kTcTile;                            // destinations per tile (96 = two steps per warp was measured: slower, the epilogue warps become the limit)
constexpr int kAgStepsPerTile = kAgTile / 4;                // 32

constexpr int kAgAttDst = 144;                              // bytes per destination in the attention scratch: 8 slots x 4 heads + 16 pad (bank shift)
constexpr int kAgAttBytes = 4 * kAgAttDst;                  // per warp
struct AgSmem {
  int a_off, b_off, ring_off, att_off, stage_off, total;
};
__host__ __device__ inline AgSmem ag_smem_layout(int F, bool staged) {
  AgSmem L;
  int o = kTcHeader;
  L.a_off = o;     o += kAgNH * kTcTile * 128;              // 4 head blocks of 128 rows x 64 bf16
  L.b_off = o;     o += kAgNH * F * 128;
  L.ring_off = o;  o += kAgWarps * kAgStages * kAgStageBytes;
  L.att_off = o;   o += kAgWarps * kAgAttBytes;
  o = (o + 1023) & ~1023;
  L.stage_off = o; o += staged ? kAgTile * F * 2 : 0;
  L.total = o + 1024;
  return L;
}

__device__ __forceinline__ void ag_cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// base + row * 128 as ONE 64-bit multiply-add (left to itself the compiler splits it into a multiply, an OR and a two-instruction add)
__device__ __forceinline__ const char* ag_row_ptr(const char* lane_base, int row) {
  unsigned long long p;
  asm("mad.wide.u32 %0, %1, 128, %2;" : "=l"(p) : "r"((unsigned)row), "l"((unsigned long long)lane_base));
  return reinterpret_cast<const char*>(p);
}
__device__ __forceinline__ float ag_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// acc += ELU(a), ELU(b) on a packed pair: ELU(v) = max(v, exp(min(v, 0)) - 1) (for v > 0 the second term is 0; for v <= 0 it is
// >= v), the scale by log2(e), the -1 and the sum each one packed instruction for the two lanes
__device__ __forceinline__ void ag_elu2_acc(unsigned long long& acc, uint32_t a, uint32_t b) {
  asm("{\n.reg .b64 m, t;\n.reg .f32 ml, mh, tl, th;\n"
      "mov.b64 m, {%1, %2};\n"
      "mul.f32x2 m, m, %3;\n"
      "mov.b64 {ml, mh}, m;\n"
      "min.f32 ml, ml, 0f00000000;\nmin.f32 mh, mh, 0f00000000;\n"
      "ex2.approx.ftz.f32 tl, ml;\nex2.approx.ftz.f32 th, mh;\n"
      "mov.b64 t, {tl, th};\n"
      "add.f32x2 t, t, %4;\n"
      "mov.b64 {tl, th}, t;\n"
      "max.f32 tl, tl, %1;\nmax.f32 th, th, %2;\n"
      "mov.b64 t, {tl, th};\n"
      "add.f32x2 %0, %0, t;\n}"
      : "+l"(acc)
      : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "l"(0x3fb8aa3b3fb8aa3bull), "l"(0xbf800000bf800000ull));
}
__device__ __forceinline__ void ag_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ag_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ag_ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void ag_mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// same, accumulators start from zero (the first MMA of a single-chunk step: no zeroing pass over the 32 accumulators)
__device__ __forceinline__ void ag_mma_bf16_zero(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ uint32_t ag_pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// bf16 hi parts of (a, b) and the bf16 of the remainders
__device__ __forceinline__ void ag_split(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  hi = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
  lo = ag_pack_bf16(a - __bfloat162float(ah), b - __bfloat162float(bh));
}
// The 32 HMMAs of an item: the stage's 32 source rows (B operand through ldmatrix.trans) times the block-diagonal attention
// fragments (A operand, bf16 hi + lo parts).  k-step 0: slots of destinations 0 (k 0-7) and 1 (k 8-15) feed fragment rows 0-7
// only; k-step 1: destinations 2, 3 feed rows 8-15.  ZERO: the accumulators start from 0 (single-chunk step).
template <bool ZERO>
__device__ __forceinline__ void ag_item_mma(float (&acc)[8][4], uint32_t sb, const uint32_t (&ld_off)[4], uint32_t hiA, uint32_t loA,
                                            uint32_t hiB, uint32_t loB, uint32_t mask_d0, uint32_t mask_d1) {
  const uint32_t a0h = hiA & mask_d0, a0l = loA & mask_d0, a2h = hiA & mask_d1, a2l = loA & mask_d1;
#pragma unroll
  for (int np = 0; np < 4; ++np) {
    uint32_t b[4];
    ag_ldmatrix_x4_trans(sb + ld_off[np], b);
    if (ZERO) ag_mma_bf16_zero(acc[2 * np], a0h, 0u, a2h, 0u, b[0], b[1]);
    else ag_mma_bf16(acc[2 * np], a0h, 0u, a2h, 0u, b[0], b[1]);
    ag_mma_bf16(acc[2 * np], a0l, 0u, a2l, 0u, b[0], b[1]);
    if (ZERO) ag_mma_bf16_zero(acc[2 * np + 1], a0h, 0u, a2h, 0u, b[2], b[3]);
    else ag_mma_bf16(acc[2 * np + 1], a0h, 0u, a2h, 0u, b[2], b[3]);
    ag_mma_bf16(acc[2 * np + 1], a0l, 0u, a2l, 0u, b[2], b[3]);
  }
  const uint32_t a1h = hiB & mask_d0, a1l = loB & mask_d0, a3h = hiB & mask_d1, a3l = loB & mask_d1;
#pragma unroll
  for (int np = 0; np < 4; ++np) {
    uint32_t b[4];
    ag_ldmatrix_x4_trans(sb + 16 * kAgRowBytes + ld_off[np], b);
    ag_mma_bf16(acc[2 * np], 0u, a1h, 0u, a3h, b[0], b[1]);
    ag_mma_bf16(acc[2 * np], 0u, a1l, 0u, a3l, b[0], b[1]);
    ag_mma_bf16(acc[2 * np + 1], 0u, a1h, 0u, a3h, b[2], b[3]);
    ag_mma_bf16(acc[2 * np + 1], 0u, a1l, 0u, a3l, b[2], b[3]);
  }
}
// softmax denominators of the lane's two destinations over the 8 slots each (the 4 lanes of a fragment row hold 2 slots each)
__device__ __forceinline__ void ag_quad_sum(float& d0, float& d1) {
  d0 += __shfl_xor_sync(kFull, d0, 1); d1 += __shfl_xor_sync(kFull, d1, 1);
  d0 += __shfl_xor_sync(kFull, d0, 2); d1 += __shfl_xor_sync(kFull, d1, 2);
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}

__global__ void __launch_bounds__(kAgThreads, 1) gat_agg_mma_kernel(const GatTcArgs A) {
  constexpr int NH = kAgNH, IN = kAgIn;
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  const bool staged = A.out_bf16 && !A.concat;
  const AgSmem L = ag_smem_layout(A.F, staged);
  const uint32_t bar_a_empty = tc_smem_u32(sm + 8);
  const uint32_t bar_t_full0 = tc_smem_u32(sm + 16), bar_t_empty0 = tc_smem_u32(sm + 32);   // [2] each, 8 B apart
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(sm + 64);
  int* arrive_cnt = reinterpret_cast<int*>(sm + 72);        // steps of the CTA whose z rows are in the A operand
  unsigned char* As = sm + L.a_off;
  unsigned char* Bs = sm + L.b_off;
  unsigned char* Ss = sm + L.stage_off;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F = A.F;
  const int ntiles = ceil_div(A.N, kAgTile);

  if (tid == 0) {
    *arrive_cnt = 0;
    tc_mbar_init(bar_a_empty, 1);
    tc_mbar_init(bar_t_full0, 1);
    tc_mbar_init(bar_t_full0 + 8, 1);
    tc_mbar_init(bar_t_empty0, 4);
    tc_mbar_init(bar_t_empty0 + 8, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kAgWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_ptr_s)),
                 "r"(A.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // W (fp32) -> bf16 B operand: element (h, f, i) -> block h, row f, 16-byte chunk (i / 8) ^ (f & 7)
  for (int idx = tid; idx < NH * F * (IN / 8); idx += kAgThreads) {
    const int c8 = idx % (IN / 8);
    const int f = (idx / (IN / 8)) % F;
    const int h = idx / ((IN / 8) * F);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(A.W + ((size_t)h * F + f) * IN) + 2 * c8);
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(A.W + ((size_t)h * F + f) * IN) + 2 * c8 + 1);
    uint4 pk;
    pk.x = ag_pack_bf16(w0.x, w0.y); pk.y = ag_pack_bf16(w0.z, w0.w);
    pk.z = ag_pack_bf16(w1.x, w1.y); pk.w = ag_pack_bf16(w1.z, w1.w);
    *reinterpret_cast<uint4*>(Bs + (size_t)h * F * 128 + f * 128 + ((c8 ^ (f & 7)) << 4)) = pk;
  }
  tc_fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  // everything above (barriers, TMEM allocation, W -> bf16 B operand) ran beside the tail of the edge-max pre-pass;
  // its outputs (s, gmax) are read from here on, through these pointers
  const float* s_in = A.s;
  const float* gmax_in = A.gmax;
  pdl_wait_prior_grid();

  // instruction descriptor: D = f32, A = B = bf16, both K-major, N = F, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(F >> 3) << 17) | ((uint32_t)(kTcTile >> 4) << 24);

  auto issue_mma = [&](int it) {                            // one thread: tile `it` of this CTA, A operand complete
    const int stg = it & 1, n = it >> 1;
    if (n > 0) tc_mbar_wait(bar_t_empty0 + 8 * stg, (uint32_t)((n - 1) & 1));
    tc_fence_after();
#pragma unroll 1
    for (int h = 0; h < NH; ++h) {
      const uint32_t d_tmem = tmem_base + (uint32_t)(stg * NH * F + h * F);
      const uint64_t ad = tc_desc_sw128(tc_smem_u32(As + (size_t)h * kTcTile * 128));
      const uint64_t bd = tc_desc_sw128(tc_smem_u32(Bs + (size_t)h * F * 128));
#pragma unroll
      for (int k = 0; k < IN / 16; ++k)                     // UMMA_K = 16 bf16 = 32 bytes = 2 descriptor units
        tc_mma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
    }
    tc_commit(bar_a_empty);                                 // A may be overwritten once these MMAs retire
    tc_commit(bar_t_full0 + 8 * stg);                       // accumulators of this tile are complete
  };

  if (warp < kAgWarps) {
    // ======================= aggregation warps: gathered rows x attention -> z -> A operand =======================
    // (written flat, with every lane-constant address term hoisted: the first version spent ~950 instructions per item
    //  on index arithmetic, runtime divisions and pipeline bookkeeping for 32 HMMA)
    const int g8 = lane >> 2, t4 = lane & 3;                  // MMA fragment coordinates
    const int hd = g8 & 3, dsel = g8 >> 2;                    // fragment row g8 = (destination dsel, head hd); row g8 + 8 = (dsel + 2, hd)
    const int r4 = lane >> 3, ch8 = lane & 7;                 // loader role: destination of the step / row within a 4-row group, 16-byte chunk
    const uint32_t ring0 = tc_smem_u32(sm + L.ring_off + (size_t)warp * kAgStages * kAgStageBytes);
    constexpr float kLog2e = 1.4426950408889634f;
    const float slope = A.slope;
    const int npg = A.nodes_per_graph;
    const int tile_stride = (int)gridDim.x * kAgTile;
    const int n_end = ntiles * kAgTile;                       // a cursor is exhausted once its tile base reaches this
    // lane constants
    const uint32_t cp_off_even = (uint32_t)(r4 * kAgRowBytes + ((ch8 ^ r4) << 4));          // rows 8m + r4
    const uint32_t cp_off_odd = (uint32_t)(r4 * kAgRowBytes + ((ch8 ^ (r4 | 4)) << 4));      // rows 8m + 4 + r4
    const char* x_lane = reinterpret_cast<const char*>(A.x) + ch8 * 16;
    const float* s_tgt_lane = s_in + NH + hd;
    uint32_t ld_off[4];                                       // ldmatrix.x4.trans lane addresses for the 4 feature pairs (k-step 0)
#pragma unroll
    for (int np = 0; np < 4; ++np)
      ld_off[np] = (uint32_t)((((r4 & 1) * 8 + ch8) * kAgRowBytes) + ((((np << 1) | (r4 >> 1)) ^ ch8) << 4));
    const uint32_t mask_d0 = dsel == 0 ? 0xffffffffu : 0u, mask_d1 = ~mask_d0;
    const uint32_t z_lane = tc_smem_u32(As) + (uint32_t)(hd * kTcTile * 128 + dsel * 128 + t4 * 4);
        // (shared-space addresses: through generic pointers the compiler emits LD.E / ST.E, tracked on the long scoreboard)
    const uint32_t att_w = tc_smem_u32(sm + L.att_off + warp * kAgAttBytes + r4 * kAgAttDst + ch8 * 16);                 // slot (r4, ch8), heads 0..3
    const uint32_t att_r = tc_smem_u32(sm + L.att_off + warp * kAgAttBytes + dsel * kAgAttDst + t4 * 32 + hd * 4);
    const float M_single = npg > 0 ? 0.f : leaky_relu(ld_pre(gmax_in + hd), slope);   // one graph: the shift is a lane constant

    // ---- cursor over the warp's steps (warp-uniform), loader-lane row ranges ----
    int cur_tb = (int)blockIdx.x * kAgTile, cur_step = warp, cur_it = 0;     // tile base node, step in tile, tile ordinal
    int cur_c = 0, cur_nch = 1, cur_beg = 0, cur_end = 0, nxt_beg = 0, nxt_end = 0;
#define AG_NEXT_STEP(tb, st, it)  \
  do {                            \
    st += kAgWarps;               \
    if (st >= kAgStepsPerTile) {  \
      st -= kAgStepsPerTile;      \
      tb += tile_stride;          \
      ++it;                       \
    }                             \
  } while (0)
    // the loads write the pipeline registers themselves: a copy out of a temporary would wait for the data one
    // instruction after the request (measured: 7 % of the aggregation warps' time in that move)
#define AG_FETCH_RP(tb, st, b, e)                                                                  \
  do {                                                                                             \
    b = 0; e = 0;                                                                                  \
    const int j_ = tb + st * 4 + r4;                                                               \
    const int ok_ = (tb < n_end && j_ < A.N) ? 1 : 0;                                              \
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p ld.global.nc.b32 %0, [%3];\n"        \
                 "@p ld.global.nc.b32 %1, [%3+4];\n}"                                              \
                 : "+r"(b), "+r"(e) : "r"(ok_), "l"(A.rowptr + (ok_ ? j_ : 0)));                   \
  } while (0)
    AG_FETCH_RP(cur_tb, cur_step, cur_beg, cur_end);
    {
      int tb = cur_tb, st = cur_step, it = cur_it;
      AG_NEXT_STEP(tb, st, it);
      AG_FETCH_RP(tb, st, nxt_beg, nxt_end);
    }
    cur_nch = max(1, __reduce_max_sync(kFull, (cur_end - cur_beg + 7) >> 3));
    // advance the cursor by one item; rowptr of the step after the next one is prefetched
    // branch-free on the load: the prefetched row pointers are written by a predicated load straight into nxt_beg / nxt_end
    // (tied asm operands), so nothing waits for them before the next step reads them
#define AG_ADVANCE()                                                                                               \
  do {                                                                                                             \
    const bool adv_ = ++cur_c >= cur_nch;                                                                          \
    if (adv_) {                                                                                                    \
      AG_NEXT_STEP(cur_tb, cur_step, cur_it);                                                                      \
      cur_c = 0;                                                                                                   \
    }                                                                                                              \
    cur_beg = adv_ ? nxt_beg : cur_beg;                                                                            \
    cur_end = adv_ ? nxt_end : cur_end;                                                                            \
    {                                                                                                              \
      int tb = cur_tb, st = cur_step, it = cur_it;                                                                 \
      AG_NEXT_STEP(tb, st, it);                                                                                    \
      /* past the last node both loads read rowptr[N]: an empty range, no zeroing (a select would wait for the load) */ \
      const int j_ = tb < n_end ? min(tb + st * 4 + r4, A.N) : A.N;                                                \
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p ld.global.nc.b32 %0, [%3];\n@p ld.global.nc.b32 %1, [%4];\n}" \
                   : "+r"(nxt_beg), "+r"(nxt_end)                                                                  \
                   : "r"(adv_ ? 1 : 0), "l"(A.rowptr + j_), "l"(A.rowptr + min(j_ + 1, A.N)));                     \
    }                                                                                                              \
    if (adv_) cur_nch = max(1, __reduce_max_sync(kFull, (cur_end - cur_beg + 7) >> 3));                            \
  } while (0)
    // loader lane: source node of its slot in the cursor's item (-1: no edge)
#define AG_LOAD_SRC(dst)                                             \
  do {                                                               \
    const int k_ = cur_beg + 8 * cur_c + ch8;                        \
    dst = (cur_tb < n_end && k_ < cur_end) ? __ldg(A.col + k_) : -1; \
  } while (0)
    // 32 rows x 8 chunks of 16 B into ring stage `stg`; lanes 8r .. 8r+7 read one 128-byte row; empty slots are zero-filled
#define AG_ISSUE_ROWS(src, stg)                                                                       \
  do {                                                                                                \
    const uint32_t sb_ = ring0 + (uint32_t)(stg) * kAgStageBytes;                                      \
    _Pragma("unroll") for (int i_ = 0; i_ < 8; ++i_) {                                                 \
      const int sn_ = __shfl_sync(kFull, src, i_ * 4 + r4);                                            \
      ag_cp_async16(sb_ + (uint32_t)(i_ * 4 * kAgRowBytes) + ((i_ & 1) ? cp_off_odd : cp_off_even),   \
                    ag_row_ptr(x_lane, max(sn_, 0)), sn_ >= 0 ? 16u : 0u);                              \
    }                                                                                                 \
  } while (0)
    // attention scalars of an item: s_src of the lane's 4 slots (-inf: empty), s_tgt and shift of its two destinations
#define AG_LOAD_TGT(node0, st2, M2)                                                                    \
  do {                                                                                                 \
    const int ja_ = node0 + dsel, jb_ = ja_ + 2;                                                        \
    const bool va_ = node0 >= 0 && ja_ < A.N, vb_ = node0 >= 0 && jb_ < A.N;                             \
    st2[0] = va_ ? ld_pre(s_tgt_lane + (size_t)(unsigned)ja_ * (2 * NH)) : 0.f;                          \
    st2[1] = vb_ ? ld_pre(s_tgt_lane + (size_t)(unsigned)jb_ * (2 * NH)) : 0.f;                          \
    M2[0] = M_single; M2[1] = M_single;                                                                 \
    if (npg > 0) {                                                                                      \
      M2[0] = va_ ? leaky_relu(ld_pre(gmax_in + (size_t)(ja_ / npg) * NH + hd), slope) : 0.f;              \
      M2[1] = vb_ ? leaky_relu(ld_pre(gmax_in + (size_t)(jb_ / npg) * NH + hd), slope) : 0.f;              \
    }                                                                                                   \
  } while (0)
    // loader role: the lane's OWN slot, all four heads in one 16-byte gather (a quarter of the L1 requests of four 4-byte
    // gathers per fragment lane); the fragment lanes pick their values up from the warp's scratch in step E
#define AG_LOAD_ATT(src, node0, sv, st2, M2)                                                           \
  do {                                                                                                 \
    sv[0] = sv[1] = sv[2] = sv[3] = -INFINITY;                                                           \
    if (src >= 0) {                                                                                     \
      const float4 v_ = ld_pre(reinterpret_cast<const float4*>(s_in + (size_t)(unsigned)src * (2 * NH))); \
      sv[0] = v_.x; sv[1] = v_.y; sv[2] = v_.z; sv[3] = v_.w;                                            \
    }                                                                                                   \
    AG_LOAD_TGT(node0, st2, M2);                                                                        \
  } while (0)

    // ---- software pipeline: item n on the tensor cores, rows + scalars of item n+1 and columns of item n+2 in flight ----
    // item descriptor: node0 = first destination of its step (-1: past the end), row0 = step * 4, it, first / last chunk flags
    int node0_0 = cur_tb < n_end ? cur_tb + cur_step * 4 : -1, row0_0 = cur_step * 4, it_0 = cur_it;
    bool first_0 = true, last_0 = cur_nch == 1;
    int src0;
    AG_LOAD_SRC(src0);
    float sv0[4], st0[2], M0[2];
    AG_LOAD_ATT(src0, node0_0, sv0, st0, M0);
    AG_ISSUE_ROWS(src0, 0);
    ag_cp_commit();
    AG_ADVANCE();
    int node0_1 = cur_tb < n_end ? cur_tb + cur_step * 4 : -1, row0_1 = cur_step * 4, it_1 = cur_it;
    bool first_1 = cur_c == 0, last_1 = cur_c == cur_nch - 1;
    int src1;
    AG_LOAD_SRC(src1);
    float acc[8][4];
    float den0 = 0.f, den1 = 0.f;
    uint32_t stage = 0;
    while (node0_0 >= 0) {
      // A. rows of item n+1
      if (node0_1 >= 0) AG_ISSUE_ROWS(src1, stage ^ 1u);
      ag_cp_commit();
      // B. columns of item n+2
      AG_ADVANCE();
      const int node0_2 = cur_tb < n_end ? cur_tb + cur_step * 4 : -1, row0_2 = cur_step * 4, it_2 = cur_it;
      const bool first_2 = cur_c == 0, last_2 = cur_c == cur_nch - 1;
      int src2;
      AG_LOAD_SRC(src2);
      // C. attention scalars of item n+1
      float sv1[4], st1[2], M1[2];
      AG_LOAD_ATT(src1, node0_1, sv1, st1, M1);
      // D. rows of item n have landed
      ag_cp_wait<1>();
      __syncwarp();
      if (first_0) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        den0 = den1 = 0.f;
      }
      // E. attention numerators (graph_attention.py:61-65,86) and the block-diagonal A fragments
      float sq[4];
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(att_w), "f"(sv0[0]), "f"(sv0[1]), "f"(sv0[2]), "f"(sv0[3]) : "memory");
      __syncwarp();
      asm volatile("ld.shared.f32 %0, [%4];\nld.shared.f32 %1, [%4+16];\nld.shared.f32 %2, [%4+288];\nld.shared.f32 %3, [%4+304];"
                   : "=f"(sq[0]), "=f"(sq[1]), "=f"(sq[2]), "=f"(sq[3]) : "r"(att_r) : "memory");
      static_assert(2 * kAgAttDst == 288, "slot B = slot A + 2 destinations");
      // (the scratch is rewritten one iteration later, behind the __syncwarp() that follows the MMAs)
      float p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float e = sq[i] + st0[i >> 1];
        p[i] = tc_ex2((fmaxf(e, e * slope) - M0[i >> 1]) * kLog2e);                 // LeakyReLU (0 <= slope <= 1); empty slot: ex2(-inf) = 0
      }
      den0 += p[0] + p[1];
      den1 += p[2] + p[3];
      uint32_t hiA, loA, hiB, loB;
      ag_split(p[0], p[1], hiA, loA);
      ag_split(p[2], p[3], hiB, loB);
      const uint32_t sb = ring0 + stage * kAgStageBytes;
      {
        // k-step 0: slots of destinations 0 (k 0-7) and 1 (k 8-15) feed fragment rows 0-7 only
        const uint32_t a0h = hiA & mask_d0, a0l = loA & mask_d0, a2h = hiA & mask_d1, a2l = loA & mask_d1;
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          ag_ldmatrix_x4_trans(sb + ld_off[np], b);
          ag_mma_bf16(acc[2 * np], a0h, 0u, a2h, 0u, b[0], b[1]);
          ag_mma_bf16(acc[2 * np], a0l, 0u, a2l, 0u, b[0], b[1]);
          ag_mma_bf16(acc[2 * np + 1], a0h, 0u, a2h, 0u, b[2], b[3]);
          ag_mma_bf16(acc[2 * np + 1], a0l, 0u, a2l, 0u, b[2], b[3]);
        }
        // k-step 1: destinations 2, 3 feed fragment rows 8-15
        const uint32_t a1h = hiB & mask_d0, a1l = loB & mask_d0, a3h = hiB & mask_d1, a3l = loB & mask_d1;
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          ag_ldmatrix_x4_trans(sb + 16 * kAgRowBytes + ld_off[np], b);
          ag_mma_bf16(acc[2 * np], 0u, a1h, 0u, a3h, b[0], b[1]);
          ag_mma_bf16(acc[2 * np], 0u, a1l, 0u, a3l, b[0], b[1]);
          ag_mma_bf16(acc[2 * np + 1], 0u, a1h, 0u, a3h, b[2], b[3]);
          ag_mma_bf16(acc[2 * np + 1], 0u, a1l, 0u, a3l, b[2], b[3]);
        }
      }
      __syncwarp();                                           // every lane has read the stage: the next issue may refill it
      // F. last chunk of the step: softmax denominators, z rows -> A operand, hand the tile to the tensor core
      if (last_0) {
        float d0 = den0, d1 = den1;
        d0 += __shfl_xor_sync(kFull, d0, 1); d1 += __shfl_xor_sync(kFull, d1, 1);
        d0 += __shfl_xor_sync(kFull, d0, 2); d1 += __shfl_xor_sync(kFull, d1, 2);
        const float inv0 = ag_rcp(d0 + 1e-10f), inv1 = ag_rcp(d1 + 1e-10f);           // graph_attention.py:96
        if (it_0 > 0 && row0_0 < 4 * kAgWarps) tc_mbar_wait(bar_a_empty, (uint32_t)((it_0 - 1) & 1));   // (the warp's first step of the tile)
        // rows qa = row0 + dsel and qa + 2; chunk nt of a row sits at ((nt ^ (row & 7)) << 4)
        const uint32_t za = z_lane + (uint32_t)row0_0 * 128u;
        const uint32_t xa = (uint32_t)(((row0_0 + dsel) & 7) << 4), xb = xa ^ 0x20u;   // (row + 2) & 7 = (row & 7) ^ 2 (row & 3 = dsel < 2)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const uint32_t va = ag_pack_bf16(acc[nt][0] * inv0, acc[nt][1] * inv0);
          const uint32_t vb = ag_pack_bf16(acc[nt][2] * inv1, acc[nt][3] * inv1);
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(za + ((uint32_t)(nt << 4) ^ xa)), "r"(va) : "memory");
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(za + 256u + ((uint32_t)(nt << 4) ^ xb)), "r"(vb) : "memory");
        }
        // the warp's last step of the tile (steps warp, warp + 12, ...): one fence and one arrival for all its z rows
        if (row0_0 + 4 * kAgWarps >= kAgTile) {
          tc_fence_async_smem();                              // generic-proxy writes -> visible to the tensor core
          __syncwarp();
          if (lane == 0) {
            __threadfence_block();
            int prev;                                         // (shared-space atomic: atomicAdd on the generic pointer compiles to ATOM.E)
            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(prev) : "r"(tc_smem_u32(arrive_cnt)) : "memory");
            const bool last = prev == kAgWarps * (it_0 + 1) - 1;
            __threadfence_block();
            if (last) issue_mma(it_0);                        // the LAST warp to finish the tile issues its MMAs
          }
          __syncwarp();
        }
      }
      // rotate the pipeline
      node0_0 = node0_1; row0_0 = row0_1; it_0 = it_1; first_0 = first_1; last_0 = last_1; src0 = src1;
#pragma unroll
      for (int i = 0; i < 4; ++i) sv0[i] = sv1[i];
      st0[0] = st1[0]; st0[1] = st1[1]; M0[0] = M1[0]; M0[1] = M1[1];
      node0_1 = node0_2; row0_1 = row0_2; it_1 = it_2; first_1 = first_2; last_1 = last_2; src1 = src2;
      stage ^= 1u;
    }
    ag_cp_wait<0>();
#undef AG_NEXT_STEP
#undef AG_FETCH_RP
#undef AG_ADVANCE
#undef AG_LOAD_SRC
#undef AG_ISSUE_ROWS
#undef AG_LOAD_ATT
#undef AG_LOAD_TGT
  } else {
    // ============ epilogue warps: TMEM -> ELU -> head mean / concat -> global (as in gat_tc_kernel) ============
    const int ew = warp & 3;                                  // TMEM lane quarter this warp may access
    const int row = ew * 32 + lane;
    const int et = (warp - kAgWarps) * 32 + lane;             // 0..127 linear id among epilogue threads
    const float inv_h = 1.f / (float)NH;
    const int out_w = A.concat ? NH * F : F;
    const int cpr = F / 8;                                    // 16-byte chunks per staged bf16 row
    const int swz_mask = ((cpr & (cpr - 1)) == 0) ? (min(cpr, 8) - 1) : 0;
    const int cpr_sh = ((cpr & (cpr - 1)) == 0) ? __ffs(cpr) - 1 : -1;       // power of two: row = idx >> cpr_sh (no runtime division)
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int stg = it & 1, nn = it >> 1;
      const int tile_base = tile * kAgTile;
      const int node = tile_base + row;
      tc_mbar_wait_relaxed(bar_t_full0 + 8 * stg, (uint32_t)(nn & 1));
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(stg * NH * F);
      for (int c0 = 0; c0 < F; c0 += 16) {
        float oacc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) oacc[i] = 0.f;
        unsigned long long oacc2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) oacc2[i] = 0ull;
#pragma unroll
        for (int h0 = 0; h0 < NH; h0 += 2) {
          uint32_t v[2][16];
          tc_tmem_ld16(t_row + (uint32_t)(h0 * F + c0), v[0]);
          tc_tmem_ld16(t_row + (uint32_t)((h0 + 1) * F + c0), v[1]);
          tc_tmem_wait_ld();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (!A.concat) {                                   // ELU per head (:118) and the head sum (:158) on packed pairs
#pragma unroll
              for (int i = 0; i < 8; ++i) ag_elu2_acc(oacc2[i], v[hh][2 * i], v[hh][2 * i + 1]);
              continue;
            }
            float e16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) e16[i] = tc_elu(__uint_as_float(v[hh][i]));          // ELU per head (:118)
            if (A.concat) {
              if (node < A.N) {
                const size_t o = (size_t)node * out_w + (size_t)(h0 + hh) * F + c0;
                if (A.out_bf16) {
                  uint4 pk[2];
                  unsigned* pw = reinterpret_cast<unsigned*>(pk);
#pragma unroll
                  for (int i = 0; i < 8; ++i) pw[i] = ag_pack_bf16(e16[2 * i], e16[2 * i + 1]);
                  uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(A.out) + o);
                  op[0] = pk[0]; op[1] = pk[1];
                } else {
                  float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(A.out) + o);
#pragma unroll
                  for (int i = 0; i < 4; ++i) op[i] = make_float4(e16[4 * i], e16[4 * i + 1], e16[4 * i + 2], e16[4 * i + 3]);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) oacc[i] += e16[i];                                  // then the head mean (:158)
            }
          }
        }
        if (!A.concat) {
#pragma unroll
          for (int i = 0; i < 8; ++i) tc_unpack2(oacc2[i], oacc[2 * i], oacc[2 * i + 1]);
          if (staged) {
            uint4 pk[2];
            unsigned* pw = reinterpret_cast<unsigned*>(pk);
#pragma unroll
            for (int i = 0; i < 8; ++i) pw[i] = ag_pack_bf16(oacc[2 * i] * inv_h, oacc[2 * i + 1] * inv_h);
            const int chunk = c0 >> 3;
            *reinterpret_cast<uint4*>(Ss + (size_t)row * F * 2 + ((chunk ^ (row & swz_mask)) << 4)) = pk[0];
            *reinterpret_cast<uint4*>(Ss + (size_t)row * F * 2 + (((chunk + 1) ^ (row & swz_mask)) << 4)) = pk[1];
          } else if (node < A.N) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(A.out) + (size_t)node * out_w + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              op[i] = make_float4(oacc[4 * i] * inv_h, oacc[4 * i + 1] * inv_h, oacc[4 * i + 2] * inv_h, oacc[4 * i + 3] * inv_h);
          }
        }
      }
      // TMEM stage drained: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(bar_t_empty0 + 8 * stg);
      if (staged) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int valid_rows = min(kAgTile, A.N - tile_base);
        const int nchunks = valid_rows * cpr;
        uint4* gdst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(A.out) + (size_t)tile_base * F);
        for (int idx = et; idx < nchunks; idx += 128) {
          const int r = cpr_sh >= 0 ? idx >> cpr_sh : idx / cpr, c = idx - r * cpr;
          gdst[idx] = *reinterpret_cast<const uint4*>(Ss + (size_t)r * F * 2 + ((c ^ (r & swz_mask)) << 4));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kAgWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(A.tmem_cols) : "memory");
  }
}

static bool ag_mma_supported(int in_dim, int F, int heads, int concat, int out_bf16) {
  static const int enabled = getenv("MG_GAT_AGG_MMA") ? atoi(getenv("MG_GAT_AGG_MMA")) : 1;
  if (!enabled || heads != kAgNH || in_dim != kAgIn) return false;
  if (F % 16 != 0 || F < 16 || 2 * heads * F > 512) return false;
  return ag_smem_layout(F, out_bf16 && !concat).total <= 227 * 1024;
}

// launch with programmatic stream serialisation: the kernel may start while its predecessor on the stream is still
// running; it calls griddepcontrol.wait before it reads the predecessor's outputs
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static const int pdl_env = getenv("MG_GAT_PDL") ? atoi(getenv("MG_GAT_PDL")) : 1;      // 0: plain stream order (debugging)
  attr[0].val.programmaticStreamSerializationAllowed = pdl_env ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// =====================================================================================================================
// The same tensor-core aggregation for WIDE rows (heads 4, in = 64 * nsl, bf16 storage) in front of the TMA-fed transform
// (gat_tma_gemm.cu): z does not fit an SM's shared memory as a UMMA operand, so it is spilled to HBM as bf16 (N, heads, in)
// — but by mma.sync warps instead of the FP32-pipe gat_aggregate_kernel.  A source row is walked in 128-byte slabs of 64
// features: sub-item (step of 4 destinations, slab, chunk of 8 in-edges) runs the item pipeline of gat_agg_mma_kernel
// (cp.async-staged rows, ldmatrix.trans, block-diagonal attention fragments).  With one chunk per step (in-degree <= 8)
// the column ids, the attention scalars and the numerator fragments of a step are computed once and reused for every
// slab.  No tiles, no TMEM: 16 aggregation warps per SM, steps dealt round-robin over all warps of the grid.
// =====================================================================================================================
constexpr int kSpWarps = 16;
constexpr int kSpThreads = kSpWarps * 32;

struct GatSpillArgs {
  const __nv_bfloat16* x;
  const int32_t* rowptr;
  const int32_t* col;
  const float* s;          // (N, 8)
  const float* gmax;       // (G, 4)
  __nv_bfloat16* z;        // (N, 4, in)
  int N, in_dim, nodes_per_graph;
  float slope;
};

__host__ __device__ inline int sp_smem_bytes() { return kSpWarps * (kAgStages * kAgStageBytes + kAgAttBytes) + 1024; }

__device__ __forceinline__ const char* sp_row_ptr(const char* lane_base, int row, unsigned row_bytes) {
  unsigned long long p;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"((unsigned)row), "r"(row_bytes), "l"((unsigned long long)lane_base));
  return reinterpret_cast<const char*>(p);
}

__global__ void __launch_bounds__(kSpThreads, 1) gat_agg_spill_kernel(const GatSpillArgs A) {
  constexpr int NH = kAgNH;
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g8 = lane >> 2, t4 = lane & 3;                    // MMA fragment coordinates
  const int hd = g8 & 3, dsel = g8 >> 2;                      // fragment row g8 = (destination dsel, head hd); row g8 + 8 = (dsel + 2, hd)
  const int r4 = lane >> 3, ch8 = lane & 7;                   // loader role: destination of the step, 16-byte chunk / edge slot
  const uint32_t ring0 = tc_smem_u32(sm + (size_t)warp * kAgStages * kAgStageBytes);
  const uint32_t att0 = tc_smem_u32(sm + (size_t)kSpWarps * kAgStages * kAgStageBytes + (size_t)warp * kAgAttBytes);
  const uint32_t att_w = att0 + (uint32_t)(r4 * kAgAttDst + ch8 * 16);
  const uint32_t att_r = att0 + (uint32_t)(dsel * kAgAttDst + t4 * 32 + hd * 4);
  constexpr float kLog2e = 1.4426950408889634f;
  const float slope = A.slope;
  const int npg = A.nodes_per_graph, N = A.N, in_dim = A.in_dim, nsl = in_dim >> 6;
  const unsigned row_bytes = (unsigned)in_dim * 2u;
  const int gw = (int)blockIdx.x * kSpWarps + warp, nw = (int)gridDim.x * kSpWarps;
  const int nsteps = (N + 3) >> 2;
  const uint32_t cp_off_even = (uint32_t)(r4 * kAgRowBytes + ((ch8 ^ r4) << 4));
  const uint32_t cp_off_odd = (uint32_t)(r4 * kAgRowBytes + ((ch8 ^ (r4 | 4)) << 4));
  const char* x_lane = reinterpret_cast<const char*>(A.x) + ch8 * 16;
  uint32_t ld_off[4];
#pragma unroll
  for (int np = 0; np < 4; ++np)
    ld_off[np] = (uint32_t)((((r4 & 1) * 8 + ch8) * kAgRowBytes) + ((((np << 1) | (r4 >> 1)) ^ ch8) << 4));
  const uint32_t mask_d0 = dsel == 0 ? 0xffffffffu : 0u, mask_d1 = ~mask_d0;

  pdl_launch_dependents();
  const float* s_in = A.s;
  const float* gmax_in = A.gmax;
  pdl_wait_prior_grid();                                      // s, gmax of the pre-pass: read with ld_pre from here on
  const float* s_tgt_lane = s_in + NH + hd;
  const float M_single = npg > 0 ? 0.f : leaky_relu(ld_pre(gmax_in + hd), slope);

  // ---- cursor over (step, slab, chunk); row pointers of the step after the next one prefetched in place ----
  int cur_step = gw, cur_sl = 0, cur_c = 0, cur_nch = 1, cur_beg = 0, cur_end = 0, nxt_beg = 0, nxt_end = 0;
  auto fetch_rp = [&](int step, int& b, int& e, bool doit) {
    const int j = step < nsteps ? min(step * 4 + r4, N) : N;  // past the last node: rowptr[N] twice, an empty range
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p ld.global.nc.b32 %0, [%3];\n@p ld.global.nc.b32 %1, [%4];\n}"
                 : "+r"(b), "+r"(e)
                 : "r"(doit ? 1 : 0), "l"(A.rowptr + j), "l"(A.rowptr + min(j + 1, N)));
  };
  fetch_rp(cur_step, cur_beg, cur_end, true);
  fetch_rp(cur_step + nw, nxt_beg, nxt_end, true);
  cur_nch = max(1, __reduce_max_sync(kFull, (cur_end - cur_beg + 7) >> 3));
  auto advance = [&]() {
    bool adv_step = false;
    if (++cur_c >= cur_nch) {
      cur_c = 0;
      if (++cur_sl >= nsl) {
        cur_sl = 0;
        cur_step += nw;
        adv_step = true;
      }
    }
    cur_beg = adv_step ? nxt_beg : cur_beg;
    cur_end = adv_step ? nxt_end : cur_end;
    fetch_rp(cur_step + nw, nxt_beg, nxt_end, adv_step);
    if (adv_step) cur_nch = max(1, __reduce_max_sync(kFull, (cur_end - cur_beg + 7) >> 3));
  };
  auto load_src = [&]() -> int {                              // loader lane: source node of its slot (-1: no edge)
    const int k = cur_beg + 8 * cur_c + ch8;
    return (cur_step < nsteps && k < cur_end) ? __ldg(A.col + k) : -1;
  };
  auto issue_rows = [&](int src, int sl, uint32_t stg) {      // 32 slab rows of 128 B into ring stage stg (zero-filled where empty)
    const uint32_t sb = ring0 + stg * kAgStageBytes;
    const char* xs = x_lane + sl * 128;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int sn = __shfl_sync(kFull, src, i * 4 + r4);
      ag_cp_async16(sb + (uint32_t)(i * 4 * kAgRowBytes) + ((i & 1) ? cp_off_odd : cp_off_even), sp_row_ptr(xs, max(sn, 0), row_bytes),
                    sn >= 0 ? 16u : 0u);
    }
  };
  auto load_att = [&](int src, int node0, float (&sv)[4], float (&st2)[2], float (&M2)[2]) {
    sv[0] = sv[1] = sv[2] = sv[3] = -INFINITY;
    if (src >= 0) {
      const float4 v = ld_pre(reinterpret_cast<const float4*>(s_in + (size_t)(unsigned)src * (2 * NH)));
      sv[0] = v.x; sv[1] = v.y; sv[2] = v.z; sv[3] = v.w;
    }
    const int ja = node0 + dsel, jb = ja + 2;
    const bool va = node0 >= 0 && ja < N, vb = node0 >= 0 && jb < N;
    st2[0] = va ? ld_pre(s_tgt_lane + (size_t)(unsigned)ja * (2 * NH)) : 0.f;
    st2[1] = vb ? ld_pre(s_tgt_lane + (size_t)(unsigned)jb * (2 * NH)) : 0.f;
    M2[0] = M_single; M2[1] = M_single;
    if (npg > 0) {
      M2[0] = va ? leaky_relu(ld_pre(gmax_in + (size_t)(ja / npg) * NH + hd), slope) : 0.f;
      M2[1] = vb ? leaky_relu(ld_pre(gmax_in + (size_t)(jb / npg) * NH + hd), slope) : 0.f;
    }
  };

  // ---- software pipeline: sub-item n on the tensor cores, rows + scalars of n+1 and column ids of n+2 in flight ----
  int node0_0 = cur_step < nsteps ? cur_step * 4 : -1, sl_0 = 0;
  bool first_0 = true, last_0 = cur_nch == 1, reuse_0 = false;
  int src0 = load_src();
  float sv0[4], st0[2], M0[2];
  load_att(src0, node0_0, sv0, st0, M0);
  if (node0_0 >= 0) issue_rows(src0, sl_0, 0u);
  ag_cp_commit();
  advance();
  int node0_1 = cur_step < nsteps ? cur_step * 4 : -1, sl_1 = cur_sl;
  bool first_1 = cur_c == 0, last_1 = cur_c == cur_nch - 1, reuse_1 = cur_nch == 1 && cur_sl > 0;
  int src1 = reuse_1 ? src0 : load_src();
  float acc[8][4];
  float den0 = 0.f, den1 = 0.f, psum0 = 0.f, psum1 = 0.f;
  uint32_t hiA = 0, loA = 0, hiB = 0, loB = 0;                // numerator fragments of the step's chunk (kept across slabs)
  uint32_t stage = 0;
  while (node0_0 >= 0) {
    // A. rows of sub-item n+1
    if (node0_1 >= 0) issue_rows(src1, sl_1, stage ^ 1u);
    ag_cp_commit();
    // B. column ids of sub-item n+2
    advance();
    const int node0_2 = cur_step < nsteps ? cur_step * 4 : -1, sl_2 = cur_sl;
    const bool first_2 = cur_c == 0, last_2 = cur_c == cur_nch - 1, reuse_2 = cur_nch == 1 && cur_sl > 0;
    const int src2 = reuse_2 ? src1 : load_src();
    // C. attention scalars of sub-item n+1 (the same as sub-item n's when it is another slab of the same chunk)
    float sv1[4], st1[2], M1[2];
    if (reuse_1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sv1[i] = sv0[i];
      st1[0] = st0[0]; st1[1] = st0[1]; M1[0] = M0[0]; M1[1] = M0[1];
    } else {
      load_att(src1, node0_1, sv1, st1, M1);
    }
    // D. rows of sub-item n have landed
    ag_cp_wait<1>();
    __syncwarp();
    // single-chunk steps (every in-degree <= 8): numerators normalised once, accumulators start from zero (see gat_agg_mma_kernel)
    const bool single = first_0 && last_0;
    if (first_0 && !single) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      den0 = den1 = 0.f;
    }
    // E. attention numerators (graph_attention.py:61-65,86) and the block-diagonal A fragments
    if (!reuse_0) {
      float sq[4];
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(att_w), "f"(sv0[0]), "f"(sv0[1]), "f"(sv0[2]), "f"(sv0[3]) : "memory");
      __syncwarp();
      asm volatile("ld.shared.f32 %0, [%4];\nld.shared.f32 %1, [%4+16];\nld.shared.f32 %2, [%4+288];\nld.shared.f32 %3, [%4+304];"
                   : "=f"(sq[0]), "=f"(sq[1]), "=f"(sq[2]), "=f"(sq[3]) : "r"(att_r) : "memory");
      float p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float e = sq[i] + st0[i >> 1];
        p[i] = tc_ex2((fmaxf(e, e * slope) - M0[i >> 1]) * kLog2e);                   // LeakyReLU (0 <= slope <= 1); empty slot: ex2(-inf) = 0
      }
      psum0 = p[0] + p[1];
      psum1 = p[2] + p[3];
      if (single) {                                           // (a reused fragment belongs to a single-chunk step too)
        float d0 = psum0, d1 = psum1;
        ag_quad_sum(d0, d1);
        const float i0 = ag_rcp(d0 + 1e-10f), i1 = ag_rcp(d1 + 1e-10f);              // graph_attention.py:96
        p[0] *= i0; p[1] *= i0; p[2] *= i1; p[3] *= i1;
      }
      ag_split(p[0], p[1], hiA, loA);
      ag_split(p[2], p[3], hiB, loB);
    }
    den0 += psum0;
    den1 += psum1;
    const uint32_t sb = ring0 + stage * kAgStageBytes;
    if (single) ag_item_mma<true>(acc, sb, ld_off, hiA, loA, hiB, loB, mask_d0, mask_d1);
    else ag_item_mma<false>(acc, sb, ld_off, hiA, loA, hiB, loB, mask_d0, mask_d1);
    __syncwarp();                                             // every lane has read the stage: the next issue may refill it
    // F. last chunk of the (step, slab): softmax denominators, the slab's 64 columns of z -> HBM (bf16)
    if (last_0) {
      float inv0 = 1.f, inv1 = 1.f;
      if (!single) {
        float d0 = den0, d1 = den1;
        ag_quad_sum(d0, d1);
        inv0 = ag_rcp(d0 + 1e-10f); inv1 = ag_rcp(d1 + 1e-10f);                       // graph_attention.py:96
      }
      const int na = node0_0 + dsel, nb = na + 2;
      char* za = reinterpret_cast<char*>(A.z) + (((size_t)na * NH + hd) * in_dim + sl_0 * 64 + 2 * t4) * 2;
      char* zb = za + (size_t)2 * NH * in_dim * 2;
      if (single) {
        if (na < N) {
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<uint32_t*>(za + nt * 16) = ag_pack_bf16(acc[nt][0], acc[nt][1]);
        }
        if (nb < N) {
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<uint32_t*>(zb + nt * 16) = ag_pack_bf16(acc[nt][2], acc[nt][3]);
        }
      } else {
        if (na < N) {
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
            *reinterpret_cast<uint32_t*>(za + nt * 16) = ag_pack_bf16(acc[nt][0] * inv0, acc[nt][1] * inv0);
        }
        if (nb < N) {
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
            *reinterpret_cast<uint32_t*>(zb + nt * 16) = ag_pack_bf16(acc[nt][2] * inv1, acc[nt][3] * inv1);
        }
      }
    }
    // rotate the pipeline
    node0_0 = node0_1; sl_0 = sl_1; first_0 = first_1; last_0 = last_1; reuse_0 = reuse_1; src0 = src1;
#pragma unroll
    for (int i = 0; i < 4; ++i) sv0[i] = sv1[i];
    st0[0] = st1[0]; st0[1] = st1[1]; M0[0] = M1[0]; M0[1] = M1[1];
    node0_1 = node0_2; sl_1 = sl_2; first_1 = first_2; last_1 = last_2; reuse_1 = reuse_2; src1 = src2;
    stage ^= 1u;
  }
  ag_cp_wait<0>();
}

bool gat_agg_spill_supported(int N, int in_dim, int heads) {
  static const int enabled = getenv("MG_GAT_AGG_SPILL_MMA") ? atoi(getenv("MG_GAT_AGG_SPILL_MMA")) : 1;
  return enabled && heads == kAgNH && in_dim >= 64 && in_dim % 64 == 0 && in_dim <= 1024 && N >= 4096;
}

// z (N, 4, in) bf16 <- attention-weighted neighbour sums; s, gmax from the pre-pass launched before on the same stream
int gat_agg_spill_launch(const void* x, const int32_t* rowptr, const int32_t* col, const float* s, const float* gmax, void* z_bf16, int N,
                         int in_dim, float slope, int nodes_per_graph, cudaStream_t st) {
  GatSpillArgs A;
  A.x = reinterpret_cast<const __nv_bfloat16*>(x);
  A.rowptr = rowptr; A.col = col; A.s = s; A.gmax = gmax; A.z = reinterpret_cast<__nv_bfloat16*>(z_bf16);
  A.N = N; A.in_dim = in_dim; A.nodes_per_graph = nodes_per_graph; A.slope = slope;
  if (cudaFuncSetAttribute(gat_agg_spill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
    set_error("gat_agg_spill_kernel: cannot raise dynamic shared memory");
    return MG_ERR_CUDA;
  }
  const int grid = std::min(ceil_div(ceil_div(N, 4), kSpWarps), num_sms());
  launch_pdl(gat_agg_spill_kernel, dim3(grid), dim3(kSpThreads), (size_t)sp_smem_bytes(), st, A);
  return check_launch("gat_agg_spill_kernel");
}


// The score pre-pass of a layer on bf16 node features (also used in front of the spilled-z path of gat_forward.cu):
// u = W^T a (one block per head), s = X U^T on mma.sync, the exact per-graph edge maximum.  Chained with programmatic
// dependent launch; the caller's next kernel follows in stream order.
template <int NH, int LPN>
static int launch_prepass(const __nv_bfloat16* x, const int32_t* rowptr, const int32_t* col, int N, int64_t E, const float* W,
                          const float* a, int F, int nodes_per_graph, float* s, float* gmax, float* u, cudaStream_t st) {
  const int G = nodes_per_graph > 0 ? N / nodes_per_graph : 1;
  int rc;
  const int sgrid = (int)std::min<int64_t>(ceil_div64((int64_t)N * 2, 256), (int64_t)num_sms() * 8);   // 16 nodes per warp step
  // single graph with enough edges per node: prune the edge-max scan (gsrc sits in the slack of the 256-byte gmax segment)
  static const int prune_env = getenv("MG_GAT_PRUNE") ? atoi(getenv("MG_GAT_PRUNE")) : 1;
  static const int prune_deg = getenv("MG_GAT_PRUNE_MIN_DEG") ? atoi(getenv("MG_GAT_PRUNE_MIN_DEG")) : 12;
  float* gsrc = (prune_env && G == 1 && NH <= 4 && E >= (int64_t)N * prune_deg) ? gmax + 16 : nullptr;
  tc_u_kernel<NH, LPN><<<NH, 256, 0, st>>>(W, a, F, G, u, gmax, gsrc);
  if ((rc = check_launch("tc_u_kernel"))) return rc;
  launch_pdl(tc_scores_kernel<NH, LPN>, dim3(sgrid), dim3(256), 0, st, x, N, (const float*)u, s, gsrc);
  if ((rc = check_launch("tc_scores_kernel"))) return rc;
  const int mgrid = std::min(ceil_div(N, 256), num_sms() * 6);
  if (gsrc) {
    launch_pdl(tc_edge_first_kernel<NH>, dim3(mgrid), dim3(256), 0, st, rowptr, col, (const float*)s, N, gmax);
    if ((rc = check_launch("tc_edge_first_kernel"))) return rc;
    if (E >= (int64_t)N * 12)          // many in-edges: survivors compacted per block and scanned by whole warps
      launch_pdl(tc_edge_max_pruned_kernel<NH>, dim3(mgrid), dim3(256), 0, st, rowptr, col, (const float*)s, N, gmax,
                 (const float*)gsrc);
    else                               // few in-edges (only with MG_GAT_PRUNE_MIN_DEG < 12): a surviving thread scans its own edges.  Measured at
                                       // k = 8: 0.136 ms against 0.118 ms unpruned — two short kernels cost more than one 18 us scan
      launch_pdl(tc_edge_max_kernel<NH, 8>, dim3(mgrid), dim3(256), 0, st, rowptr, col, (const float*)s, N, nodes_per_graph, gmax,
                 (const float*)gsrc);
    if ((rc = check_launch("tc_edge_max_pruned_kernel"))) return rc;
  } else if (E > (int64_t)N * 12)      // high in-degree: more gathers in flight per destination (at k = 8 both widths take 18-19 us: the pass is bound by ~0.5 L1-miss sectors per clock and SM, not by its dependent round trips)
    launch_pdl(tc_edge_max_kernel<NH, 8>, dim3(mgrid), dim3(256), 0, st, rowptr, col, (const float*)s, N, nodes_per_graph, gmax,
               (const float*)gsrc);
  else
    launch_pdl(tc_edge_max_kernel<NH, 4>, dim3(mgrid), dim3(256), 0, st, rowptr, col, (const float*)s, N, nodes_per_graph, gmax,
               (const float*)gsrc);
  return check_launch("tc_edge_max_kernel");
}

template <int NH, int LPN>
static int launch_tc(const GatTcArgs& A, const float* a, float* s, float* gmax, float* u, size_t smem, int grid,
                     cudaStream_t st) {
  auto k = gat_tc_kernel<NH, LPN>;
  // per launch (cheap, and correct when one process drives several devices: function attributes are per device)
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
    set_error("gat_tc_kernel: cannot raise dynamic shared memory");
    return MG_ERR_CUDA;
  }
  int rc;
  if ((rc = launch_prepass<NH, LPN>(A.x, A.rowptr, A.col, A.N, A.E, A.W, a, A.F, A.nodes_per_graph, s, gmax, u, st))) return rc;
  if (NH == kAgNH && LPN * 8 == kAgIn && ag_mma_supported(LPN * 8, A.F, NH, A.concat, A.out_bf16)) {
    if (cudaFuncSetAttribute(gat_agg_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("gat_agg_mma_kernel: cannot raise dynamic shared memory");
      return MG_ERR_CUDA;
    }
    const int ag_grid = std::min(ceil_div(A.N, kAgTile), num_sms());
    launch_pdl(gat_agg_mma_kernel, dim3(ag_grid), dim3(kAgThreads), (size_t)ag_smem_layout(A.F, A.out_bf16 && !A.concat).total, st, A);
    return check_launch("gat_agg_mma_kernel");
  }
  launch_pdl(k, dim3(grid), dim3(kTcThreads), smem, st, A);
  return check_launch("gat_tc_kernel");
}

// The exact per-graph edge maximum alone (it only reads s, rowptr and col: any feature dtype), behind a score kernel that has
// reset gmax to -inf: used by the FP32-pipe pre-pass of gat_forward.cu, whose own thread-per-destination kernel took 50 us
// (k = 8) / 230 us (k = 32) at N = 262 144 against 18 / 70 us here.
bool gat_tc_edge_max_supported(int heads) {
  static const int enabled = getenv("MG_GAT_TC_EDGE_MAX") ? atoi(getenv("MG_GAT_TC_EDGE_MAX")) : 1;
  return enabled && (heads == 1 || heads == 2 || heads == 4);
}
template <int NH>
static int launch_edge_max_only(const int32_t* rowptr, const int32_t* col, const float* s, int N, int64_t E, int nodes_per_graph, float* gmax,
                                cudaStream_t st) {
  const int mgrid = std::min(ceil_div(N, 256), num_sms() * 6);
  if (E > (int64_t)N * 12)
    launch_pdl(tc_edge_max_kernel<NH, 8>, dim3(mgrid), dim3(256), 0, st, rowptr, col, s, N, nodes_per_graph, gmax, (const float*)nullptr);
  else
    launch_pdl(tc_edge_max_kernel<NH, 4>, dim3(mgrid), dim3(256), 0, st, rowptr, col, s, N, nodes_per_graph, gmax, (const float*)nullptr);
  return check_launch("tc_edge_max_kernel");
}
int gat_tc_edge_max(const int32_t* rowptr, const int32_t* col, const float* s, int N, int64_t E, int heads, int nodes_per_graph, float* gmax,
                    cudaStream_t st) {
  if (heads == 4) return launch_edge_max_only<4>(rowptr, col, s, N, E, nodes_per_graph, gmax, st);
  if (heads == 2) return launch_edge_max_only<2>(rowptr, col, s, N, E, nodes_per_graph, gmax, st);
  if (heads == 1) return launch_edge_max_only<1>(rowptr, col, s, N, E, nodes_per_graph, gmax, st);
  set_error("gat_tc_edge_max: heads=%d", heads);
  return MG_ERR_UNSUPPORTED;
}

// bf16 score pre-pass for layers that do not run on gat_tc_kernel / gat_agg_mma_kernel (spilled z, FP32-pipe fused kernel)
bool gat_tc_prepass_supported(int N, int in_dim, int heads) {
  static const int enabled = getenv("MG_GAT_TC_PREPASS") ? atoi(getenv("MG_GAT_TC_PREPASS")) : 1;
  if (!enabled || N < 4096) return false;
  return (heads == 1 || heads == 2 || heads == 4) && (in_dim == 32 || in_dim == 64 || in_dim == 128 || in_dim == 256);
}
int gat_tc_prepass(const void* x, const int32_t* rowptr, const int32_t* col, int N, int64_t E, const float* W, const float* a, int in_dim,
                   int F, int heads, int nodes_per_graph, float* s, float* gmax, float* u, cudaStream_t st) {
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  const int lpn = in_dim / 8;
#define MG_PRE(NHH, LL) \
  if (heads == NHH && lpn == LL) return launch_prepass<NHH, LL>(xb, rowptr, col, N, E, W, a, F, nodes_per_graph, s, gmax, u, st);
  MG_PRE(4, 4) MG_PRE(4, 8) MG_PRE(4, 16) MG_PRE(4, 32) MG_PRE(2, 4) MG_PRE(2, 8) MG_PRE(2, 16) MG_PRE(2, 32)
  MG_PRE(1, 4) MG_PRE(1, 8) MG_PRE(1, 16) MG_PRE(1, 32)
#undef MG_PRE
  set_error("gat_tc_prepass: no variant for heads=%d in=%d", heads, in_dim);
  return MG_ERR_UNSUPPORTED;
}

// Shapes the tensor-pipe kernel takes (everything else stays on the FP32-pipe kernels).
bool gat_tc_supported(int N, int in_dim, int F, int heads, int concat, int out_bf16) {
  static const int enabled = getenv("MG_GAT_TC") ? atoi(getenv("MG_GAT_TC")) : 1;
  static const int min_nodes = getenv("MG_GAT_TC_MIN_NODES") ? atoi(getenv("MG_GAT_TC_MIN_NODES")) : 4096;
  if (!enabled || N < min_nodes) return false;
  if (!(heads == 1 || heads == 2 || heads == 4)) return false;
  if (!(in_dim == 32 || in_dim == 64 || in_dim == 128 || in_dim == 256) || heads * in_dim > 256) return false;
  if (F % 16 != 0 || F < 16 || F > 256 || 2 * heads * F > 512) return false;
  const TcSmem L = tc_smem_layout(heads, in_dim, F, out_bf16 && !concat);
  return L.total <= 227 * 1024;
}

int gat_tc_launch(const void* x, const int32_t* rowptr, const int32_t* col, float* s, float* gmax, float* u, const float* W,
                  const float* a, int N, int64_t E, int in_dim, int F, int heads, int concat, float slope, int nodes_per_graph, void* out,
                  int out_bf16, cudaStream_t st) {
  GatTcArgs A;
  A.x = reinterpret_cast<const __nv_bfloat16*>(x);
  A.rowptr = rowptr; A.col = col; A.s = s; A.gmax = gmax; A.W = W; A.out = out;
  A.N = N; A.E = E; A.F = F; A.concat = concat; A.out_bf16 = out_bf16; A.nodes_per_graph = nodes_per_graph; A.slope = slope;
  int cols = 32;
  while (cols < 2 * heads * F) cols <<= 1;
  A.tmem_cols = cols;
  const TcSmem L = tc_smem_layout(heads, in_dim, F, out_bf16 && !concat);
  const int grid = std::min(ceil_div(N, kTcTile), num_sms());
  const int lpn = in_dim / 8;
#define MG_TC(NHH, LL) \
  if (heads == NHH && lpn == LL) return launch_tc<NHH, LL>(A, a, s, gmax, u, (size_t)L.total, grid, st);
  MG_TC(4, 8) MG_TC(4, 4) MG_TC(2, 16) MG_TC(2, 8) MG_TC(2, 4) MG_TC(1, 32) MG_TC(1, 16) MG_TC(1, 8) MG_TC(1, 4)
#undef MG_TC
  set_error("gat_tc: no variant for heads=%d in=%d", heads, in_dim);
  return MG_ERR_UNSUPPORTED;
}

}  // namespace mg
