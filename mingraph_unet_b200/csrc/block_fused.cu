// Fused per-image graph block (forward): ONE launch runs, for every image of the batch,
//   patch GAT -> predictor GAT -> softmax/argmax -> N-cut loss -> region mean-pool -> region GAT
// i.e. scripts/train_end_to_end.py:329-389 of the reference (model/gat/graph_attention.py:40-160,
// model/graph_partition/mincut_refinement.py:43-160,192-193) on the 4-connected patch grid.
//
// B200 mapping: each image is owned by one thread-block CLUSTER (<= 8 CTAs, one per SM).  The
// image's nodes are split row-major over the cluster's CTAs; the stages that need a per-image
// reduction (the global softmax shift of graph_attention.py:86, the N-cut sums, the region
// means) exchange a few floats through distributed shared memory and a hardware cluster
// barrier instead of a kernel boundary.  Neighbours are closed-form (up, left, right, down =
// ascending COO edge id), so no edge list is read at all.  Node-level intermediates that
// neighbouring CTAs need (h, predictor scalars, S) go through global memory (L2 resident) and
// become visible at the cluster barriers (release/acquire at cluster scope).
//
// Weights are pre-arranged once per weight version by mg_block_prepare (transposed W, attention
// vectors u = W^T a) and staged into shared memory with one TMA bulk copy (cp.async.bulk).
#include <cooperative_groups.h>

#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mg {

constexpr int kBT = 512;          // threads per CTA
constexpr int kTileN = 128;       // nodes per aggregation/transform tile
constexpr int kMaxHeads = 4;
constexpr int kMaxSeg = 8;        // K
constexpr int kMaxCluster = 8;
constexpr int kSmemMax = 220 * 1024;  // dynamic shared memory budget per CTA (227 KB max on sm_100a)

struct BlockShape {
  int B, Hp, Wp, N;               // images, patch grid, nodes per image
  int in_dim, in_pad, D;          // node feature width (padded to 4), GAT output width
  int H1, H2, H3, K;              // heads of patch / predictor / region GAT, segments
  int cluster, npc;               // CTAs per image, nodes per CTA
  float slope1, slope2, slope3;   // LeakyReLU slopes
};

// prepared-weight blob layout (floats).  Part A is staged to shared memory by every CTA.
struct PrepLayout {
  int w1t, u1, w2, u2, sizeA;     // A: W1t [H1][in_pad][D] | u1 [2H1][in_pad] | W2 [H2][K][D] | u2 [2H2][D]
  int w3t, u3, total;             // B: W3t [H3][D][D] (i-major, f fastest) | u3 [2H3][D]
};
__host__ __device__ inline int round_up4(int v) { return (v + 3) & ~3; }
__host__ __device__ inline PrepLayout prep_layout(const BlockShape& s) {
  PrepLayout p;
  int o = 0;
  p.w1t = o; o += s.H1 * s.in_pad * s.D;
  p.u1 = o;  o += 2 * s.H1 * s.in_pad;
  p.w2 = o;  o += s.H2 * s.K * s.D;
  p.u2 = o;  o += 2 * s.H2 * s.D;
  o = round_up4(o);
  p.sizeA = o;
  p.w3t = o; o += s.H3 * s.D * s.D;
  p.u3 = o;  o += 2 * s.H3 * s.D;
  p.total = round_up4(o);
  return p;
}

// shared-memory layout of the main kernel (float offsets)
struct SmemLayout {
  int prepA, prepB, s1, alpha, zs, tmp, wts, deg, S, lab, red, cmax1, cmax2, exp_part, exp_rsum, exp_rcnt, exp_fl, region, total;
};
__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ inline SmemLayout smem_layout(const BlockShape& s) {
  const PrepLayout p = prep_layout(s);
  SmemLayout L;
  int o = 4;                                    // [0..3]: two mbarriers (8 bytes each)
  L.prepA = o;    o += p.sizeA;
  L.prepB = o;    o += p.total - p.sizeA;       // W3t | u3 (used by the cluster's rank-0 CTA)
  L.s1 = o;       o += round_up4((s.npc + 2 * s.Wp) * 2 * s.H1);
  L.alpha = o;    o += kTileN * s.H1 * 4;
  L.zs = o;       o += kTileN * s.H1 * s.in_pad;
  L.tmp = o;      o += round_up4(imax(imax(s.npc * s.H2 * s.K, 3 * s.npc * s.K), (kBT / s.D) * s.K * s.D));
  L.wts = o;      o += s.npc * 4;
  L.deg = o;      o += round_up4(s.npc);
  L.S = o;        o += round_up4(s.npc * s.K);
  L.lab = o;      o += round_up4(s.npc);
  L.red = o;      o += (kBT / 32) * 8;          // block reductions: per-warp partials of up to 8 values
  L.cmax1 = o;    o += 4;
  L.cmax2 = o;    o += 4;
  L.exp_part = o; o += round_up4(2 * s.K);
  L.exp_rsum = o; o += s.K * s.D;
  L.exp_rcnt = o; o += round_up4(s.K);
  L.exp_fl = o;   o += 4;                        // this CTA's feature-consistency loss partial
  // rank-0 scratch of the region stage: R [K][D], s3 [K][2H3], a3 [K][H3][K], z3 [K][H3][D], y3 [K][H3][D]
  L.region = o;   o += s.K * s.D + round_up4(s.K * 2 * s.H3) + round_up4(s.K * s.H3 * s.K) + 2 * s.K * s.H3 * s.D;
  L.total = o;
  return L;
}

// Multi-GPU exchange fused into the kernel (mg_block_forward_push): while it computes, the kernel also stores the small
// per-image outputs (loss | region_out | labels, the layout of one packed buffer) into slice `rank` of EVERY rank's
// gathered buffer through NVLink peer mappings, and the last CTA to finish publishes the step's sequence number in every
// rank's flag array with a system-scope release.  Nothing here waits for another GPU.
struct PeerOut {
  float* const* bufs;       // DEVICE array [world]: rank p's gathered buffer as mapped into this GPU (own rank included)
  uint32_t* const* flags;   // DEVICE array [world]: rank p's flag array
  int world;                // 0: no exchange
  long long slice_off;      // floats: offset of this (slot, rank) slice inside one parity half
  long long parity_stride;  // floats between the two parity halves (steps alternate: a slice is rewritten every 2nd step)
  long long flag_index;     // this (slot, rank) flag
  uint32_t* seq;            // device: steps this slot has completed (advanced by the last CTA)
  uint32_t* done;           // device: CTAs of the current launch that have finished (re-armed by the last CTA)
};

struct BlockArgs {
  BlockShape s;
  PeerOut peer;
  const void* x;            // (B, N, in) f32|bf16
  const float* prep;        // prepared weights (mg_block_prepare)
  float* h;                 // (B, N, D)   patch-GAT output
  float* q;                 // (B, N, NQ)  scratch: predictor scalars (s2_src | s2_tgt | t[h2][c])
  float* S;                 // (B, N, K)
  int32_t* labels;          // (B, N)
  float* loss;              // (B)
  float* region_in;         // (B, K, D) or null
  float* region_out;        // (B, K, D)
  // FeatureConsistencyLoss folded into the patch-GAT epilogue (model/unet/feature_loss.py:103-123): null = off
  const float* fl_unet;     // (B, N, D) f32  U-Net patch features
  const float* fl_y;        // (B, N)    f32  patch labels
  float fl_margin;
  float* fl_out;            // (B)       per-image sums over the patches
};

// ---------------------------------------------------------------------------------------------
// weight preparation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) block_prepare_kernel(BlockShape s, const float* __restrict__ W1,
                                                           const float* __restrict__ a1, const float* __restrict__ W2,
                                                           const float* __restrict__ a2, const float* __restrict__ W3,
                                                           const float* __restrict__ a3, float* __restrict__ prep) {
  const PrepLayout p = prep_layout(s);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p.total; idx += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (idx < p.u1) {                                  // W1t[h][i][f] = W1[h][f][i], zero padded in i
      const int r = idx - p.w1t, f = r % s.D, i = (r / s.D) % s.in_pad, h = r / (s.D * s.in_pad);
      if (i < s.in_dim) v = W1[((size_t)h * s.D + f) * s.in_dim + i];
    } else if (idx < p.w2) {                           // u1[q][i] = sum_f a1[h][half*D + f] W1[h][f][i]
      const int r = idx - p.u1, i = r % s.in_pad, qq = r / s.in_pad;
      const int h = qq % s.H1, half = qq / s.H1;
      if (i < s.in_dim)
        for (int f = 0; f < s.D; ++f) v = fmaf(a1[(size_t)h * 2 * s.D + half * s.D + f], W1[((size_t)h * s.D + f) * s.in_dim + i], v);
    } else if (idx < p.u2) {                           // W2[h][c][d] as given
      v = W2[idx - p.w2];
    } else if (idx < p.u2 + 2 * s.H2 * s.D) {          // u2[q][d] = sum_c a2[h][half*K + c] W2[h][c][d]
      const int r = idx - p.u2, d = r % s.D, qq = r / s.D;
      const int h = qq % s.H2, half = qq / s.H2;
      for (int c = 0; c < s.K; ++c) v = fmaf(a2[(size_t)h * 2 * s.K + half * s.K + c], W2[((size_t)h * s.K + c) * s.D + d], v);
    } else if (idx < p.w3t) {
      v = 0.f;                                         // alignment padding
    } else if (idx < p.u3) {                           // W3t[h][i][f] = W3[h][f][i]
      const int r = idx - p.w3t, f = r % s.D, i = (r / s.D) % s.D, h = r / (s.D * s.D);
      v = W3[((size_t)h * s.D + f) * s.D + i];
    } else if (idx < p.u3 + 2 * s.H3 * s.D) {          // u3[q][i]
      const int r = idx - p.u3, i = r % s.D, qq = r / s.D;
      const int h = qq % s.H3, half = qq / s.H3;
      for (int f = 0; f < s.D; ++f) v = fmaf(a3[(size_t)h * 2 * s.D + half * s.D + f], W3[((size_t)h * s.D + f) * s.D + i], v);
    }
    prep[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> this CTA's shared memory, completion on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// destination of this step's payload in rank p's gathered buffer
__device__ __forceinline__ float* peer_slice(const PeerOut& po, int p, uint32_t parity) {
  return po.bufs[p] + po.slice_off + (long long)parity * po.parity_stride;
}
// every thread of the CTA has issued its peer stores: make them visible system-wide, count the CTA in, and let the last
// CTA of the launch publish the sequence number on every rank (fence / atomic / fence / release chain)
__device__ __forceinline__ void peer_cta_done(const PeerOut& po, uint32_t seq_now) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t old = atomicAdd(po.done, 1u);
    if (old == gridDim.x - 1) {
      __threadfence_system();
      *po.done = 0u;
      *po.seq = seq_now + 1u;
      for (int p = 0; p < po.world; ++p) st_release_sys_u32(po.flags[p] + po.flag_index, seq_now + 1u);
    }
  }
}

template <typename TX>
__device__ __forceinline__ void load4(const TX* p, float* o);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float* o) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float* o) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  o[0] = __uint_as_float(u.x << 16); o[1] = __uint_as_float(u.x & 0xffff0000u);
  o[2] = __uint_as_float(u.y << 16); o[3] = __uint_as_float(u.y & 0xffff0000u);
}

// In-neighbours of grid node n in ascending COO edge id: slot 0 up, 1 left, 2 right, 3 down.
// Fixed slots with a validity flag keep every index static (no local-memory arrays).
struct Nbr {
  int id[4];
  bool ok[4];
};
__device__ __forceinline__ Nbr grid_nbrs(int n, int Hp, int Wp) {
  const int r = n / Wp, c = n - r * Wp;
  Nbr b;
  b.ok[0] = r > 0;        b.id[0] = b.ok[0] ? n - Wp : n;
  b.ok[1] = c > 0;        b.id[1] = b.ok[1] ? n - 1 : n;
  b.ok[2] = c + 1 < Wp;   b.id[2] = b.ok[2] ? n + 1 : n;
  b.ok[3] = r + 1 < Hp;   b.id[3] = b.ok[3] ? n + Wp : n;
  return b;
}

// deterministic block reduction of nv (<= MAXV <= 8) per-thread values; result in out[0..nv)
template <bool IS_MAX, int MAXV>
__device__ __forceinline__ void block_reduce(const float (&v)[MAXV], int nv, float* red /* [warps][8] */, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv) {
      float t = v[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float u = __shfl_xor_sync(kFull, t, o);
        t = IS_MAX ? fmaxf(t, u) : t + u;
      }
      if (lane == 0) red[warp * 8 + i] = t;
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < nv) {
    float t = red[threadIdx.x];
    for (int w = 1; w < kBT / 32; ++w) t = IS_MAX ? fmaxf(t, red[w * 8 + threadIdx.x]) : t + red[w * 8 + threadIdx.x];
    out[threadIdx.x] = t;
  }
  __syncthreads();
}

__device__ __forceinline__ float sel4(const float (&a)[kMaxHeads], int h) {
  float v = a[0];
#pragma unroll
  for (int i = 1; i < kMaxHeads; ++i) v = (h == i) ? a[i] : v;
  return v;
}

// ELU(alpha=1).  exp(v)-1 instead of expm1: the absolute error (<= 1 ulp of 1 = 6e-8) is what the
// 1e-5 parity budget is about, and it is a third of the instructions.
// Branch-free, with the hardware ex2 (relative error ~2^-21 on (0,1], i.e. ~1e-7 absolute).
__device__ __forceinline__ float elu_fast(float v) {
  const float e = __expf(fminf(v, 0.f)) - 1.f;
  return v > 0.f ? v : e;
}

// Sum over the FGT lanes of a group of 32 per-lane values with a shuffle reduce-scatter: after the
// log2(FGT) stages lane gl holds the 32/FGT finished sums j = gl*(32/FGT) ... in v[0 .. 32/FGT).
template <int N, int OFF>
struct ReduceScatter {
  static __device__ __forceinline__ void run(float (&v)[32], int gl) {
    const bool up = (gl & OFF) != 0;
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
      const float send = up ? v[j] : v[j + N / 2];
      const float keep = up ? v[j + N / 2] : v[j];
      v[j] = keep + __shfl_xor_sync(kFull, send, OFF);
    }
    ReduceScatter<N / 2, OFF / 2>::run(v, gl);
  }
};
template <int N>
struct ReduceScatter<N, 0> {
  static __device__ __forceinline__ void run(float (&)[32], int) {}
};

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
template <typename TX>
__global__ void __launch_bounds__(kBT, 1) block_forward_kernel(const BlockArgs A) {
  cg::cluster_group cluster = cg::this_cluster();
  const BlockShape& s = A.s;
  const SmemLayout L = smem_layout(s);
  const PrepLayout P = prep_layout(s);
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int crank = (int)cluster.block_rank();
  const int b = blockIdx.x / s.cluster;
  const int N = s.N, Hp = s.Hp, Wp = s.Wp, D = s.D, H1 = s.H1, H2 = s.H2, K = s.K;
  const int in_dim = s.in_dim, in_pad = s.in_pad;
  const int NQ = 2 * H2 + H2 * K;
  const int n0 = min(N, crank * s.npc), n1 = min(N, n0 + s.npc), cnt = n1 - n0;
  const int h0 = max(0, n0 - Wp), h1 = min(N, n1 + Wp);
  const size_t gb = (size_t)b * N;

  float* w1t = sm + L.prepA + P.w1t;
  float* u1 = sm + L.prepA + P.u1;
  float* w2 = sm + L.prepA + P.w2;
  float* u2 = sm + L.prepA + P.u2;
  float* s1 = sm + L.s1;
  float* alpha = sm + L.alpha;
  float* zs = sm + L.zs;
  float* tmp = sm + L.tmp;
  float* wts = sm + L.wts;
  float* deg = sm + L.deg;
  float* Sown = sm + L.S;
  int* lab = reinterpret_cast<int*>(sm + L.lab);
  float* red = sm + L.red;
  void* barA = sm;
  void* barB = sm + 2;
  const PeerOut& po = A.peer;
  // steps this slot has completed so far: its parity selects the half of the peers' buffers this launch writes (the
  // last CTA advances it only after every CTA has read it: a CTA counts itself in after this load)
  const uint32_t seq_now = po.world > 0 ? *reinterpret_cast<volatile const uint32_t*>(po.seq) : 0u;
  const long long lab_off = (long long)s.B * (1 + K * D);       // labels follow loss [B] | region_out [B][K][D]

  // ---- P0: stage the prepared weights with TMA bulk copies (region weights only on rank 0) ------
  if (tid == 0) {
    mbar_init(barA, 1);
    mbar_init(barB, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytesA = (uint32_t)P.sizeA * 4u;
    mbar_expect_tx(barA, bytesA);
    tma_bulk_g2s(sm + L.prepA, A.prep, bytesA, barA);
    if (crank == 0) {
      const uint32_t bytesB = (uint32_t)(P.total - P.sizeA) * 4u;
      mbar_expect_tx(barB, bytesB);
      tma_bulk_g2s(sm + L.prepB, A.prep + P.sizeA, bytesB, barB);
    }
  }
  mbar_wait(barA, 0);

  // ---- P1: attention scalars of the patch GAT for own + halo nodes ---------------------------
  const TX* xg = reinterpret_cast<const TX*>(A.x) + gb * in_dim;
  const bool vec_in = (in_dim & 3) == 0;
  for (int t = tid; t < h1 - h0; t += kBT) {
    const int n = h0 + t;
    float acc[2 * kMaxHeads];
#pragma unroll
    for (int qq = 0; qq < 2 * kMaxHeads; ++qq) acc[qq] = 0.f;
    const TX* row = xg + (size_t)n * in_dim;
    for (int i = 0; i < in_dim; i += 4) {
      float xv[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec_in) {
        load4<TX>(row + i, xv);
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (i + v < in_dim) xv[v] = to_f32<TX>(row[i + v]);
      }
#pragma unroll
      for (int qq = 0; qq < 2 * kMaxHeads; ++qq) {
        if (qq < 2 * H1) {
          const float4 uv = *reinterpret_cast<const float4*>(u1 + qq * in_pad + i);
          acc[qq] = fmaf(xv[3], uv.w, fmaf(xv[2], uv.z, fmaf(xv[1], uv.y, fmaf(xv[0], uv.x, acc[qq]))));
        }
      }
    }
#pragma unroll
    for (int qq = 0; qq < 2 * kMaxHeads; ++qq)
      if (qq < 2 * H1) s1[t * 2 * H1 + qq] = acc[qq];
  }
  __syncthreads();

  // per-image max of s_src[i] + s_tgt[j] over edges (graph_attention.py:86), per head
  {
    float m[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) m[h] = -INFINITY;
    for (int t = tid; t < cnt; t += kBT) {
      const int n = n0 + t;
      const Nbr nb = grid_nbrs(n, Hp, Wp);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h)
          if (h < H1 && nb.ok[k]) m[h] = fmaxf(m[h], s1[(nb.id[k] - h0) * 2 * H1 + h] + s1[(n - h0) * 2 * H1 + H1 + h]);
    }
    block_reduce<true, kMaxHeads>(m, H1, red, sm + L.cmax1);
  }
  cluster.sync();                                                                   // #1
  // a few threads fetch the other CTAs' maxima through DSMEM and broadcast them in local shared memory
  if (tid < H1) {
    float m = -INFINITY;
    for (int r = 0; r < s.cluster; ++r) m = fmaxf(m, cluster.map_shared_rank(sm + L.cmax1, r)[tid]);
    red[tid] = leaky_relu(m, s.slope1);
  }
  __syncthreads();
  float M1[kMaxHeads];
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) M1[h] = h < H1 ? red[h] : 0.f;

  // ---- P2: patch GAT for own nodes, tile by tile ----------------------------------------------
  const int FG = D >> 2;                          // feature quads per node (power of two <= 32)
  const int zs_stride = H1 * in_pad;
  for (int tile = 0; tile < cnt; tile += kTileN) {
    const int tn = min(kTileN, cnt - tile);
    // (a) attention coefficients alpha[q][h][slot]
    for (int idx = tid; idx < tn * H1; idx += kBT) {
      const int qn = idx / H1, h = idx - qn * H1;
      const int n = n0 + tile + qn;
      const Nbr nb = grid_nbrs(n, Hp, Wp);
      const float st = s1[(n - h0) * 2 * H1 + H1 + h];
      const float mh = sel4(M1, h);
      float p[4], den = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        p[k] = nb.ok[k] ? __expf(leaky_relu(s1[(nb.id[k] - h0) * 2 * H1 + h] + st, s.slope1) - mh) : 0.f;
        den += p[k];
      }
      const float dn = den + 1e-10f;                                               // graph_attention.py:96
      *reinterpret_cast<float4*>(alpha + (qn * H1 + h) * 4) = make_float4(p[0] / dn, p[1] / dn, p[2] / dn, p[3] / dn);
    }
    __syncthreads();
    // (b) z[q][h][i] = sum_k alpha_k x[nbr_k][i]   (one gather of x serves every head)
    const int IQ = in_pad >> 2;
    for (int idx = tid; idx < tn * IQ; idx += kBT) {
      const int qn = idx / IQ, i4 = (idx - qn * IQ) * 4;
      const int n = n0 + tile + qn;
      const Nbr nb = grid_nbrs(n, Hp, Wp);
      float xv[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int v = 0; v < 4; ++v) xv[k][v] = 0.f;
        if (nb.ok[k]) {
          const TX* row = xg + (size_t)nb.id[k] * in_dim + i4;
          if (vec_in) {
            load4<TX>(row, xv[k]);
          } else {
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (i4 + v < in_dim) xv[k][v] = to_f32<TX>(row[v]);
          }
        }
      }
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h) {
        if (h < H1) {
          const float4 al = *reinterpret_cast<const float4*>(alpha + (qn * H1 + h) * 4);
          float4 z;
          z.x = fmaf(al.w, xv[3][0], fmaf(al.z, xv[2][0], fmaf(al.y, xv[1][0], al.x * xv[0][0])));
          z.y = fmaf(al.w, xv[3][1], fmaf(al.z, xv[2][1], fmaf(al.y, xv[1][1], al.x * xv[0][1])));
          z.z = fmaf(al.w, xv[3][2], fmaf(al.z, xv[2][2], fmaf(al.y, xv[1][2], al.x * xv[0][2])));
          z.w = fmaf(al.w, xv[3][3], fmaf(al.z, xv[2][3], fmaf(al.y, xv[1][3], al.x * xv[0][3])));
          *reinterpret_cast<float4*>(zs + (size_t)qn * zs_stride + h * in_pad + i4) = z;
        }
      }
    }
    // rows of the tile beyond tn must not feed garbage into the transform
    for (int idx = tid + tn * zs_stride; idx < kTileN * zs_stride; idx += kBT) zs[idx] = 0.f;
    __syncthreads();
    // (c) h = mean_h ELU(W_h z_h); 4 nodes x 4 features per item; then the predictor scalars
    //     q[n][v] = h_n . vec_v  (vec = u2 rows, W2 rows) reduced over the FG lanes of a node quad
    // node quads that hold real rows (a short last tile costs what it holds), rounded up to whole warps: the
    // reduce-scatter below shuffles with the full mask, so a warp enters the loop with all its lanes or not at all
    const int items = ((((tn + 3) / 4) * FG + 31) / 32) * 32;
    const float inv_h = 1.f / (float)H1;
    for (int it = tid; it < items; it += kBT) {
      const int ng = it / FG, fg = it - ng * FG;
      const int q0 = ng * 4, f0 = fg * 4;
      float o[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[r][c] = 0.f;
      for (int h = 0; h < H1; ++h) {
        // acc[r][c] += z[r][i] * W[i][c] on packed fp32 pairs (FFMA2: the same IEEE fmaf per element — labels stay bit-exact — in
        // half the instructions of the phase that is bound by issue slots)
        unsigned long long acc2[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc2[r][0] = acc2[r][1] = 0ull;
        const float* wh = w1t + (size_t)h * in_pad * D + f0;
        const float* zh = zs + (size_t)q0 * zs_stride + h * in_pad;
#pragma unroll 1
        for (int i = 0; i < in_pad; i += 4) {
          float4 zv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) zv[r] = *reinterpret_cast<const float4*>(zh + (size_t)r * zs_stride + i);
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const float4 wv = *reinterpret_cast<const float4*>(wh + (size_t)(i + ii) * D);
            const unsigned long long w01 = f32x2_pack(wv.x, wv.y), w23 = f32x2_pack(wv.z, wv.w);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float zz = ii == 0 ? zv[r].x : (ii == 1 ? zv[r].y : (ii == 2 ? zv[r].z : zv[r].w));
              const unsigned long long z2 = f32x2_pack(zz, zz);
              f32x2_fma(acc2[r][0], z2, w01);
              f32x2_fma(acc2[r][1], z2, w23);
            }
          }
        }
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          f32x2_unpack(acc2[r][0], acc[r][0], acc[r][1]);
          f32x2_unpack(acc2[r][1], acc[r][2], acc[r][3]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) o[r][c] += elu_fast(acc[r][c]);              // ELU per head, then mean (:118,:158)
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) o[r][c] *= inv_h;
        if (q0 + r < tn)
          *reinterpret_cast<float4*>(A.h + (gb + n0 + tile + q0 + r) * D + f0) = make_float4(o[r][0], o[r][1], o[r][2], o[r][3]);
      }
      // feature-consistency loss (feature_loss.py:103-123) on the rows while they are in registers: squared distance
      // to the U-Net patch features over the quad's FG lanes, then the per-patch contrastive term
      if (A.fl_unet) {
        float d2[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          d2[r] = 0.f;
          if (q0 + r < tn) {
            const float4 fu = __ldg(reinterpret_cast<const float4*>(A.fl_unet + (gb + n0 + tile + q0 + r) * D + f0));
            const float e0 = fu.x - o[r][0], e1 = fu.y - o[r][1], e2 = fu.z - o[r][2], e3 = fu.w - o[r][3];
            d2[r] = (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
          }
        }
        for (int off = FG >> 1; off > 0; off >>= 1) {
#pragma unroll
          for (int r = 0; r < 4; ++r) d2[r] += __shfl_xor_sync(kFull, d2[r], off);
        }
        if (fg == 0) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            if (q0 + r < tn) {
              const float y = __ldg(A.fl_y + gb + n0 + tile + q0 + r);
              const float hinge = fmaxf(A.fl_margin - sqrtf(d2[r] + 1e-8f), 0.f);                 // :116-118
              deg[tile + q0 + r] = y * d2[r] + (1.f - y) * (hinge * hinge);                        // :110,119 (deg is free until P3)
            }
          }
        }
      }
      // predictor scalars: 8 vectors x 4 nodes = 32 partial dot products per lane, summed over the node quad's
      // FG lanes with a reduce-scatter (30 shuffles instead of 128)
      if (FG == 16 || FG == 8) {
#pragma unroll 1
        for (int v0 = 0; v0 < NQ; v0 += 8) {
          float pr[32];
#pragma unroll
          for (int vv = 0; vv < 8; ++vv) {
            const int v = min(v0 + vv, NQ - 1);
            const float* vec = v < 2 * H2 ? u2 + v * D + f0 : w2 + (v - 2 * H2) * D + f0;
            const float4 w4 = *reinterpret_cast<const float4*>(vec);
#pragma unroll
            for (int r = 0; r < 4; ++r)
              pr[vv * 4 + r] = fmaf(o[r][3], w4.w, fmaf(o[r][2], w4.z, fmaf(o[r][1], w4.y, o[r][0] * w4.x)));
          }
          if (FG == 16) {
            ReduceScatter<32, 8>::run(pr, fg);
            const int v = v0 + (fg >> 1), r = (fg & 1) * 2;
            if (v < NQ) {
              if (q0 + r < tn) A.q[(gb + n0 + tile + q0 + r) * NQ + v] = pr[0];
              if (q0 + r + 1 < tn) A.q[(gb + n0 + tile + q0 + r + 1) * NQ + v] = pr[1];
            }
          } else {
            ReduceScatter<32, 4>::run(pr, fg);
            const int v = v0 + fg;
            if (v < NQ) {
#pragma unroll
              for (int r = 0; r < 4; ++r)
                if (q0 + r < tn) A.q[(gb + n0 + tile + q0 + r) * NQ + v] = pr[r];
            }
          }
        }
      } else {
#pragma unroll 1
        for (int v = 0; v < NQ; ++v) {
          const float* vec = v < 2 * H2 ? u2 + v * D + f0 : w2 + (v - 2 * H2) * D + f0;
          const float4 vv = *reinterpret_cast<const float4*>(vec);
          float pr[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) pr[r] = fmaf(o[r][3], vv.w, fmaf(o[r][2], vv.z, fmaf(o[r][1], vv.y, o[r][0] * vv.x)));
          for (int off = FG >> 1; off > 0; off >>= 1) {
#pragma unroll
            for (int r = 0; r < 4; ++r) pr[r] += __shfl_xor_sync(kFull, pr[r], off);
          }
          if (fg == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
              if (q0 + r < tn) A.q[(gb + n0 + tile + q0 + r) * NQ + v] = pr[r];
          }
        }
      }
    }
    __syncthreads();
  }
  if (A.fl_unet && tid < 32) {                            // this CTA's patches, fixed order: strided per lane, then the shuffle tree
    float acc = 0.f;
    for (int t = lane; t < cnt; t += 32) acc += deg[t];
    acc = warp_sum(acc);
    if (lane == 0) (sm + L.exp_fl)[0] = acc;
  }
  cluster.sync();                                                                   // #2: h, q visible

  // ---- P3: predictor max (per image, per head) + N-cut edge weights -----------------------------
  const float* qg = A.q + gb * NQ;
  const float* hg = A.h + gb * D;
  {
    float mh[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) mh[h] = -INFINITY;
    for (int t = tid; t < cnt; t += kBT) {
      const int n = n0 + t;
      const Nbr nb = grid_nbrs(n, Hp, Wp);
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h) {
        if (h < H2) {
          const float st = __ldcg(qg + (size_t)n * NQ + H2 + h);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (nb.ok[k]) mh[h] = fmaxf(mh[h], __ldcg(qg + (size_t)nb.id[k] * NQ + h) + st);
        }
      }
    }
    block_reduce<true, kMaxHeads>(mh, H2, red, sm + L.cmax2);
  }
  // w[n][slot] = exp(-|h_n - h_nbr|^2 / 2) (mincut_refinement.py:43-51): 8 lanes per node, all rows in flight
  {
    const int gl = tid & 7;
    const int groups = kBT / 8;
    const int rounds = ceil_div(cnt, groups);
    for (int rd = 0; rd < rounds; ++rd) {
      const int t = rd * groups + (tid >> 3);
      const bool live = t < cnt;
      const int n = n0 + (live ? t : 0);
      const Nbr nb = grid_nbrs(n, Hp, Wp);
      float d2[4] = {0.f, 0.f, 0.f, 0.f};
      for (int d = gl * 4; d < D; d += 32) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(hg + (size_t)n * D + d));
        float4 c[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) c[k] = __ldcg(reinterpret_cast<const float4*>(hg + (size_t)nb.id[k] * D + d));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float e0 = a.x - c[k].x, e1 = a.y - c[k].y, e2 = a.z - c[k].z, e3 = a.w - c[k].w;
          d2[k] += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        d2[k] += __shfl_xor_sync(kFull, d2[k], 4);
        d2[k] += __shfl_xor_sync(kFull, d2[k], 2);
        d2[k] += __shfl_xor_sync(kFull, d2[k], 1);
      }
      if (live && gl == 0) {
        float w[4], dsum = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          w[k] = nb.ok[k] ? __expf(-d2[k] / 2.0f) : 0.f;
          dsum += w[k];
        }
        *reinterpret_cast<float4*>(wts + t * 4) = make_float4(w[0], w[1], w[2], w[3]);
        deg[t] = dsum;                                                             // degree by source (:92-96)
      }
    }
  }
  cluster.sync();                                                                   // #3
  if (tid < H2) {
    float m = -INFINITY;
    for (int r = 0; r < s.cluster; ++r) m = fmaxf(m, cluster.map_shared_rank(sm + L.cmax2, r)[tid]);
    red[tid] = leaky_relu(m, s.slope2);
  }
  __syncthreads();
  float M2[kMaxHeads];
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) M2[h] = h < H2 ? red[h] : 0.f;

  // ---- P4: predictor GAT (transform-first: K scalars per head), softmax, argmax -------------------
  // step 1: thread per (node, head): ELU(sum_k alpha_k t[nbr_k][h][c]) -> tmp[node][h][c]
  for (int idx = tid; idx < cnt * H2; idx += kBT) {
    const int t = idx / H2, h = idx - t * H2;
    const int n = n0 + t;
    const Nbr nb = grid_nbrs(n, Hp, Wp);
    const float st = __ldcg(qg + (size_t)n * NQ + H2 + h);
    const float mh = sel4(M2, h);
    float p[4], den = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      p[k] = nb.ok[k] ? __expf(leaky_relu(__ldcg(qg + (size_t)nb.id[k] * NQ + h) + st, s.slope2) - mh) : 0.f;
      den += p[k];
    }
    const float dn = den + 1e-10f;
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] /= dn;
#pragma unroll 1
    for (int c = 0; c < K; ++c) {
      float agg = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) agg = fmaf(p[k], __ldcg(qg + (size_t)nb.id[k] * NQ + 2 * H2 + h * K + c), agg);
      tmp[(t * H2 + h) * K + c] = elu_fast(agg);
    }
  }
  __syncthreads();
  // step 2: thread per node: head mean, softmax (mincut_refinement.py:193), first-max argmax (train_end_to_end.py:356)
  {
    const float inv_h2 = 1.f / (float)H2;
    for (int t = tid; t < cnt; t += kBT) {
      const int n = n0 + t;
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < K; ++c) {
        float l = 0.f;
        for (int h = 0; h < H2; ++h) l += tmp[(t * H2 + h) * K + c];
        l *= inv_h2;
        Sown[t * K + c] = l;
        mx = fmaxf(mx, l);
      }
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < K; ++c) sum += expf(Sown[t * K + c] - mx);
      float best = -INFINITY;
      int arg = 0;
#pragma unroll 1
      for (int c = 0; c < K; ++c) {
        const float pc = expf(Sown[t * K + c] - mx) / sum;
        Sown[t * K + c] = pc;
        A.S[(gb + n) * K + c] = pc;
        if (pc > best) { best = pc; arg = c; }
      }
      lab[t] = arg;
      A.labels[gb + n] = arg;
      for (int p = 0; p < po.world; ++p)                       // consecutive threads, consecutive addresses: full sectors
        reinterpret_cast<int32_t*>(peer_slice(po, p, seq_now & 1u))[lab_off + (long long)gb + n] = arg;
    }
  }
  cluster.sync();                                                                   // #4: S visible

  // ---- P5: N-cut partial sums + region partial sums ------------------------------------------------
  {
    // thread per (node, segment): assoc, cut, label-count terms -> tmp[3][cnt*K]
    const float* Sg = A.S + gb * K;
    const int nk = cnt * K;
    for (int idx = tid; idx < nk; idx += kBT) {
      const int t = idx / K, c = idx - t * K;
      const int n = n0 + t;
      const Nbr nb = grid_nbrs(n, Hp, Wp);
      const float sc = Sown[idx];
      const float4 w = *reinterpret_cast<const float4*>(wts + t * 4);
      float cut = 0.f;
      if (nb.ok[0]) cut += w.x * sc * (1.f - __ldcg(Sg + (size_t)nb.id[0] * K + c));   // (:112-113,149)
      if (nb.ok[1]) cut += w.y * sc * (1.f - __ldcg(Sg + (size_t)nb.id[1] * K + c));
      if (nb.ok[2]) cut += w.z * sc * (1.f - __ldcg(Sg + (size_t)nb.id[2] * K + c));
      if (nb.ok[3]) cut += w.w * sc * (1.f - __ldcg(Sg + (size_t)nb.id[3] * K + c));
      tmp[idx] = sc * deg[t];                                                      // assoc (:102)
      tmp[nk + idx] = cut;
      tmp[2 * nk + idx] = (lab[t] == c) ? 1.f : 0.f;
    }
    __syncthreads();
    // one warp per (quantity, segment): fixed-order strided sum + shuffle tree -> exported partials
    const int warp = tid >> 5;
    for (int r = warp; r < 3 * K; r += kBT / 32) {
      const int which = r / K, c = r - which * K;
      float acc = 0.f;
      for (int t = lane; t < cnt; t += 32) acc += tmp[which * nk + t * K + c];
      acc = warp_sum(acc);
      if (lane == 0) {
        if (which < 2) (sm + L.exp_part)[which * K + c] = acc;
        else (sm + L.exp_rcnt)[c] = acc;                                           // exact: counts < 2^24
      }
    }
    __syncthreads();
  }
  {
    // region sums: thread (sub, d) walks own nodes sub, sub+NS, ... and adds h[n][d] to its label's slot
    float* rsum = sm + L.exp_rsum;
    const int NS = kBT / D;                               // node subsets (D <= 128)
    const int d = tid % D, sub = tid / D;
    float acc[kMaxSeg];
#pragma unroll
    for (int c = 0; c < kMaxSeg; ++c) acc[c] = 0.f;
    for (int t = sub; t < cnt; t += NS) {
      const float hv = __ldcg(hg + (size_t)(n0 + t) * D + d);
      const int l = lab[t];
#pragma unroll
      for (int c = 0; c < kMaxSeg; ++c) acc[c] += (l == c) ? hv : 0.f;
    }
    // combine the NS subsets in fixed order through tmp
#pragma unroll
    for (int c = 0; c < kMaxSeg; ++c)
      if (c < K) tmp[(sub * K + c) * D + d] = acc[c];
    __syncthreads();
    for (int idx = tid; idx < K * D; idx += kBT) {
      float t = 0.f;
      for (int sb = 0; sb < NS; ++sb) t += tmp[sb * K * D + idx];
      rsum[idx] = t;
    }
  }
  cluster.sync();                                                                   // #5: partials exported

  // ---- P6 (rank 0): loss, region means, region GAT on the complete digraph -------------------------
  const int H3 = s.H3;
  float* R = sm + L.region;
  if (crank == 0) {
    if (tid == 0) {
      float l = 0.f;
      for (int c = 0; c < K; ++c) {
        float assoc = 0.f, cut = 0.f;
        for (int r = 0; r < s.cluster; ++r) {
          const float* part = cluster.map_shared_rank(sm + L.exp_part, r);
          assoc += part[c];
          cut += part[K + c];
        }
        if (assoc > 1e-8f) l += cut / assoc;                                       // mincut_refinement.py:151-152
      }
      A.loss[b] = l;
      for (int p = 0; p < po.world; ++p) peer_slice(po, p, seq_now & 1u)[b] = l;
      if (A.fl_out) {                                     // per-image sum over the patches, CTAs in rank order
        float fl = 0.f;
        for (int r = 0; r < s.cluster; ++r) fl += cluster.map_shared_rank(sm + L.exp_fl, r)[0];
        A.fl_out[b] = fl;
      }
    }
    for (int idx = tid; idx < K * D; idx += kBT) {
      const int c = idx / D;
      float t = 0.f, n = 0.f;
      for (int r = 0; r < s.cluster; ++r) {
        t += cluster.map_shared_rank(sm + L.exp_rsum, r)[idx];
        n += cluster.map_shared_rank(sm + L.exp_rcnt, r)[c];
      }
      const float m = n > 0.f ? t / n : 0.f;                                       // train_end_to_end.py:368-373
      R[idx] = m;
      if (A.region_in) A.region_in[(size_t)b * K * D + idx] = m;
    }
  }
  cluster.sync();                                                                   // #6: remote smem no longer needed
  if (crank != 0) {
    if (po.world > 0) peer_cta_done(po, seq_now);               // its peer stores were the labels of P4
    return;
  }

  float* s3 = R + K * D;
  float* a3 = s3 + round_up4(K * 2 * H3);
  float* z3 = a3 + round_up4(K * H3 * K);
  float* y3 = z3 + K * H3 * D;
  if (K > 1) {
    mbar_wait(barB, 0);                                  // W3t | u3 landed long ago
    const float* w3t = sm + L.prepB;
    const float* u3 = w3t + (P.u3 - P.w3t);
    // scores: warp per (node, q), lanes over the feature dimension
    const int warp = tid >> 5;
    for (int idx = warp; idx < K * 2 * H3; idx += kBT / 32) {
      const int k = idx / (2 * H3), qq = idx - k * 2 * H3;
      float acc = 0.f;
      for (int i = lane; i < D; i += 32) acc = fmaf(R[k * D + i], u3[qq * D + i], acc);
      acc = warp_sum(acc);
      if (lane == 0) s3[idx] = acc;
    }
    __syncthreads();
    // attention over in-edges (sources ascending, skipping the node itself)
    for (int idx = tid; idx < K * H3; idx += kBT) {
      const int j = idx / H3, h = idx - j * H3;
      float m = -INFINITY;
      for (int jj = 0; jj < K; ++jj)
        for (int i = 0; i < K; ++i)
          if (i != jj) m = fmaxf(m, s3[i * 2 * H3 + h] + s3[jj * 2 * H3 + H3 + h]);
      const float M3 = leaky_relu(m, s.slope3);
      float den = 0.f;
      for (int i = 0; i < K; ++i) {
        float p = 0.f;
        if (i != j) {
          p = expf(leaky_relu(s3[i * 2 * H3 + h] + s3[j * 2 * H3 + H3 + h], s.slope3) - M3);
          den += p;
        }
        a3[(j * H3 + h) * K + i] = p;
      }
      const float dn = den + 1e-10f;
      for (int i = 0; i < K; ++i) a3[(j * H3 + h) * K + i] /= dn;
    }
    __syncthreads();
    for (int idx = tid; idx < K * H3 * D; idx += kBT) {
      const int i = idx % D, jh = idx / D, j = jh / H3;
      float acc = 0.f;
      for (int src = 0; src < K; ++src)
        if (src != j) acc = fmaf(a3[jh * K + src], R[src * D + i], acc);
      z3[idx] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < K * H3 * D; idx += kBT) {
      const int f = idx % D, jh = idx / D, h = jh % H3;
      const float* wcol = w3t + (size_t)h * D * D + f;
      const float* zr = z3 + (size_t)jh * D;
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll 4
      for (int i = 0; i < D; i += 4) {
        acc0 = fmaf(zr[i], wcol[(size_t)i * D], acc0);
        acc1 = fmaf(zr[i + 1], wcol[(size_t)(i + 1) * D], acc1);
        acc2 = fmaf(zr[i + 2], wcol[(size_t)(i + 2) * D], acc2);
        acc3 = fmaf(zr[i + 3], wcol[(size_t)(i + 3) * D], acc3);
      }
      y3[idx] = elu_fast((acc0 + acc1) + (acc2 + acc3));
    }
    __syncthreads();
    const float inv_h3 = 1.f / (float)H3;
    for (int idx = tid; idx < K * D; idx += kBT) {
      const int j = idx / D, f = idx - j * D;
      float t = 0.f;
      for (int h = 0; h < H3; ++h) t += y3[(j * H3 + h) * D + f];
      const float y = t * inv_h3;
      A.region_out[(size_t)b * K * D + idx] = y;
      for (int p = 0; p < po.world; ++p) peer_slice(po, p, seq_now & 1u)[s.B + (long long)b * K * D + idx] = y;
    }
  } else {
    for (int idx = tid; idx < K * D; idx += kBT) {             // :387-389 passthrough
      A.region_out[(size_t)b * K * D + idx] = R[idx];
      for (int p = 0; p < po.world; ++p) peer_slice(po, p, seq_now & 1u)[s.B + (long long)b * K * D + idx] = R[idx];
    }
  }
  if (po.world > 0) peer_cta_done(po, seq_now);
}

static bool shape_supported(const BlockShape& s, const char** why) {
  *why = nullptr;
  if (s.in_dim < 1 || s.in_dim > 64) *why = "node feature width must be in [1, 64]";
  else if (s.D != 32 && s.D != 64 && s.D != 128) *why = "GAT output width must be 32, 64 or 128";
  else if (s.H1 < 1 || s.H1 > kMaxHeads || s.H2 < 1 || s.H2 > kMaxHeads || s.H3 < 1 || s.H3 > kMaxHeads) *why = "heads must be in [1, 4]";
  else if (s.K < 1 || s.K > kMaxSeg) *why = "num_segments must be in [1, 8]";
  else if (s.Hp * s.Wp < 2) *why = "a 1x1 patch grid has no edges";
  return *why == nullptr;
}

template <typename TX>
__global__ void block_forward_kernel(const BlockArgs A);

// How many clusters of `c` CTAs (kBT threads, `smem` bytes each) the device keeps resident at once.  A cluster needs its
// CTAs inside one GPC, so this is NOT num_sms / c: on B200 only 15 clusters of 8 fit (measured: the kernel takes 38.9 us
// for 1..15 images and 67.6 us for 16, a second wave).  Cached per (cluster size, smem size class).
static int max_active_clusters(int c, size_t smem) {
  static int cache[9][8];
  static bool have[9][8];
  const int sc = (int)std::min<size_t>(7, smem / (32 * 1024));
  if (have[c][sc]) return cache[c][sc];
  auto kern = block_forward_kernel<float>;
  int n = 0;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(c * 64), 1, 1);
  cfg.blockDim = dim3(kBT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)c;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = num_sms() / c;                  // no device / query failed: optimistic estimate
  }
  cache[c][sc] = n;
  have[c][sc] = true;
  return n;
}

static int plan_cluster(BlockShape* s) {
  // experiment / tuning override: MG_BLOCK_CLUSTER = 1, 2, 4 or 8 CTAs per image (fewer CTAs leave SMs to the HBM-bound
  // kernels of the neighbouring pipeline step)
  static const int forced = getenv("MG_BLOCK_CLUSTER") ? atoi(getenv("MG_BLOCK_CLUSTER")) : 0;
  if (forced >= 1 && forced <= kMaxCluster) {
    s->cluster = forced;
    s->npc = ceil_div(s->N, forced);
    return forced;
  }
  // smallest cluster (power of two <= 8) with <= 256 nodes per CTA, else 8
  int c = 1;
  while (c < kMaxCluster && ceil_div(s->N, c) > 256) c <<= 1;
  // prefer filling the machine when the batch is small
  while (c < kMaxCluster && s->B * c * 2 <= num_sms() && ceil_div(s->N, c * 2) >= 64) c <<= 1;
  // ... but never at the price of a second wave: all B clusters must be resident together
  const int c_min = c;
  (void)c_min;
  while (c > 1 && ceil_div(s->N, c / 2) <= 256) {
    BlockShape t = *s;
    t.cluster = c; t.npc = ceil_div(s->N, c);
    if (s->B <= max_active_clusters(c, (size_t)smem_layout(t).total * 4)) break;
    c >>= 1;
  }
  s->cluster = c;
  s->npc = ceil_div(s->N, c);
  return c;
}

}  // namespace mg

using namespace mg;

static void fill_shape(BlockShape* s, int B, int Hp, int Wp, int in_dim, int D, int H1, int H2, int H3, int K, float sl1,
                       float sl2, float sl3) {
  s->B = B; s->Hp = Hp; s->Wp = Wp; s->N = Hp * Wp;
  s->in_dim = in_dim; s->in_pad = round_up4(in_dim); s->D = D;
  s->H1 = H1; s->H2 = H2; s->H3 = H3; s->K = K;
  s->slope1 = sl1; s->slope2 = sl2; s->slope3 = sl3;
  s->cluster = 1; s->npc = s->N;
}

extern "C" {

int64_t mg_block_prep_floats(int in_dim, int D, int H1, int H2, int H3, int K) {
  BlockShape s;
  fill_shape(&s, 1, 2, 2, in_dim, D, H1, H2, H3, K, 0.2f, 0.2f, 0.2f);
  return prep_layout(s).total;
}

int mg_block_supported(int B, int Hp, int Wp, int in_dim, int D, int H1, int H2, int H3, int K) {
  BlockShape s;
  fill_shape(&s, B, Hp, Wp, in_dim, D, H1, H2, H3, K, 0.2f, 0.2f, 0.2f);
  const char* why;
  if (B < 1 || Hp < 1 || Wp < 1 || !shape_supported(s, &why)) return 0;
  plan_cluster(&s);
  if ((int64_t)smem_layout(s).total * 4 > kSmemMax) return 0;
  return 1;
}

int mg_block_prepare(const float* W1, const float* a1, const float* W2, const float* a2, const float* W3, const float* a3,
                     int in_dim, int D, int H1, int H2, int H3, int K, float* prep, mg_stream_t stream) {
  MG_REQUIRE(W1 && a1 && W2 && a2 && W3 && a3 && prep, MG_ERR_INVALID, "mg_block_prepare: null pointer");
  BlockShape s;
  fill_shape(&s, 1, 2, 2, in_dim, D, H1, H2, H3, K, 0.2f, 0.2f, 0.2f);
  const char* why;
  MG_REQUIRE(shape_supported(s, &why), MG_ERR_UNSUPPORTED, "mg_block_prepare: %s", why);
  const int total = prep_layout(s).total;
  block_prepare_kernel<<<std::min(ceil_div(total, 256), num_sms() * 4), 256, 0, (cudaStream_t)stream>>>(s, W1, a1, W2, a2, W3,
                                                                                                    a3, prep);
  return check_launch("block_prepare_kernel");
}

int mg_block_forward_ex(const void* x, int x_dtype, int B, int Hp, int Wp, int in_dim, int D, int H1, int H2, int H3, int K,
                        float slope1, float slope2, float slope3, const float* prep, float* h, float* q_work, float* S,
                        int32_t* labels, float* loss, float* region_in, float* region_out, const mg_peer_out_t* peer,
                        const mg_block_feature_loss_t* floss, mg_stream_t stream) {
  MG_REQUIRE(x && prep && h && q_work && S && labels && loss && region_out, MG_ERR_INVALID, "mg_block_forward: null pointer");
  MG_REQUIRE(B > 0 && Hp > 0 && Wp > 0, MG_ERR_INVALID, "mg_block_forward: bad sizes");
  MG_REQUIRE(x_dtype == MG_F32 || x_dtype == MG_BF16, MG_ERR_INVALID, "mg_block_forward: x dtype");
  MG_REQUIRE(slope1 >= 0.f && slope2 >= 0.f && slope3 >= 0.f, MG_ERR_UNSUPPORTED, "mg_block_forward: negative LeakyReLU slope");
  BlockShape s;
  fill_shape(&s, B, Hp, Wp, in_dim, D, H1, H2, H3, K, slope1, slope2, slope3);
  const char* why;
  MG_REQUIRE(shape_supported(s, &why), MG_ERR_UNSUPPORTED, "mg_block_forward: %s", why);
  MG_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)prep % 16 == 0) && ((uintptr_t)h % 16 == 0), MG_ERR_INVALID,
             "mg_block_forward: x, prep and h must be 16-byte aligned");
  plan_cluster(&s);
  const SmemLayout L = smem_layout(s);
  const size_t smem = (size_t)L.total * 4;
  MG_REQUIRE(smem <= (size_t)kSmemMax, MG_ERR_UNSUPPORTED, "mg_block_forward: shape too large for the fused kernel (%zu B smem)", smem);
  BlockArgs A;
  A.s = s; A.x = x; A.prep = prep; A.h = h; A.q = q_work; A.S = S; A.labels = labels; A.loss = loss;
  A.region_in = region_in; A.region_out = region_out;
  A.fl_unet = nullptr; A.fl_y = nullptr; A.fl_out = nullptr; A.fl_margin = 0.f;
  if (floss) {
    MG_REQUIRE(floss->f_unet && floss->y && floss->loss_per_image, MG_ERR_INVALID, "mg_block_forward_ex: bad feature-loss descriptor");
    MG_REQUIRE((uintptr_t)floss->f_unet % 16 == 0, MG_ERR_INVALID, "mg_block_forward_ex: f_unet must be 16-byte aligned");
    A.fl_unet = floss->f_unet; A.fl_y = floss->y; A.fl_margin = floss->margin; A.fl_out = floss->loss_per_image;
  }
  A.peer.world = 0;
  if (peer) {
    MG_REQUIRE(peer->peer_bufs_dev && peer->peer_flags_dev && peer->seq && peer->done && peer->world > 0 && peer->world <= 64,
               MG_ERR_INVALID, "mg_block_forward_ex: bad peer descriptor");
    MG_REQUIRE(peer->slice_offset >= 0 && peer->parity_stride >= 0 && peer->flag_index >= 0, MG_ERR_INVALID,
               "mg_block_forward_ex: negative peer offsets");
    A.peer.bufs = reinterpret_cast<float* const*>(const_cast<void* const*>(reinterpret_cast<const void* const*>(peer->peer_bufs_dev)));
    A.peer.flags = reinterpret_cast<uint32_t* const*>(const_cast<void* const*>(reinterpret_cast<const void* const*>(peer->peer_flags_dev)));
    A.peer.world = peer->world;
    A.peer.slice_off = peer->slice_offset;
    A.peer.parity_stride = peer->parity_stride;
    A.peer.flag_index = peer->flag_index;
    A.peer.seq = peer->seq;
    A.peer.done = peer->done;
  }

  auto kern = x_dtype == MG_F32 ? block_forward_kernel<float> : block_forward_kernel<__nv_bfloat16>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax) != cudaSuccess) {
    set_error("mg_block_forward: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
    return MG_ERR_CUDA;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * s.cluster), 1, 1);
  cfg.blockDim = dim3(kBT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)s.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, A);
  if (e != cudaSuccess) {
    set_error("block_forward_kernel: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return MG_ERR_CUDA;
  }
  return check_launch("block_forward_kernel");
}

int mg_block_forward_push(const void* x, int x_dtype, int B, int Hp, int Wp, int in_dim, int D, int H1, int H2, int H3, int K,
                          float slope1, float slope2, float slope3, const float* prep, float* h, float* q_work, float* S,
                          int32_t* labels, float* loss, float* region_in, float* region_out, const mg_peer_out_t* peer,
                          mg_stream_t stream) {
  return mg_block_forward_ex(x, x_dtype, B, Hp, Wp, in_dim, D, H1, H2, H3, K, slope1, slope2, slope3, prep, h, q_work, S,
                             labels, loss, region_in, region_out, peer, nullptr, stream);
}

int mg_block_forward(const void* x, int x_dtype, int B, int Hp, int Wp, int in_dim, int D, int H1, int H2, int H3, int K,
                     float slope1, float slope2, float slope3, const float* prep, float* h, float* q_work, float* S,
                     int32_t* labels, float* loss, float* region_in, float* region_out, mg_stream_t stream) {
  return mg_block_forward_ex(x, x_dtype, B, Hp, Wp, in_dim, D, H1, H2, H3, K, slope1, slope2, slope3, prep, h, q_work, S,
                             labels, loss, region_in, region_out, nullptr, nullptr, stream);
}

}  // extern "C"
