// K3 + K4 — fused multi-head GAT layer forward (MultiHeadGATLayer.forward, eval mode,
// model/gat/graph_attention.py:40-118,150-160).
//
// Launch sequence (all on the caller's stream, no host sync):
//   1. gat_scores_kernel   s[n] = (x_n . u_src[h], x_n . u_tgt[h]);  resets the per-graph max
//   2. gat_edge_max_kernel per-graph, per-head max of s_src[i]+s_tgt[j] over edges (atomic max: exact)
//   3a. gat_fused_kernel   CSR gather -> softmax-weighted aggregate (registers) -> shared-memory
//                          tile -> W_h transform from shared memory -> ELU -> head mean/concat.
//                          No atomics, no (E,F) intermediates; x is gathered once for all heads.
//   3b. (weights too large for shared memory) gat_aggregate_kernel spills z (N,heads,in) and
//       gat_transform_kernel runs a tiled FP32 GEMM with the same epilogue.
//
// HBM model: compulsory bytes = N*in*b + N*out*b + 4*(E+N+1) + 8*heads*N.
#include "gat_kernels.cuh"

namespace mg {

// ------------------------------------------------------------------------------------------
// attention vectors u[q][i], q in [0,2*heads): u_src[h] = W_h^T a_h[:F], u_tgt[h] = W_h^T a_h[F:]
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void compute_u(const float* __restrict__ W, const float* __restrict__ a, int in_dim, int F,
                                          int heads, float* u /* [2*heads][in_dim] */, int tid, int nthreads) {
  for (int idx = tid; idx < 2 * heads * in_dim; idx += nthreads) {
    const int q = idx / in_dim, i = idx - q * in_dim;
    const int h = q < heads ? q : q - heads;
    const float* av = a + (size_t)h * 2 * F + (q < heads ? 0 : F);
    const float* Wh = W + (size_t)h * F * in_dim + i;
    float acc = 0.f;
    for (int f = 0; f < F; ++f) acc = fmaf(__ldg(av + f), __ldg(Wh + (size_t)f * in_dim), acc);
    u[idx] = acc;
  }
}

// block = (head, 32 input columns); thread = (f-partition of 8, column): every thread's loads are independent and
// coalesced along the input dimension, the 8 partial sums are combined in a fixed order through shared memory
__global__ void __launch_bounds__(256) gat_u_kernel(const float* __restrict__ W, const float* __restrict__ a, int in_dim, int F,
                                                    int heads, float* __restrict__ u) {
  __shared__ float part[8][2][32];
  const int h = blockIdx.y, i0 = blockIdx.x * 32;
  const int il = threadIdx.x & 31, fp = threadIdx.x >> 5;
  const int i = i0 + il;
  float acc0 = 0.f, acc1 = 0.f;
  if (i < in_dim) {
    const float* Wh = W + (size_t)h * F * in_dim + i;
    const float* ah = a + (size_t)h * 2 * F;
#pragma unroll 8
    for (int f = fp; f < F; f += 8) {
      const float w = __ldg(Wh + (size_t)f * in_dim);
      acc0 = fmaf(__ldg(ah + f), w, acc0);
      acc1 = fmaf(__ldg(ah + F + f), w, acc1);
    }
  }
  part[fp][0][il] = acc0;
  part[fp][1][il] = acc1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int half = threadIdx.x >> 5, c = threadIdx.x & 31;
    float v = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p) v += part[p][half][c];
    if (i0 + c < in_dim) u[(size_t)(half * heads + h) * in_dim + i0 + c] = v;
  }
}

// warp per node; u in shared memory; also resets gmax (consumed by the next kernel on the stream).
// The 2*heads dot products of a node are reduced over the warp TOGETHER: each butterfly step halves the number of
// values a lane carries (NQP-1 + log2(32/NQP) shuffles instead of 5 per value), after which lane L holds the total of
// scalar q(L) given by the lane bits consumed by the halving steps.
template <typename TX, int V, int T, int NQP>
__global__ void __launch_bounds__(256) gat_scores_kernel(const TX* __restrict__ x, int N, int in_dim,
                                                         const float* __restrict__ W, const float* __restrict__ a,
                                                         const float* __restrict__ u_global, int F, int heads,
                                                         int num_graphs, float* __restrict__ s,
                                                         float* __restrict__ gmax) {
  extern __shared__ float u_s[];   // [2*heads][in_dim]
  const int nq = 2 * heads;
  if (u_global) {
    for (int i = threadIdx.x; i < nq * in_dim; i += blockDim.x) u_s[i] = __ldg(u_global + i);
  } else {
    compute_u(W, a, in_dim, F, heads, u_s, threadIdx.x, blockDim.x);
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < num_graphs * heads; i += blockDim.x) gmax[i] = -INFINITY;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  // scalar index this lane ends up holding, and the lane that stores it
  int my_q = 0;
  {
    int n = NQP, bit = 0;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      if (n > 1) { n >>= 1; if (lane & o) my_q += n; bit |= o; }
    }
    (void)bit;
  }
  constexpr int kTail = 32 / NQP;                  // lanes sharing one scalar after the halving steps
  for (int n = warp_global; n < N; n += nwarps) {
    float xv[V * T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int d = LaneDims<V, T>::dim(lane, t);
      if (d < in_dim) {
        VecLoad<TX, V>::ld(x + (size_t)n * in_dim + d, &xv[t * V]);
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) xv[t * V + v] = 0.f;
      }
    }
    float vals[NQP];
#pragma unroll
    for (int q = 0; q < NQP; ++q) {
      float acc = 0.f;
      if (q < nq) {
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int d = LaneDims<V, T>::dim(lane, t);
          if (d < in_dim) {
            if (V == 4) {
              const float4 u4 = *reinterpret_cast<const float4*>(u_s + q * in_dim + d);
              acc = fmaf(xv[t * V], u4.x, acc);
              acc = fmaf(xv[t * V + 1], u4.y, acc);
              acc = fmaf(xv[t * V + 2], u4.z, acc);
              acc = fmaf(xv[t * V + 3], u4.w, acc);
            } else if (V == 2) {
              const float2 u2 = *reinterpret_cast<const float2*>(u_s + q * in_dim + d);
              acc = fmaf(xv[t * V], u2.x, acc);
              acc = fmaf(xv[t * V + 1], u2.y, acc);
            } else {
#pragma unroll
              for (int v = 0; v < V; ++v) acc = fmaf(xv[t * V + v], u_s[q * in_dim + d + v], acc);
            }
          }
        }
      }
      vals[q] = acc;
    }
    {
      int cnt = NQP;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        if (cnt > 1) {
          cnt >>= 1;
          const bool upper = (lane & o) != 0;
#pragma unroll
          for (int i = 0; i < NQP / 2; ++i) {
            if (i < cnt) {
              const float send = upper ? vals[i] : vals[i + cnt];
              const float keep = upper ? vals[i + cnt] : vals[i];
              vals[i] = keep + __shfl_xor_sync(kFull, send, o);
            }
          }
        } else {
          vals[0] += __shfl_xor_sync(kFull, vals[0], o);
        }
      }
    }
    if ((lane & (kTail - 1)) == 0 && my_q < nq) s[(size_t)n * nq + my_q] = vals[0];
  }
}

// thread per destination node; exact per-graph max via order-independent atomic max
__global__ void __launch_bounds__(256) gat_edge_max_kernel(const int32_t* __restrict__ rowptr,
                                                           const int32_t* __restrict__ col,
                                                           const float* __restrict__ s, int N, int heads,
                                                           int nodes_per_graph, float* __restrict__ gmax) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int nq = 2 * heads;
  const int jc = j < N ? j : N - 1;
  const int g = nodes_per_graph > 0 ? jc / nodes_per_graph : 0;
  const int g0 = __shfl_sync(kFull, g, 0);
  const bool uniform = __all_sync(kFull, g == g0);
  const int beg = j < N ? rowptr[j] : 0, end = j < N ? rowptr[j + 1] : 0;
  for (int h = 0; h < heads; ++h) {
    float m = -INFINITY;
    if (end > beg) {
      const float st = s[(size_t)j * nq + heads + h];
      for (int k = beg; k < end; ++k) m = fmaxf(m, s[(size_t)col[k] * nq + h] + st);
    }
    if (uniform) {
      m = warp_max(m);
      if (lane == 0 && m > -INFINITY) atomic_max_f32(gmax + (size_t)g0 * heads + h, m);
    } else if (m > -INFINITY) {
      atomic_max_f32(gmax + (size_t)g * heads + h, m);
    }
  }
}


// out[n][f] = mean_h/concat_h ELU( sum_i z[n][h][i] * W[h][f][i] ); 64x64 tile, 4x4 per thread, K-step 16
constexpr int kTM = 64, kTN = 64, kTK = 16;
__global__ void __launch_bounds__(256) gat_transform_kernel(const float* __restrict__ z, const float* __restrict__ W,
                                                            int N, int in_dim, int F, int heads, int concat,
                                                            void* __restrict__ out, int out_bf16) {
  __shared__ float As[kTK][kTM + 4];
  __shared__ float Bs[kTK][kTN + 4];
  const int n_base = blockIdx.x * kTM, f_base = blockIdx.y * kTN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // tx -> f, ty -> n
  const int out_w = concat ? heads * F : F;
  float oacc[4][4] = {};
  for (int h = 0; h < heads; ++h) {
    float acc[4][4] = {};
    for (int k0 = 0; k0 < in_dim; k0 += kTK) {
      // stage A (z rows of head h) and B (W_h rows) transposed to [k][row]
      for (int idx = threadIdx.x; idx < kTM * kTK; idx += 256) {
        const int r = idx / kTK, k = idx - r * kTK;
        const int n = n_base + r, kk = k0 + k;
        As[k][r] = (n < N && kk < in_dim) ? __ldg(z + ((size_t)n * heads + h) * in_dim + kk) : 0.f;
        const int f = f_base + r;
        Bs[k][r] = (f < F && kk < in_dim) ? __ldg(W + ((size_t)h * F + f) * in_dim + kk) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kTK; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(aa[r], bb[c], acc[r][c]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float v = elu1(acc[r][c]);
        if (concat) {
          const int n = n_base + ty * 4 + r, f = f_base + tx * 4 + c;
          if (n < N && f < F) {
            const size_t o = (size_t)n * out_w + (size_t)h * F + f;
            if (out_bf16) reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16_rn(v);
            else reinterpret_cast<float*>(out)[o] = v;
          }
        } else {
          oacc[r][c] += v;
        }
      }
  }
  if (!concat) {
    const float inv_h = 1.f / (float)heads;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int n = n_base + ty * 4 + r, f = f_base + tx * 4 + c;
        if (n < N && f < F) {
          const size_t o = (size_t)n * out_w + f;
          const float v = oacc[r][c] * inv_h;
          if (out_bf16) reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16_rn(v);
          else reinterpret_cast<float*>(out)[o] = v;
        }
      }
  }
}

// ------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------
static bool pick_dims(int in_dim, DimCfg* c) {
  static const DimCfg cands[] = {{1, 1}, {2, 1}, {4, 1}, {4, 2}, {4, 4}};
  for (const DimCfg& d : cands)
    if (in_dim % d.V == 0 && in_dim <= 32 * d.V * d.T) { *c = d; return true; }
  return false;
}

static inline int round4(int v) { return (v + 3) & ~3; }

struct FusedPlan { bool ok; int tile_nodes, rn; size_t smem; };
static FusedPlan plan_fused(int in_dim, int F, int heads, DimCfg d, int N = 1 << 30) {
  FusedPlan p{false, 0, 1, 0};
  if (d.V * d.T > 4 || in_dim > 128) return p;             // fused variants are instantiated for in <= 128
  const int in_pad = round4(in_dim), f_pad = round4(F);
  const int64_t wt = (int64_t)heads * in_pad * f_pad * 4;
  const int64_t scratch = 8 * (int64_t)sizeof(WarpScratch);
  const int FG = f_pad / 4;
  // prefer two resident blocks per SM (<= ~100 KB each) so gather and transform phases overlap
  for (int pass = 0; pass < 2 && !p.ok; ++pass) {
    const int64_t budget = pass == 0 ? 100 * 1024 : kSmemBudget;
    // small graphs (a training shard is 4 x 1024 nodes): smaller tiles so that every SM gets one — 128-node tiles left
    // 116 of 148 SMs idle and the layer took 71 us for 4096 nodes
    int tn_max = 128;
    while (tn_max > 16 && ceil_div(N, tn_max) < num_sms()) tn_max >>= 1;
    for (int tn = tn_max; tn >= 8; tn >>= 1) {
      const int64_t tot = wt + (int64_t)tn * heads * in_pad * 4 + scratch;
      if (tot <= budget) { p.ok = true; p.tile_nodes = tn; p.smem = (size_t)tot; break; }
    }
  }
  if (!p.ok) return p;
  p.rn = 1;
  for (int rn = 8; rn >= 1; rn >>= 1)
    if (p.tile_nodes % rn == 0 && (p.tile_nodes / rn) * FG >= 256) { p.rn = rn; break; }
  return p;
}

// per-dtype instantiations live in gat_inst_f32.cu / gat_inst_bf16.cu
int gat_launch_fused_f32(const GatFusedArgs& A, int NH, DimCfg d, size_t smem, int grid, cudaStream_t st);
int gat_launch_fused_bf16(const GatFusedArgs& A, int NH, DimCfg d, size_t smem, int grid, cudaStream_t st);
int gat_launch_agg_f32(const GatAggArgs& a, int NH, DimCfg d, float* z, float* den, int grid, cudaStream_t st);
int gat_launch_agg_bf16(const GatAggArgs& a, int NH, DimCfg d, float* z, float* den, int grid, cudaStream_t st);


template <typename TX>
static int launch_scores(const void* x, int N, int in_dim, const float* W, const float* a, const float* u_global, int F,
                         int heads, int G, float* s, float* gmax, DimCfg d, cudaStream_t st) {
  const size_t smem = (size_t)2 * heads * in_dim * 4;
  const int grid = (int)std::min<int64_t>(ceil_div64((int64_t)N * 32, 256), (int64_t)num_sms() * 8);
  const TX* xx = reinterpret_cast<const TX*>(x);
  const int nq = 2 * heads;
  const int nqp = nq <= 2 ? 2 : (nq <= 4 ? 4 : (nq <= 8 ? 8 : 16));
#define MG_SC(VV, TT)                                                                                                  \
  if (d.V == VV && d.T == TT) {                                                                                        \
    if (nqp == 2) gat_scores_kernel<TX, VV, TT, 2><<<grid, 256, smem, st>>>(xx, N, in_dim, W, a, u_global, F, heads, G, s, gmax);        \
    else if (nqp == 4) gat_scores_kernel<TX, VV, TT, 4><<<grid, 256, smem, st>>>(xx, N, in_dim, W, a, u_global, F, heads, G, s, gmax);   \
    else if (nqp == 8) gat_scores_kernel<TX, VV, TT, 8><<<grid, 256, smem, st>>>(xx, N, in_dim, W, a, u_global, F, heads, G, s, gmax);   \
    else gat_scores_kernel<TX, VV, TT, 16><<<grid, 256, smem, st>>>(xx, N, in_dim, W, a, u_global, F, heads, G, s, gmax);               \
    return check_launch("gat_scores_kernel");                                                                          \
  }
  MG_SC(1, 1) MG_SC(2, 1) MG_SC(4, 1) MG_SC(4, 2) MG_SC(4, 4)
#undef MG_SC
  set_error("gat_scores: no variant for V=%d T=%d", d.V, d.T);
  return MG_ERR_UNSUPPORTED;
}


// scores s (N, 2*heads) = [s_src | s_tgt] and the per-graph raw edge maxima gmax (G, heads); shared with backward
int gat_scores_and_max(const void* x, int x_dtype, const int32_t* rowptr, const int32_t* col, int N, const float* W,
                       const float* a, int in_dim, int out_dim, int heads, int nodes_per_graph, float* s, float* gmax,
                       float* u, cudaStream_t st, int64_t E) {
  DimCfg d;
  if (!pick_dims(in_dim, &d)) {
    set_error("gat: in_dim=%d unsupported", in_dim);
    return MG_ERR_UNSUPPORTED;
  }
  const int G = nodes_per_graph > 0 ? N / nodes_per_graph : 1;
  int rc;
  // attention vectors: recomputed per block when cheap, otherwise one small kernel
  const float* u_global = nullptr;
  if ((int64_t)in_dim * out_dim * heads > 32768 || u != nullptr) {
    gat_u_kernel<<<dim3(ceil_div(in_dim, 32), heads), 256, 0, st>>>(W, a, in_dim, out_dim, heads, u);
    if ((rc = check_launch("gat_u_kernel"))) return rc;
    u_global = u;
  }
  if (x_dtype == MG_F32)
    rc = launch_scores<float>(x, N, in_dim, W, a, u_global, out_dim, heads, G, s, gmax, d, st);
  else
    rc = launch_scores<__nv_bfloat16>(x, N, in_dim, W, a, u_global, out_dim, heads, G, s, gmax, d, st);
  if (rc) return rc;
  if (E >= 0 && N >= 4096 && gat_tc_edge_max_supported(heads))        // several gathers in flight per destination (gat_tc.cu)
    return gat_tc_edge_max(rowptr, col, s, N, E, heads, nodes_per_graph, gmax, st);
  gat_edge_max_kernel<<<ceil_div(N, 256), 256, 0, st>>>(rowptr, col, s, N, heads, nodes_per_graph, gmax);
  return check_launch("gat_edge_max_kernel");
}

struct WorkLayout { size_t s_off, gmax_off, u_off, wb_off, z_off, total; };
static WorkLayout work_layout(int N, int in_dim, int out_dim, int heads, int G, bool need_z) {
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  WorkLayout w;
  size_t o = 0;
  w.s_off = o;    o = al(o + (size_t)N * 2 * heads * 4);
  w.gmax_off = o; o = al(o + (size_t)G * heads * 4);
  w.u_off = o;    o = al(o + (size_t)2 * heads * in_dim * 4);
  w.wb_off = o;   o = al(o + (size_t)gat_transform_tma_wbytes(in_dim, out_dim, heads));     // bf16 copy of W (TMA transform)
  w.z_off = o;    if (need_z) o = al(o + (size_t)N * heads * in_dim * 4);
  w.total = o;
  return w;
}

}  // namespace mg

using namespace mg;

extern "C" {

int64_t mg_gat_work_bytes(int N, int in_dim, int out_dim, int heads, int num_graphs) {
  if (N <= 0 || in_dim <= 0 || out_dim <= 0 || heads <= 0) return 0;
  DimCfg d;
  bool need_z = true;
  if (pick_dims(in_dim, &d))
    need_z = !plan_fused(in_dim, out_dim, heads, d).ok || gat_transform_tc_supported(N, in_dim, out_dim, heads, 1) ||
             gat_transform_tc_supported(N, in_dim, out_dim, heads, 3) || gat_transform_tma_supported(N, in_dim, out_dim, heads);
  return (int64_t)work_layout(N, in_dim, out_dim, heads, num_graphs > 0 ? num_graphs : 1, need_z).total;
}

int mg_gat_uses_tensor_pipe(int N, int in_dim, int out_dim, int heads, int concat, int x_dtype, int out_dtype) {
  return (x_dtype == MG_BF16 && gat_tc_supported(N, in_dim, out_dim, heads, concat ? 1 : 0, out_dtype == MG_BF16 ? 1 : 0)) ? 1 : 0;
}

int mg_gat_forward(const void* x, int x_dtype, const int32_t* rowptr, const int32_t* col, int N, int64_t E,
                   const float* W, const float* a, int in_dim, int out_dim, int heads, int concat, float slope,
                   int nodes_per_graph, float dropout_p, uint64_t seed, const uint64_t* seed_dev, void* out, int out_dtype,
                   void* work,
                   float* save_den, float* save_z, mg_stream_t stream) {
  MG_REQUIRE(x && rowptr && W && a && out && work, MG_ERR_INVALID, "mg_gat_forward: null pointer");
  MG_REQUIRE(N > 0 && in_dim > 0 && out_dim > 0, MG_ERR_INVALID, "mg_gat_forward: bad sizes N=%d in=%d out=%d", N, in_dim,
             out_dim);
  MG_REQUIRE(heads >= 1 && heads <= 8, MG_ERR_UNSUPPORTED, "mg_gat_forward: heads=%d (supported 1..8)", heads);
  MG_REQUIRE(E > 0 && col, MG_ERR_INVALID,
             "mg_gat_forward: empty edge_index (the reference raises in torch.max, graph_attention.py:86)");
  MG_REQUIRE(slope >= 0.f, MG_ERR_UNSUPPORTED, "mg_gat_forward: negative LeakyReLU slope is not monotone");
  MG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, MG_ERR_INVALID, "mg_gat_forward: dropout_p must be in [0, 1)");
  MG_REQUIRE(x_dtype == MG_F32 || x_dtype == MG_BF16, MG_ERR_INVALID, "mg_gat_forward: x dtype");
  MG_REQUIRE(out_dtype == MG_F32 || out_dtype == MG_BF16, MG_ERR_INVALID, "mg_gat_forward: out dtype");
  MG_REQUIRE(nodes_per_graph >= 0 && (nodes_per_graph == 0 || N % nodes_per_graph == 0), MG_ERR_INVALID,
             "mg_gat_forward: N=%d is not a multiple of nodes_per_graph=%d", N, nodes_per_graph);
  DimCfg d;
  MG_REQUIRE(pick_dims(in_dim, &d), MG_ERR_UNSUPPORTED, "mg_gat_forward: in_dim=%d unsupported (in <= 32, even in <= 64, or in %% 4 == 0 and in <= 512; pad the\n"
             "feature dimension with zero columns otherwise)", in_dim);
  const int G = nodes_per_graph > 0 ? N / nodes_per_graph : 1;
  const int NH = heads <= 1 ? 1 : (heads <= 2 ? 2 : (heads <= 4 ? 4 : 8));
  FusedPlan plan = plan_fused(in_dim, out_dim, heads, d, N);
  // large transforms go to the tensor pipe: aggregate to z, then a tcgen05 GEMM (tf32 for bf16 storage, 3xTF32 for fp32)
  const int tc_passes = x_dtype == MG_BF16 ? 1 : 3;
  const bool tc_gemm = gat_transform_tc_supported(N, in_dim, out_dim, heads, tc_passes);
  const bool any_tc_gemm = gat_transform_tc_supported(N, in_dim, out_dim, heads, 1) || gat_transform_tc_supported(N, in_dim, out_dim, heads, 3) ||
                           gat_transform_tma_supported(N, in_dim, out_dim, heads);
  // bf16 storage, inference: z spilled as bf16 and transformed by the persistent TMA-fed kernel (gat_tma_gemm.cu)
  const bool tma_gemm = x_dtype == MG_BF16 && !save_z && gat_transform_tma_supported(N, in_dim, out_dim, heads);
  const WorkLayout wl = work_layout(N, in_dim, out_dim, heads, G, !plan.ok || any_tc_gemm);
  if (tc_gemm || tma_gemm) plan.ok = false;
  unsigned char* wb = reinterpret_cast<unsigned char*>(work);
  float* s = reinterpret_cast<float*>(wb + wl.s_off);
  float* gmax = reinterpret_cast<float*>(wb + wl.gmax_off);
  float* u = reinterpret_cast<float*>(wb + wl.u_off);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;

  // bf16 storage, inference: score pre-pass + transform on the tensor pipe (gat_tc.cu)
  if (x_dtype == MG_BF16 && dropout_p == 0.f && !save_den && !save_z &&
      gat_tc_supported(N, in_dim, out_dim, heads, concat ? 1 : 0, out_dtype == MG_BF16 ? 1 : 0))
    return gat_tc_launch(x, rowptr, col, s, gmax, u, W, a, N, E, in_dim, out_dim, heads, concat ? 1 : 0, slope, nodes_per_graph, out,
                         out_dtype == MG_BF16 ? 1 : 0, st);

  // bf16 features: the mma.sync score pre-pass (4-5x faster than the FP32-pipe kernels at in = 128 / 256); everything else the generic one
  if (x_dtype == MG_BF16 && gat_tc_prepass_supported(N, in_dim, heads))
    rc = gat_tc_prepass(x, rowptr, col, N, E, W, a, in_dim, out_dim, heads, nodes_per_graph, s, gmax, u, st);
  else
    rc = gat_scores_and_max(x, x_dtype, rowptr, col, N, W, a, in_dim, out_dim, heads, nodes_per_graph, s, gmax, u, st, E);
  if (rc) return rc;

  GatAggArgs ag;
  ag.x = x; ag.rowptr = rowptr; ag.col = col; ag.s = s; ag.gmax = gmax;
  ag.N = N; ag.in_dim = in_dim; ag.heads = heads; ag.nodes_per_graph = nodes_per_graph; ag.slope = slope;
  ag.dropout_p = dropout_p; ag.seed = seed; ag.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  ag.z_bf16 = tma_gemm ? 1 : 0;
  ag.dim_parts = 1;

  if (plan.ok) {
    GatFusedArgs A;
    A.agg = ag; A.W = W; A.out = out; A.save_den = save_den; A.save_z = save_z;
    A.F = out_dim; A.concat = concat ? 1 : 0; A.out_bf16 = out_dtype == MG_BF16;
    A.tile_nodes = plan.tile_nodes; A.rn = plan.rn; A.in_pad = round4(in_dim); A.f_pad = round4(out_dim);
    const int ntiles = ceil_div(N, plan.tile_nodes);
    const int per_sm = std::max<int>(1, (int)std::min<int64_t>(4, (220 * 1024) / (int64_t)(plan.smem + 1024)));
    const int grid = std::min(ntiles, num_sms() * per_sm);
    if (x_dtype == MG_F32) return gat_launch_fused_f32(A, NH, d, plan.smem, grid, st);
    return gat_launch_fused_bf16(A, NH, d, plan.smem, grid, st);
  }
  // unfused: z -> global (also serves as save_z when the caller wants it)
  float* z = save_z ? save_z : reinterpret_cast<float*>(wb + wl.z_off);
  // wide rows: one warp per 256 input dims of a destination.  Measured at in = 512, N = 262 144, k = 8 (ncu): one warp per
  // row (T=4) is latency-bound (1.2 ms), one warp per 128 dims is issue-bound (842 M warp instructions: every part
  // recomputes the attention numerators); 256 dims per warp sits between the two.
  DimCfg dagg = d;
  if (in_dim > 256 && in_dim % 256 == 0) { dagg = DimCfg{4, 2}; ag.dim_parts = in_dim / 256; }
  const int grid = (int)std::min<int64_t>(ceil_div64((int64_t)N * ag.dim_parts * 32, 256), (int64_t)num_sms() * 16);
  if (tma_gemm && !save_den && dropout_p == 0.f && gat_agg_spill_supported(N, in_dim, heads))
    rc = gat_agg_spill_launch(x, rowptr, col, s, gmax, z, N, in_dim, slope, nodes_per_graph, st);      // mma.sync aggregation warps
  else if (x_dtype == MG_F32) rc = gat_launch_agg_f32(ag, NH, dagg, z, save_den, grid, st);
  else rc = gat_launch_agg_bf16(ag, NH, dagg, z, save_den, grid, st);
  if (rc) return rc;
  if (tma_gemm)
    return gat_transform_tma_launch(z, W, wb + wl.wb_off, N, in_dim, out_dim, heads, concat ? 1 : 0, out, out_dtype == MG_BF16 ? 1 : 0, st);
  if (tc_gemm)
    return gat_transform_tc_launch(z, W, N, in_dim, out_dim, heads, concat ? 1 : 0, out, out_dtype == MG_BF16 ? 1 : 0, tc_passes, st);
  dim3 g2(ceil_div(N, kTM), ceil_div(out_dim, kTN));
  gat_transform_kernel<<<g2, 256, 0, st>>>(z, W, N, in_dim, out_dim, heads, concat ? 1 : 0, out,
                                           out_dtype == MG_BF16 ? 1 : 0);
  return check_launch("gat_transform_kernel");
}

}  // extern "C"
