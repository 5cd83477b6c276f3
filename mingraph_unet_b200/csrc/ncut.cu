// K6 — soft normalized-cut kernels (MinCutRefinement, model/graph_partition/mincut_refinement.py)
//   * row softmax + argmax of the predictor logits              (:193, train_end_to_end.py:356)
//   * per-edge Gaussian weights w_e = exp(-|h_src - h_tgt|^2 / 2) (:43-51)
//   * the loss  sum_c [assoc_c > 1e-8] cut_c / assoc_c          (:92-102,112-113,149-152)
// No atomics: every node writes its own partial terms, a per-graph block reduces them in a
// fixed order, so the loss is bitwise reproducible run to run.
#include "common.cuh"

namespace mg {

constexpr int kMaxK = 32;   // segments handled in registers

// thread per node
__global__ void __launch_bounds__(256) softmax_argmax_kernel(const float* __restrict__ logits, int N, int K,
                                                            float* __restrict__ S, int32_t* __restrict__ labels) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    const float* row = logits + (size_t)n * K;
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, __ldg(row + k));
    float sum = 0.f;
    for (int k = 0; k < K; ++k) sum += expf(__ldg(row + k) - m);
    float best = -INFINITY;
    int arg = 0;
    for (int k = 0; k < K; ++k) {
      const float p = expf(__ldg(row + k) - m) / sum;
      if (S) S[(size_t)n * K + k] = p;
      if (p > best) { best = p; arg = k; }              // first maximum, like torch.argmax
    }
    if (labels) labels[n] = arg;
  }
}

// squared distance of two rows with a group of G lanes (G in {8,16,32}); result on every lane of the group
template <int G>
__device__ __forceinline__ float group_sqdist(const float* __restrict__ a, const float* __restrict__ b, int D, int gl,
                                              bool vec4) {
  float acc = 0.f;
  if (vec4) {
    for (int d = gl * 4; d < D; d += G * 4) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(a + d));
      const float4 y = __ldg(reinterpret_cast<const float4*>(b + d));
      const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
      acc += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
  } else {
    for (int d = gl; d < D; d += G) {
      const float t = __ldg(a + d) - __ldg(b + d);
      acc = fmaf(t, t, acc);
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  return acc;
}

constexpr int kEG = 8;   // lanes per edge / per node group

__global__ void __launch_bounds__(256) ncut_edge_weights_kernel(const float* __restrict__ h, int N, int D,
                                                               const int64_t* __restrict__ ei, int64_t E,
                                                               float* __restrict__ w) {
  const int gl = threadIdx.x & (kEG - 1);
  const bool vec4 = (D & 3) == 0;
  const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / kEG;
  // all lanes of a warp run the same number of iterations (shuffles need the full warp)
  const int64_t iters = ceil_div64(E, groups);
  int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kEG;
  for (int64_t it = 0; it < iters; ++it, e += groups) {
    const bool ok = e < E;
    int64_t s = ok ? ei[e] : 0, t = ok ? ei[E + e] : 0;
    const bool inb = s >= 0 && s < N && t >= 0 && t < N;
    if (!inb) s = t = 0;
    const float d2 = group_sqdist<kEG>(h + (size_t)s * D, h + (size_t)t * D, D, gl, vec4);
    if (ok && gl == 0) w[e] = inb ? expf(-d2 / 2.0f) : __int_as_float(0x7fc00000);
  }
}

// group of kEG lanes per SOURCE node i: deg_i, assoc terms S_ic*deg_i, cut terms
//   node_terms[(i, c)]     = S_ic * deg_i
//   node_terms[(i, K + c)] = sum_{e: src=i} w_e * S_ic * (1 - S_{tgt,c})
__global__ void __launch_bounds__(256) ncut_node_terms_kernel(const float* __restrict__ h, const float* __restrict__ S,
                                                             const int32_t* __restrict__ rowptr,
                                                             const int32_t* __restrict__ col, int N, int D, int K,
                                                             float* __restrict__ node_terms) {
  const int gl = threadIdx.x & (kEG - 1);
  const bool vec4 = (D & 3) == 0;
  const int groups = (gridDim.x * blockDim.x) / kEG;
  const int iters = ceil_div(N, groups);
  int i = (blockIdx.x * blockDim.x + threadIdx.x) / kEG;
  for (int it = 0; it < iters; ++it, i += groups) {
    const bool ok = i < N;
    const int ic = ok ? i : 0;
    const int beg = ok ? __ldg(rowptr + ic) : 0, end = ok ? __ldg(rowptr + ic + 1) : 0;
    // every group of the warp must take part in the shuffles: iterate to the warp-wide max degree
    int deg_n = end - beg;
    int maxdeg = deg_n;
#pragma unroll
    for (int o = 16; o >= kEG; o >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(kFull, maxdeg, o));
    float deg = 0.f;
    float cut[kMaxK / kEG];                      // lane gl owns segments c = gl + q*kEG
#pragma unroll
    for (int q = 0; q < kMaxK / kEG; ++q) cut[q] = 0.f;
    const float* hi = h + (size_t)ic * D;
    for (int k = 0; k < maxdeg; ++k) {
      const bool live = k < deg_n;
      const int j = live ? __ldg(col + beg + k) : ic;
      const float d2 = group_sqdist<kEG>(hi, h + (size_t)j * D, D, gl, vec4);
      if (live) {
        const float w = expf(-d2 / 2.0f);
        deg += w;
#pragma unroll
        for (int q = 0; q < kMaxK / kEG; ++q) {
          const int c = gl + q * kEG;
          if (c < K) cut[q] += w * __ldg(S + (size_t)ic * K + c) * (1.f - __ldg(S + (size_t)j * K + c));
        }
      }
    }
    if (ok) {
#pragma unroll
      for (int q = 0; q < kMaxK / kEG; ++q) {
        const int c = gl + q * kEG;
        if (c < K) {
          node_terms[(size_t)i * 2 * K + c] = __ldg(S + (size_t)i * K + c) * deg;
          node_terms[(size_t)i * 2 * K + K + c] = cut[q];
        }
      }
    }
  }
}

// one block per graph: fixed-order tree reduction of the 2K columns over the graph's nodes
__global__ void __launch_bounds__(256) ncut_reduce_kernel(const float* __restrict__ node_terms, int nodes_per_graph,
                                                         int K, float* __restrict__ loss, float* __restrict__ stats) {
  __shared__ float red[8][2 * kMaxK];
  __shared__ float tot[2 * kMaxK];
  const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* base = node_terms + (size_t)g * nodes_per_graph * 2 * K;
  const int twoK = 2 * K;
  // thread t accumulates flat elements t, t+256, ... ; since 256 % twoK may be != 0 handle per column
  for (int c = 0; c < twoK; ++c) {
    float acc = 0.f;
    for (int n = threadIdx.x; n < nodes_per_graph; n += blockDim.x) acc += __ldg(base + (size_t)n * twoK + c);
    acc = warp_sum(acc);
    if (lane == 0) red[warp][c] = acc;
  }
  __syncthreads();
  if (threadIdx.x < twoK) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    tot[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f;
    for (int c = 0; c < K; ++c)
      if (tot[c] > 1e-8f) l += tot[K + c] / tot[c];          // mincut_refinement.py:151-152
    loss[g] = l;
  }
  if (stats && threadIdx.x < twoK) stats[(size_t)g * twoK + threadIdx.x] = tot[threadIdx.x];   // [assoc | cut]
}

// ------------------------------------------------------------------------------------------------
// backward of the loss w.r.t. the node features h and the soft assignments S (autograd dual of
// mincut_refinement.py:43-51,92-113,149-152).  With A_c = assoc_c, C_c = cut_c (saved), active c: A_c > 1e-8,
//   dL/dC_c = 1/A_c, dL/dA_c = -C_c/A_c^2,
//   dL/dw_e = sum_c S[s,c] ((1 - S[t,c])/A_c - C_c/A_c^2)                  (deg_s contains w_e)
//   dL/dS[i,c] = sum_{e: src=i} w_e ((1 - S[t,c])/A_c - C_c/A_c^2)  -  sum_{e: tgt=i} w_e S[s,c]/A_c
//   dL/dh_s -= dL/dw_e w_e (h_s - h_t),   dL/dh_t += dL/dw_e w_e (h_s - h_t)
// Owner-computes: a group of 8 lanes per node walks its out-edges (source terms) and its in-edges (target
// terms); no atomics.  D % 4 == 0, D <= 128.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ncut_backward_kernel(const float* __restrict__ h, const float* __restrict__ S,
                                                           const int32_t* __restrict__ rp_out, const int32_t* __restrict__ col_out,
                                                           const int32_t* __restrict__ rp_in, const int32_t* __restrict__ col_in,
                                                           int N, int D, int K, int nodes_per_graph,
                                                           const float* __restrict__ stats, const float* __restrict__ gloss,
                                                           float* __restrict__ gh, float* __restrict__ gS) {
  const int gl = threadIdx.x & (kEG - 1);
  const int groups = (gridDim.x * blockDim.x) / kEG;
  const int iters = ceil_div(N, groups);
  int i = (blockIdx.x * blockDim.x + threadIdx.x) / kEG;
  for (int it = 0; it < iters; ++it, i += groups) {
    const bool ok = i < N;
    const int ic = ok ? i : 0;
    const int g = nodes_per_graph > 0 ? ic / nodes_per_graph : 0;
    const float gL = __ldg(gloss + g);
    // per-lane segment coefficients: lane owns c = gl + q*kEG
    float invA[kMaxK / kEG], cA[kMaxK / kEG], Si[kMaxK / kEG], gSi[kMaxK / kEG];
#pragma unroll
    for (int q = 0; q < kMaxK / kEG; ++q) {
      const int c = gl + q * kEG;
      invA[q] = cA[q] = Si[q] = gSi[q] = 0.f;
      if (c < K) {
        const float A = __ldg(stats + (size_t)g * 2 * K + c), C = __ldg(stats + (size_t)g * 2 * K + K + c);
        if (A > 1e-8f) { invA[q] = gL / A; cA[q] = -gL * C / (A * A); }
        Si[q] = __ldg(S + (size_t)ic * K + c);
      }
    }
    float acc[4][4];                                  // g_h_i for dims d = (gl + 8t)*4 .. +3, t < 4  (D <= 128)
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[t][v] = 0.f;
    const float* hi = h + (size_t)ic * D;
    for (int pass = 0; pass < 2; ++pass) {            // 0: out-edges (i is the source), 1: in-edges (i is the target)
      const int32_t* rp = pass == 0 ? rp_out : rp_in;
      const int32_t* cl = pass == 0 ? col_out : col_in;
      const int beg = ok ? __ldg(rp + ic) : 0, end = ok ? __ldg(rp + ic + 1) : 0;
      int deg_n = end - beg, maxdeg = deg_n;
#pragma unroll
      for (int o = 16; o >= kEG; o >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(kFull, maxdeg, o));
      for (int k = 0; k < maxdeg; ++k) {
        const bool live = k < deg_n;
        const int j = live ? __ldg(cl + beg + k) : ic;
        const float* hj = h + (size_t)j * D;
        const float d2 = group_sqdist<kEG>(hi, hj, D, gl, true);
        const float w = expf(-d2 / 2.0f);
        // gw = dL/dw_e (summed over segments), and this node's dL/dS contributions
        float gw = 0.f;
#pragma unroll
        for (int q = 0; q < kMaxK / kEG; ++q) {
          const int c = gl + q * kEG;
          if (c < K && live) {
            const float Sj = __ldg(S + (size_t)j * K + c);
            if (pass == 0) {
              const float coef = (1.f - Sj) * invA[q] + cA[q];
              gw += Si[q] * coef;
              gSi[q] += w * coef;
            } else {
              gw += Sj * ((1.f - Si[q]) * invA[q] + cA[q]);
              gSi[q] -= w * Sj * invA[q];
            }
          }
        }
        gw += __shfl_xor_sync(kFull, gw, 4);
        gw += __shfl_xor_sync(kFull, gw, 2);
        gw += __shfl_xor_sync(kFull, gw, 1);
        if (live) {
          // out-edge: dL/dh_i -= gw w (h_i - h_j); in-edge (j -> i): dL/dh_i += gw w (h_j - h_i): same expression
          const float sc = -gw * w;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int d = (gl + kEG * t) * 4;
            if (d < D) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(hi + d));
              const float4 b = __ldg(reinterpret_cast<const float4*>(hj + d));
              acc[t][0] = fmaf(sc, a.x - b.x, acc[t][0]);
              acc[t][1] = fmaf(sc, a.y - b.y, acc[t][1]);
              acc[t][2] = fmaf(sc, a.z - b.z, acc[t][2]);
              acc[t][3] = fmaf(sc, a.w - b.w, acc[t][3]);
            }
          }
        }
      }
    }
    if (ok) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int d = (gl + kEG * t) * 4;
        if (d < D) *reinterpret_cast<float4*>(gh + (size_t)i * D + d) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
      }
#pragma unroll
      for (int q = 0; q < kMaxK / kEG; ++q) {
        const int c = gl + q * kEG;
        if (c < K) gS[(size_t)i * K + c] = gSi[q];
      }
    }
  }
}

}  // namespace mg

using namespace mg;

extern "C" {

int mg_softmax_argmax(const float* logits, int N, int K, float* S, int32_t* labels, mg_stream_t stream) {
  MG_REQUIRE(logits && N > 0 && K > 0, MG_ERR_INVALID, "mg_softmax_argmax: bad arguments");
  const int grid = std::min(ceil_div(N, 256), num_sms() * 8);
  softmax_argmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, N, K, S, labels);
  return check_launch("softmax_argmax_kernel");
}

int mg_ncut_edge_weights(const float* h, int N, int D, const int64_t* edge_index, int64_t E, float* w,
                         mg_stream_t stream) {
  MG_REQUIRE(h && N > 0 && D > 0 && E >= 0, MG_ERR_INVALID, "mg_ncut_edge_weights: bad arguments");
  if (E == 0) return MG_OK;
  MG_REQUIRE(edge_index && w, MG_ERR_INVALID, "mg_ncut_edge_weights: null pointer");
  MG_REQUIRE((D & 3) != 0 || ((uintptr_t)h % 16 == 0), MG_ERR_INVALID, "mg_ncut_edge_weights: h must be 16-byte aligned");
  const int grid = (int)std::min<int64_t>(ceil_div64(E * kEG, 256), (int64_t)num_sms() * 8);
  ncut_edge_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h, N, D, edge_index, E, w);
  return check_launch("ncut_edge_weights_kernel");
}

int64_t mg_ncut_work_bytes(int N, int K, int num_graphs) {
  (void)num_graphs;
  return (int64_t)N * 2 * K * 4 + 256;
}

int mg_ncut_loss(const float* h, const float* S, const int32_t* rowptr_out, const int32_t* col_out, int N, int D, int K,
                 int nodes_per_graph, float* loss, float* stats, void* work, mg_stream_t stream) {
  MG_REQUIRE(h && S && rowptr_out && loss && work && N > 0 && D > 0, MG_ERR_INVALID, "mg_ncut_loss: bad arguments");
  MG_REQUIRE(K >= 1 && K <= kMaxK, MG_ERR_UNSUPPORTED, "mg_ncut_loss: K=%d (supported 1..%d)", K, kMaxK);
  MG_REQUIRE(nodes_per_graph >= 0 && (nodes_per_graph == 0 || N % nodes_per_graph == 0), MG_ERR_INVALID,
             "mg_ncut_loss: N=%d is not a multiple of nodes_per_graph=%d", N, nodes_per_graph);
  MG_REQUIRE((D & 3) != 0 || ((uintptr_t)h % 16 == 0), MG_ERR_INVALID, "mg_ncut_loss: h must be 16-byte aligned");
  const int npg = nodes_per_graph > 0 ? nodes_per_graph : N;
  const int G = N / npg;
  float* terms = reinterpret_cast<float*>(work);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)std::min<int64_t>(ceil_div64((int64_t)N * kEG, 256), (int64_t)num_sms() * 8);
  ncut_node_terms_kernel<<<grid, 256, 0, st>>>(h, S, rowptr_out, col_out, N, D, K, terms);
  int rc;
  if ((rc = check_launch("ncut_node_terms_kernel"))) return rc;
  ncut_reduce_kernel<<<G, 256, 0, st>>>(terms, npg, K, loss, stats);
  return check_launch("ncut_reduce_kernel");
}

int mg_ncut_backward(const float* h, const float* S, const int32_t* rowptr_out, const int32_t* col_out,
                     const int32_t* rowptr_in, const int32_t* col_in, int N, int D, int K, int nodes_per_graph,
                     const float* stats, const float* grad_loss, float* grad_h, float* grad_S, mg_stream_t stream) {
  MG_REQUIRE(h && S && rowptr_out && rowptr_in && stats && grad_loss && grad_h && grad_S && N > 0, MG_ERR_INVALID,
             "mg_ncut_backward: bad arguments");
  MG_REQUIRE(K >= 1 && K <= kMaxK, MG_ERR_UNSUPPORTED, "mg_ncut_backward: K=%d (supported 1..%d)", K, kMaxK);
  MG_REQUIRE((D & 3) == 0 && D <= 128 && ((uintptr_t)h % 16 == 0) && ((uintptr_t)grad_h % 16 == 0), MG_ERR_UNSUPPORTED,
             "mg_ncut_backward: D=%d must be a multiple of 4, <= 128, with 16-byte aligned rows", D);
  MG_REQUIRE(nodes_per_graph >= 0 && (nodes_per_graph == 0 || N % nodes_per_graph == 0), MG_ERR_INVALID,
             "mg_ncut_backward: N=%d is not a multiple of nodes_per_graph=%d", N, nodes_per_graph);
  const int grid = (int)std::min<int64_t>(ceil_div64((int64_t)N * kEG, 256), (int64_t)num_sms() * 8);
  ncut_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h, S, rowptr_out, col_out, rowptr_in, col_in, N, D, K,
                                                              nodes_per_graph, stats, grad_loss, grad_h, grad_S);
  return check_launch("ncut_backward_kernel");
}

}  // extern "C"
