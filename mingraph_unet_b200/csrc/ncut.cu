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
                                                         int K, float* __restrict__ loss) {
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
                 int nodes_per_graph, float* loss, void* work, mg_stream_t stream) {
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
  ncut_reduce_kernel<<<G, 256, 0, st>>>(terms, npg, K, loss);
  return check_launch("ncut_reduce_kernel");
}

}  // extern "C"
