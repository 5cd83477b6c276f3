// Device building blocks of the fused GAT layer (shared by forward variants).
//
// Math (model/gat/graph_attention.py:53-118, re-associated):
//   s_src[n,h] = x_n . (W_h^T a_h[:F])      s_tgt[n,h] = x_n . (W_h^T a_h[F:])
//   e_ij = LeakyReLU(s_src[i] + s_tgt[j])    M = max_edges e  (per graph, per head)
//   p_ij = exp(e_ij - M)    den_j = sum_i p_ij    z_j = sum_i p_ij x_i / (den_j + 1e-10)
//   out_j = mean_h / concat_h  ELU(W_h z_j)
// so one gather of x_i serves every head and W x is never materialised.
#pragma once
#include "common.cuh"

namespace mg {

// Lane ownership of the input dimension: lane owns V consecutive elements at
// (t*32 + lane)*V for t < T; requires in_dim % V == 0 and in_dim <= 32*V*T.
template <int V, int T>
struct LaneDims {
  static constexpr int kPerLane = V * T;
  static __device__ __forceinline__ int dim(int lane, int t) { return (t * 32 + lane) * V; }
};

struct GatAggArgs {
  const void* x;          // (N, in)
  const int32_t* rowptr;  // in-CSR
  const int32_t* col;
  const float* s;         // (N, 2*heads): s_src | s_tgt
  const float* gmax;      // (G, heads) raw max of s_src+s_tgt over the graph's edges
  int N, in_dim, heads;
  int nodes_per_graph;
  float slope;
  float dropout_p;        // attention dropout (graph_attention.py:97); 0 in eval mode
  unsigned long long seed;
  const unsigned long long* seed_dev;   // optional device-side addend (CUDA-graph replays draw fresh masks)
  int z_bf16;             // gat_aggregate_kernel: spill z as bf16 (operand of the TMA-fed tensor-pipe transform)
  int dim_parts;          // gat_aggregate_kernel: warps per destination, each owning 32*V*T consecutive input dims
};

// Counter-based Bernoulli mask for attention dropout: a pure function of (seed, in-CSR slot, head), so the
// backward pass regenerates the forward's mask without storing it.  Returns 0 or 1/(1-p).
__device__ __forceinline__ float dropout_keep_scale(unsigned long long seed, unsigned slot, unsigned head, float p) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)slot * 8ull + head + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(unsigned)(z >> 40) * (1.0f / 16777216.0f);
  return u >= p ? 1.f / (1.f - p) : 0.f;
}

// Per-warp scratch in shared memory: attention numerators of one edge chunk and its sources.
struct WarpScratch {
  float p[32];
  int src[32];
};

// Aggregates one destination node j with a full warp.
//   z[h][t*V+v] <- sum_i p_ij x_i[dim]   (NOT yet normalised)
//   returns den[h] in den_out[h] (identical on every lane)
template <typename TX, int NH, int V, int T>
__device__ __forceinline__ void gat_aggregate_node(const GatAggArgs& a, int j, int lane, WarpScratch* sc,
                                                   float (&z)[NH][V * T], float (&den_out)[NH], int dim0 = 0) {
  constexpr int EPC = 32 / NH;                      // edges per chunk (one (edge,head) pair per lane)
  constexpr int U = (V * T >= 8) ? 2 : ((V * T >= 4) ? 4 : 8);   // edges whose rows are in flight together
  static_assert(EPC % U == 0 || U > EPC, "chunk/unroll mismatch");
  constexpr int UU = U > EPC ? EPC : U;
  const TX* __restrict__ x = reinterpret_cast<const TX*>(a.x);
  const int el = lane / NH, hl = lane % NH;
  const int two_h = 2 * a.heads;
  const bool head_ok = hl < a.heads;
  const int beg = __ldg(a.rowptr + j), end = __ldg(a.rowptr + j + 1);
  const int g = a.nodes_per_graph > 0 ? j / a.nodes_per_graph : 0;
  float stgt = 0.f, M = 0.f;
  if (head_ok) {
    stgt = __ldg(a.s + (size_t)j * two_h + a.heads + hl);
    M = leaky_relu(__ldg(a.gmax + (size_t)g * a.heads + hl), a.slope);
  }
#pragma unroll
  for (int h = 0; h < NH; ++h)
#pragma unroll
    for (int q = 0; q < V * T; ++q) z[h][q] = 0.f;
  float den_lane = 0.f;
  bool dim_ok[T];
#pragma unroll
  for (int t = 0; t < T; ++t) dim_ok[t] = dim0 + LaneDims<V, T>::dim(lane, t) < a.in_dim;

  for (int c = beg; c < end; c += EPC) {
    const int k = c + el;
    const bool valid = (k < end);
    int srcn = 0;
    float pv = 0.f;
    if (valid) {
      srcn = __ldg(a.col + k);
      if (head_ok) {
        const float e = leaky_relu(__ldg(a.s + (size_t)srcn * two_h + hl) + stgt, a.slope);
        pv = expf(e - M);
      }
    }
    den_lane += pv;                                  // the softmax denominator is taken before dropout (:94-97)
    if (a.dropout_p > 0.f && valid && head_ok)
      pv *= dropout_keep_scale(a.seed + (a.seed_dev ? __ldg(a.seed_dev) : 0ull), (unsigned)k, (unsigned)hl, a.dropout_p);
    __syncwarp();
    sc->p[lane] = pv;
    if (hl == 0) sc->src[el] = srcn;
    __syncwarp();
    const int ne = min(EPC, end - c);
#pragma unroll
    for (int e0 = 0; e0 < EPC; e0 += UU) {
      if (e0 < ne) {                                 // warp-uniform
        float xv[UU][V * T];
        bool ok[UU];
#pragma unroll
        for (int u = 0; u < UU; ++u) {
          ok[u] = (e0 + u) < ne;
          const int sn = ok[u] ? sc->src[e0 + u] : 0;
          const TX* row = x + (size_t)sn * a.in_dim;
#pragma unroll
          for (int t = 0; t < T; ++t) {
            if (ok[u] && dim_ok[t]) {
              VecLoad<TX, V>::ld(row + dim0 + LaneDims<V, T>::dim(lane, t), &xv[u][t * V]);
            } else {
#pragma unroll
              for (int v = 0; v < V; ++v) xv[u][t * V + v] = 0.f;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UU; ++u) {
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            const float ph = sc->p[(e0 + u) * NH + h];   // broadcast LDS; 0 for masked slots
#pragma unroll
            for (int q = 0; q < V * T; ++q) z[h][q] = fmaf(ph, xv[u][q], z[h][q]);
          }
        }
      }
    }
  }
  // den: reduce over the edge slots that share this lane's head
#pragma unroll
  for (int o = NH; o < 32; o <<= 1) den_lane += __shfl_xor_sync(kFull, den_lane, o);
#pragma unroll
  for (int h = 0; h < NH; ++h) den_out[h] = __shfl_sync(kFull, den_lane, h);
}

}  // namespace mg
