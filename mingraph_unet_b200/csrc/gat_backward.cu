// K5 — backward of the multi-head GAT layer (the autograd dual of model/gat/graph_attention.py:53-118,
// 150-160: IndexBackward / ScatterAddBackward / MmBackward chains of the reference's graph).
//
// With the forward's re-association (z_j^h = sum_i a_ij x_i, y = W_h z, o = ELU(y), out = mean/concat):
//   g_y   = g_o * ELU'(y)                       g_W_h += g_y (x) z          g_z = W_h^T g_y
//   c_j   = g_z_j . z_j                         (= sum_k a_kj g_a_kj, the softmax correction term)
//   g_a   = m_ij (g_z_j . x_i)                  (m = dropout mask / (1-p), 1 in eval)
//   g_pre = a_ij (g_a - c_j) * LeakyReLU'(pre)  pre = s_src[i] + s_tgt[j]
//   g_s_tgt[j] = sum_i g_pre    g_s_src[i] = sum_j g_pre
//   g_x_i = sum_j sum_h m a_ij g_z_j^h  +  sum_h (g_s_src[i,h] u_src^h + g_s_tgt[i,h] u_tgt^h)
//   g_u_src^h = sum_n g_s_src[n,h] x_n  (same for tgt);  g_W_h += a_src (x) g_u_src + a_tgt (x) g_u_tgt
//   g_a_h = [W_h g_u_src | W_h g_u_tgt]
// The softmax shift M = LeakyReLU(max_edges pre) also carries gradient in the reference (MaxBackward): with
// a_ij = p_ij/(D_j + eps), d a_ij / dM = -a_ij eps/(D_j + eps), so g_M = -sum_j eps/(D_j+eps) c_j, routed to the
// arg-max edge(s) (evenly over ties, like torch's full-reduction max).  It only matters where D_j ~ eps.
//
// Two passes over the edges keep every reduction owner-computes (no atomics, deterministic):
// by TARGET (in-CSR) for g_pre / g_s_tgt, by SOURCE (out-CSR) for g_x / g_s_src; weight gradients are
// split over node chunks and reduced in a fixed order.
#include "gat_kernels.cuh"

namespace mg {

constexpr int kBwdTile = 32;      // nodes per tile in the node kernel
constexpr int kMaxSplits = 256;   // node chunks for the weight-gradient reductions

struct BwdArgs {
  const void* x;            // (N, in)
  const float* W;           // (H, F, in)
  const float* a;           // (H, 2F)
  const float* u;           // (2H, in)   attention vectors
  const float* s;           // (N, 2H)    scores
  const float* gmax;        // (G, H)
  const float* den;         // (N, H)     saved by forward
  const float* z;           // (N, H, in) saved by forward (normalised aggregate)
  const float* gout;        // (N, out_w)
  const int32_t *rowptr_in, *col_in, *rowptr_out, *col_out, *slot_out2in;
  float* gy;                // (N, H, F)
  float* gz;                // (N, H, in)
  float* c;                 // (N, H)
  float* ea;                // (E, H)  m * alpha   (in-CSR slot order)
  float* eg;                // (E, H)  g_pre
  float* gs;                // (N, 2H) g_s_src | g_s_tgt
  float* gmpart;            // (N, 2H) per-node terms of g_M | arg-max tie counts
  float* gM;                // (G, 2H) g_M | tie count
  float* gx;                // (N, in)
  float* partW;             // (S, H, F, in)
  float* partU;             // (S, 2H, in)
  float* gW;                // (H, F, in)
  float* ga;                // (H, 2F)
  int N, in_dim, F, heads, concat, nodes_per_graph, splits;
  float slope, dropout_p;
  unsigned long long seed;
  const unsigned long long* seed_dev;
};

// ---------------------------------------------------------------------------------------------
// B1: per node/head  y = W z, g_y = g_o * ELU'(y), g_z = W^T g_y, c = g_z . z
// block = tile of kBwdTile nodes; W_h staged in shared memory head by head
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gat_bwd_node_kernel(const BwdArgs A) {
  extern __shared__ __align__(16) float smem[];
  const int in_dim = A.in_dim, F = A.F, H = A.heads;
  float* Ws = smem;                               // [F][in_dim + 1]
  float* Zs = Ws + (size_t)F * (in_dim + 1);       // [kBwdTile][in_dim + 1]
  float* Gy = Zs + (size_t)kBwdTile * (in_dim + 1);  // [kBwdTile][F + 1]
  const int out_w = A.concat ? H * F : F;
  const float ginv = A.concat ? 1.f : 1.f / (float)H;
  const int ntiles = ceil_div(A.N, kBwdTile);
  // grid (tiles, heads): heads are independent, so a small graph (the K-node region graphs of a training shard) still
  // spreads over heads blocks instead of walking the heads one after the other in a single block
  const int h = blockIdx.y;
  for (int idx = threadIdx.x; idx < F * in_dim; idx += blockDim.x) {
    const int f = idx / in_dim, i = idx - f * in_dim;
    Ws[f * (in_dim + 1) + i] = __ldg(A.W + (size_t)h * F * in_dim + idx);
  }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int nb = tile * kBwdTile;
    const int tn = min(kBwdTile, A.N - nb);           // only the tile's real nodes are worked on
    __syncthreads();
    for (int idx = threadIdx.x; idx < tn * in_dim; idx += blockDim.x) {
      const int q = idx / in_dim, i = idx - q * in_dim;
      Zs[q * (in_dim + 1) + i] = __ldg(A.z + ((size_t)(nb + q) * H + h) * in_dim + i);
    }
    __syncthreads();
    // y and g_y: thread per (node, f)
    for (int idx = threadIdx.x; idx < tn * F; idx += blockDim.x) {
      const int q = idx / F, f = idx - q * F;
      const int n = nb + q;
      float y0 = 0.f, y1 = 0.f;
      int i = 0;
      for (; i + 1 < in_dim; i += 2) {
        y0 = fmaf(Zs[q * (in_dim + 1) + i], Ws[f * (in_dim + 1) + i], y0);
        y1 = fmaf(Zs[q * (in_dim + 1) + i + 1], Ws[f * (in_dim + 1) + i + 1], y1);
      }
      if (i < in_dim) y0 = fmaf(Zs[q * (in_dim + 1) + i], Ws[f * (in_dim + 1) + i], y0);
      const float y = y0 + y1;
      const float go = __ldg(A.gout + (size_t)n * out_w + (A.concat ? h * F : 0) + f) * ginv;
      const float g = y > 0.f ? go : go * expf(y);                                 // ELU'(y) = exp(y) for y <= 0
      A.gy[((size_t)n * H + h) * F + f] = g;
      Gy[q * (F + 1) + f] = g;
    }
    __syncthreads();
    // g_z: thread per (node, i)
    for (int idx = threadIdx.x; idx < tn * in_dim; idx += blockDim.x) {
      const int q = idx / in_dim, i = idx - q * in_dim;
      float g0 = 0.f, g1 = 0.f;
      int f = 0;
      for (; f + 1 < F; f += 2) {
        g0 = fmaf(Gy[q * (F + 1) + f], Ws[f * (in_dim + 1) + i], g0);
        g1 = fmaf(Gy[q * (F + 1) + f + 1], Ws[(f + 1) * (in_dim + 1) + i], g1);
      }
      if (f < F) g0 = fmaf(Gy[q * (F + 1) + f], Ws[f * (in_dim + 1) + i], g0);
      const float g = g0 + g1;
      A.gz[((size_t)(nb + q) * H + h) * in_dim + i] = g;
      Zs[q * (in_dim + 1) + i] *= g;                                               // z_i * g_z_i, reduced below
    }
    __syncthreads();
    // c = sum_i z_i g_z_i : one warp per node (fixed order)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = warp; q < tn; q += blockDim.x >> 5) {
      float acc = 0.f;
      for (int i = lane; i < in_dim; i += 32) acc += Zs[q * (in_dim + 1) + i];
      acc = warp_sum(acc);
      if (lane == 0) A.c[(size_t)(nb + q) * H + h] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient partials: partW[split][h][f][i] = sum_{n in chunk} g_y[n][h][f] z[n][h][i]
// grid (F-tiles * in-tiles, heads, splits); 16x16 outputs per block... one thread per output (f, i)
// ---------------------------------------------------------------------------------------------
constexpr int kWT = 16;
__global__ void __launch_bounds__(kWT* kWT) gat_bwd_weight_kernel(const BwdArgs A) {
  __shared__ float Gs[32][kWT + 1];
  __shared__ float Zs[32][kWT + 1];
  const int in_tiles = ceil_div(A.in_dim, kWT);
  const int ft = blockIdx.x / in_tiles, it = blockIdx.x - ft * in_tiles;
  const int h = blockIdx.y, sp = blockIdx.z;
  const int tf = threadIdx.x / kWT, ti = threadIdx.x % kWT;
  const int f = ft * kWT + tf, i = it * kWT + ti;
  const int chunk = ceil_div(A.N, A.splits);
  const int nbeg = sp * chunk, nend = min(A.N, nbeg + chunk);
  float acc = 0.f;
  for (int n0 = nbeg; n0 < nend; n0 += 32) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * kWT; idx += blockDim.x) {
      const int q = idx / kWT, k = idx % kWT;
      const int n = n0 + q;
      const bool ok = n < nend;
      Gs[q][k] = (ok && ft * kWT + k < A.F) ? __ldg(A.gy + ((size_t)n * A.heads + h) * A.F + ft * kWT + k) : 0.f;
      Zs[q][k] = (ok && it * kWT + k < A.in_dim) ? __ldg(A.z + ((size_t)n * A.heads + h) * A.in_dim + it * kWT + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 32; ++q) acc = fmaf(Gs[q][tf], Zs[q][ti], acc);
  }
  if (f < A.F && i < A.in_dim) A.partW[(((size_t)sp * A.heads + h) * A.F + f) * A.in_dim + i] = acc;
}

// ---------------------------------------------------------------------------------------------
// gradient of the per-graph softmax shift: per-node terms, then one block per graph (fixed order)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gat_bwd_gm_node_kernel(const BwdArgs A) {
  const int H = A.heads, twoH = 2 * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < A.N * H; idx += gridDim.x * blockDim.x) {
    const int j = idx / H, h = idx - j * H;
    const int g = A.nodes_per_graph > 0 ? j / A.nodes_per_graph : 0;
    const float raw = __ldg(A.gmax + (size_t)g * H + h);
    const float stgt = __ldg(A.s + (size_t)j * twoH + H + h);
    const int beg = __ldg(A.rowptr_in + j), end = __ldg(A.rowptr_in + j + 1);
    float cnt = 0.f;
    for (int k = beg; k < end; ++k) cnt += (__ldg(A.s + (size_t)__ldg(A.col_in + k) * twoH + h) + stgt == raw) ? 1.f : 0.f;
    const float dn = __ldg(A.den + idx) + 1e-10f;
    A.gmpart[(size_t)j * twoH + h] = end > beg ? (1e-10f / dn) * __ldg(A.c + idx) : 0.f;
    A.gmpart[(size_t)j * twoH + H + h] = cnt;
  }
}
__global__ void __launch_bounds__(256) gat_bwd_gm_reduce_kernel(const BwdArgs A, int nodes_per_graph) {
  __shared__ float red[8][16];
  const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int twoH = 2 * A.heads;
  const float* base = A.gmpart + (size_t)g * nodes_per_graph * twoH;
  for (int c = 0; c < twoH; ++c) {
    float acc = 0.f;
    for (int n = threadIdx.x; n < nodes_per_graph; n += blockDim.x) acc += __ldg(base + (size_t)n * twoH + c);
    acc = warp_sum(acc);
    if (lane == 0) red[warp][c] = acc;
  }
  __syncthreads();
  if (threadIdx.x < twoH) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    A.gM[(size_t)g * twoH + threadIdx.x] = threadIdx.x < A.heads ? -t : t;
  }
}

// ---------------------------------------------------------------------------------------------
// B2: by target.  warp per node j; lanes over the input dimension
// ---------------------------------------------------------------------------------------------
template <typename TX>
__global__ void __launch_bounds__(256) gat_bwd_target_kernel(const BwdArgs A) {
  const TX* __restrict__ x = reinterpret_cast<const TX*>(A.x);
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int H = A.heads, in_dim = A.in_dim, twoH = 2 * H;
  for (int j = wg; j < A.N; j += nw) {
    const int beg = __ldg(A.rowptr_in + j), end = __ldg(A.rowptr_in + j + 1);
    const int g = A.nodes_per_graph > 0 ? j / A.nodes_per_graph : 0;
    for (int h = 0; h < H; ++h) {
      const float stgt = __ldg(A.s + (size_t)j * twoH + H + h);
      const float raw = __ldg(A.gmax + (size_t)g * H + h);
      const float M = leaky_relu(raw, A.slope);
      const float gM_share = __ldg(A.gM + (size_t)g * twoH + h) / fmaxf(__ldg(A.gM + (size_t)g * twoH + H + h), 1.f);
      const float dn = __ldg(A.den + (size_t)j * H + h) + 1e-10f;
      const float cj = __ldg(A.c + (size_t)j * H + h);
      const float* gzj = A.gz + ((size_t)j * H + h) * in_dim;
      float gst = 0.f;
      for (int k = beg; k < end; ++k) {
        const int src = __ldg(A.col_in + k);
        float dot = 0.f;
        for (int i = lane; i < in_dim; i += 32) dot = fmaf(__ldg(gzj + i), to_f32<TX>(x[(size_t)src * in_dim + i]), dot);
        dot = warp_sum(dot);
        const float pre = __ldg(A.s + (size_t)src * twoH + h) + stgt;
        const float alpha = expf(leaky_relu(pre, A.slope) - M) / dn;
        const float m = A.dropout_p > 0.f ? dropout_keep_scale(A.seed + (A.seed_dev ? __ldg(A.seed_dev) : 0ull), (unsigned)k, (unsigned)h, A.dropout_p) : 1.f;
        const float ge = alpha * (m * dot - cj) + (pre == raw ? gM_share : 0.f);
        const float gpre = pre > 0.f ? ge : ge * A.slope;
        if (lane == 0) {
          A.ea[(size_t)k * H + h] = m * alpha;
          A.eg[(size_t)k * H + h] = gpre;
        }
        gst += gpre;
      }
      if (lane == 0) A.gs[(size_t)j * twoH + H + h] = gst;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// B3: by source.  warp per node i; g_x_i and g_s_src[i]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gat_bwd_source_kernel(const BwdArgs A) {
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int H = A.heads, in_dim = A.in_dim, twoH = 2 * H;
  for (int i = wg; i < A.N; i += nw) {
    const int beg = __ldg(A.rowptr_out + i), end = __ldg(A.rowptr_out + i + 1);
    // g_s_src[i][h]: lane h sums over the out-edges (fixed order)
    float gsrc = 0.f;
    if (lane < H)
      for (int k = beg; k < end; ++k) gsrc += __ldg(A.eg + (size_t)__ldg(A.slot_out2in + k) * H + lane);
    if (lane < H) A.gs[(size_t)i * twoH + lane] = gsrc;
    float gsrc_h[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) gsrc_h[h] = __shfl_sync(kFull, gsrc, h);          // every lane takes part
    for (int d = lane; d < in_dim; d += 32) {
      float acc = 0.f;
      for (int k = beg; k < end; ++k) {
        const int j = __ldg(A.col_out + k);
        const int t = __ldg(A.slot_out2in + k);
        for (int h = 0; h < H; ++h) acc = fmaf(__ldg(A.ea + (size_t)t * H + h), __ldg(A.gz + ((size_t)j * H + h) * in_dim + d), acc);
      }
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        if (h < H) {
          acc = fmaf(gsrc_h[h], __ldg(A.u + (size_t)h * in_dim + d), acc);
          acc = fmaf(__ldg(A.gs + (size_t)i * twoH + H + h), __ldg(A.u + (size_t)(H + h) * in_dim + d), acc);
        }
      }
      A.gx[(size_t)i * in_dim + d] = acc;
    }
  }
}

// partU[split][q][i] = sum_{n in chunk} gs[n][q] x[n][i]
template <typename TX>
__global__ void __launch_bounds__(256) gat_bwd_u_kernel(const BwdArgs A) {
  const TX* __restrict__ x = reinterpret_cast<const TX*>(A.x);
  const int twoH = 2 * A.heads, in_dim = A.in_dim;
  const int sp = blockIdx.x;
  const int chunk = ceil_div(A.N, A.splits);
  const int nbeg = sp * chunk, nend = min(A.N, nbeg + chunk);
  for (int idx = threadIdx.x; idx < twoH * in_dim; idx += blockDim.x) {
    const int q = idx / in_dim, i = idx - q * in_dim;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;      // independent chains: the loads of four nodes are in flight together
    int n = nbeg;
    for (; n + 3 < nend; n += 4) {
      acc0 = fmaf(__ldg(A.gs + (size_t)n * twoH + q), to_f32<TX>(x[(size_t)n * in_dim + i]), acc0);
      acc1 = fmaf(__ldg(A.gs + (size_t)(n + 1) * twoH + q), to_f32<TX>(x[(size_t)(n + 1) * in_dim + i]), acc1);
      acc2 = fmaf(__ldg(A.gs + (size_t)(n + 2) * twoH + q), to_f32<TX>(x[(size_t)(n + 2) * in_dim + i]), acc2);
      acc3 = fmaf(__ldg(A.gs + (size_t)(n + 3) * twoH + q), to_f32<TX>(x[(size_t)(n + 3) * in_dim + i]), acc3);
    }
    for (; n < nend; ++n) acc0 = fmaf(__ldg(A.gs + (size_t)n * twoH + q), to_f32<TX>(x[(size_t)n * in_dim + i]), acc0);
    A.partU[((size_t)sp * twoH + q) * in_dim + i] = (acc0 + acc1) + (acc2 + acc3);
  }
}

// finalize: g_u = sum_splits partU; gW = sum_splits partW + a_src (x) g_u_src + a_tgt (x) g_u_tgt; ga = [W g_u_src | W g_u_tgt]
__global__ void __launch_bounds__(256) gat_bwd_finalize_kernel(const BwdArgs A, float* __restrict__ gu /* (2H, in) */) {
  const int H = A.heads, F = A.F, in_dim = A.in_dim;
  // phase 1 (grid-stride over 2H*in): g_u
  // the kernel is launched with ONE block so phases can be separated by __syncthreads
  for (int idx = threadIdx.x; idx < 2 * H * in_dim; idx += blockDim.x) {
    float acc = 0.f;
    for (int sp = 0; sp < A.splits; ++sp) acc += A.partU[(size_t)sp * 2 * H * in_dim + idx];
    gu[idx] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < H * 2 * F; idx += blockDim.x) {
    const int h = idx / (2 * F), r = idx - h * 2 * F;
    const int half = r / F, f = r - half * F;
    const float* guq = gu + (size_t)(half * H + h) * in_dim;
    float acc = 0.f;
    for (int i = 0; i < in_dim; ++i) acc = fmaf(__ldg(A.W + ((size_t)h * F + f) * in_dim + i), guq[i], acc);
    A.ga[idx] = acc;
  }
}

__global__ void __launch_bounds__(256) gat_bwd_weight_reduce_kernel(const BwdArgs A, const float* __restrict__ gu) {
  const int H = A.heads, F = A.F, in_dim = A.in_dim;
  const int total = H * F * in_dim;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int i = idx % in_dim, f = (idx / in_dim) % F, h = idx / (in_dim * F);
    float acc = 0.f;
    for (int sp = 0; sp < A.splits; ++sp) acc += A.partW[(size_t)sp * total + idx];
    acc = fmaf(__ldg(A.a + (size_t)h * 2 * F + f), gu[(size_t)h * in_dim + i], acc);
    acc = fmaf(__ldg(A.a + (size_t)h * 2 * F + F + f), gu[(size_t)(H + h) * in_dim + i], acc);
    A.gW[idx] = acc;
  }
}

// slot_out2in[s] = in-CSR slot of the edge sitting at out-CSR slot s
__global__ void edge_slot_inv_kernel(const int32_t* __restrict__ eid_in, int64_t E, int32_t* __restrict__ inv) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < E; t += (int64_t)gridDim.x * blockDim.x) inv[eid_in[t]] = (int32_t)t;
}
__global__ void edge_slot_map_kernel(const int32_t* __restrict__ eid_out, const int32_t* __restrict__ inv, int64_t E,
                                     int32_t* __restrict__ out) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < E; s += (int64_t)gridDim.x * blockDim.x) out[s] = inv[eid_out[s]];
}

struct BwdLayout { size_t s, gmax, u, gy, gz, c, ea, eg, gs, gmpart, gM, partW, partU, gu, total; int splits; };
static BwdLayout bwd_layout(int N, int64_t E, int in_dim, int F, int heads, int G) {
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  BwdLayout L;
  L.splits = std::max(1, std::min(kMaxSplits, ceil_div(N, 64)));      // 64-node chunks: enough blocks for a 4096-node shard
  size_t o = 0;
  L.s = o;     o = al(o + (size_t)N * 2 * heads * 4);
  L.gmax = o;  o = al(o + (size_t)G * heads * 4);
  L.u = o;     o = al(o + (size_t)2 * heads * in_dim * 4);
  L.gy = o;    o = al(o + (size_t)N * heads * F * 4);
  L.gz = o;    o = al(o + (size_t)N * heads * in_dim * 4);
  L.c = o;     o = al(o + (size_t)N * heads * 4);
  L.ea = o;    o = al(o + (size_t)E * heads * 4);
  L.eg = o;    o = al(o + (size_t)E * heads * 4);
  L.gs = o;    o = al(o + (size_t)N * 2 * heads * 4);
  L.gmpart = o; o = al(o + (size_t)N * 2 * heads * 4);
  L.gM = o;    o = al(o + (size_t)G * 2 * heads * 4);
  L.partW = o; o = al(o + (size_t)L.splits * heads * F * in_dim * 4);
  L.partU = o; o = al(o + (size_t)L.splits * 2 * heads * in_dim * 4);
  L.gu = o;    o = al(o + (size_t)2 * heads * in_dim * 4);
  L.total = o;
  return L;
}

}  // namespace mg

using namespace mg;

extern "C" {

int mg_edge_slot_map(const int32_t* eid_in, const int32_t* eid_out, int64_t E, int32_t* work, int32_t* slot_out2in,
                     mg_stream_t stream) {
  MG_REQUIRE(E >= 0 && (E == 0 || (eid_in && eid_out && work && slot_out2in)), MG_ERR_INVALID, "mg_edge_slot_map: bad arguments");
  if (E == 0) return MG_OK;
  const int grid = (int)std::min<int64_t>(ceil_div64(E, 256), (int64_t)num_sms() * 8);
  edge_slot_inv_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(eid_in, E, work);
  int rc;
  if ((rc = check_launch("edge_slot_inv_kernel"))) return rc;
  edge_slot_map_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(eid_out, work, E, slot_out2in);
  return check_launch("edge_slot_map_kernel");
}

int64_t mg_gat_backward_work_bytes(int N, int64_t E, int in_dim, int out_dim, int heads, int num_graphs) {
  if (N <= 0 || in_dim <= 0 || out_dim <= 0 || heads <= 0) return 0;
  return (int64_t)bwd_layout(N, E, in_dim, out_dim, heads, num_graphs > 0 ? num_graphs : 1).total;
}

int mg_gat_backward(const void* x, int x_dtype, const int32_t* rowptr_in, const int32_t* col_in, const int32_t* rowptr_out,
                    const int32_t* col_out, const int32_t* slot_out2in, int N, int64_t E, const float* W, const float* a,
                    int in_dim, int out_dim, int heads, int concat, float slope, int nodes_per_graph, float dropout_p,
                    uint64_t seed, const uint64_t* seed_dev, const float* den, const float* z, const float* grad_out, float* grad_x, float* grad_W,
                    float* grad_a, void* work, const void* fwd_work, mg_stream_t stream) {
  MG_REQUIRE(x && rowptr_in && col_in && rowptr_out && col_out && slot_out2in && W && a && den && z && grad_out && grad_x &&
                 grad_W && grad_a && work,
             MG_ERR_INVALID, "mg_gat_backward: null pointer");
  MG_REQUIRE(N > 0 && E > 0 && in_dim > 0 && out_dim > 0 && heads >= 1 && heads <= 8, MG_ERR_INVALID, "mg_gat_backward: bad sizes");
  MG_REQUIRE(x_dtype == MG_F32 || x_dtype == MG_BF16, MG_ERR_INVALID, "mg_gat_backward: x dtype");
  MG_REQUIRE(nodes_per_graph >= 0 && (nodes_per_graph == 0 || N % nodes_per_graph == 0), MG_ERR_INVALID,
             "mg_gat_backward: N=%d is not a multiple of nodes_per_graph=%d", N, nodes_per_graph);
  const size_t node_smem = ((size_t)out_dim * (in_dim + 1) + (size_t)kBwdTile * (in_dim + 1) + (size_t)kBwdTile * (out_dim + 1)) * 4;
  MG_REQUIRE(node_smem <= 200 * 1024, MG_ERR_UNSUPPORTED,
             "mg_gat_backward: one head's weights (%d x %d) do not fit in shared memory", out_dim, in_dim);
  const int G = nodes_per_graph > 0 ? N / nodes_per_graph : 1;
  const BwdLayout L = bwd_layout(N, E, in_dim, out_dim, heads, G);
  unsigned char* wb = reinterpret_cast<unsigned char*>(work);
  cudaStream_t st = (cudaStream_t)stream;
  BwdArgs A;
  A.x = x; A.W = W; A.a = a; A.den = den; A.z = z; A.gout = grad_out;
  // attention scalars, per-graph maxima and u = W^T a: recomputed here, or taken from the forward's workspace of the same
  // layer call (mg_gat_forward's work buffer starts with the same three segments: saves three launches per layer)
  const unsigned char* sb = fwd_work ? reinterpret_cast<const unsigned char*>(fwd_work) : wb;
  A.u = reinterpret_cast<float*>(const_cast<unsigned char*>(sb) + L.u);
  A.s = reinterpret_cast<float*>(const_cast<unsigned char*>(sb) + L.s);
  A.gmax = reinterpret_cast<float*>(const_cast<unsigned char*>(sb) + L.gmax);
  A.rowptr_in = rowptr_in; A.col_in = col_in; A.rowptr_out = rowptr_out; A.col_out = col_out; A.slot_out2in = slot_out2in;
  A.gy = reinterpret_cast<float*>(wb + L.gy); A.gz = reinterpret_cast<float*>(wb + L.gz);
  A.c = reinterpret_cast<float*>(wb + L.c); A.ea = reinterpret_cast<float*>(wb + L.ea);
  A.eg = reinterpret_cast<float*>(wb + L.eg); A.gs = reinterpret_cast<float*>(wb + L.gs);
  A.gmpart = reinterpret_cast<float*>(wb + L.gmpart); A.gM = reinterpret_cast<float*>(wb + L.gM);
  A.gx = grad_x; A.partW = reinterpret_cast<float*>(wb + L.partW); A.partU = reinterpret_cast<float*>(wb + L.partU);
  A.gW = grad_W; A.ga = grad_a;
  A.N = N; A.in_dim = in_dim; A.F = out_dim; A.heads = heads; A.concat = concat ? 1 : 0;
  A.nodes_per_graph = nodes_per_graph; A.splits = L.splits; A.slope = slope; A.dropout_p = dropout_p; A.seed = seed;
  A.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  float* gu = reinterpret_cast<float*>(wb + L.gu);
  int rc;
  // recompute the attention scalars and the per-graph shift (cheaper than saving them)
  if (!fwd_work &&
      (rc = gat_scores_and_max(x, x_dtype, rowptr_in, col_in, N, W, a, in_dim, out_dim, heads, nodes_per_graph,
                               const_cast<float*>(A.s), const_cast<float*>(A.gmax), const_cast<float*>(A.u), st)))
    return rc;
  if (node_smem > 48 * 1024) cudaFuncSetAttribute(gat_bwd_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const dim3 node_grid((unsigned)std::min(ceil_div(N, kBwdTile), num_sms() * 2), (unsigned)heads);
  gat_bwd_node_kernel<<<node_grid, 256, node_smem, st>>>(A);
  if ((rc = check_launch("gat_bwd_node_kernel"))) return rc;
  dim3 wgrid(ceil_div(out_dim, kWT) * ceil_div(in_dim, kWT), heads, L.splits);
  gat_bwd_weight_kernel<<<wgrid, kWT * kWT, 0, st>>>(A);
  if ((rc = check_launch("gat_bwd_weight_kernel"))) return rc;
  gat_bwd_gm_node_kernel<<<std::min(ceil_div(N * heads, 256), num_sms() * 8), 256, 0, st>>>(A);
  if ((rc = check_launch("gat_bwd_gm_node_kernel"))) return rc;
  gat_bwd_gm_reduce_kernel<<<G, 256, 0, st>>>(A, nodes_per_graph > 0 ? nodes_per_graph : N);
  if ((rc = check_launch("gat_bwd_gm_reduce_kernel"))) return rc;
  const int warp_grid = (int)std::min<int64_t>(ceil_div64((int64_t)N * 32, 256), (int64_t)num_sms() * 8);
  if (x_dtype == MG_F32) gat_bwd_target_kernel<float><<<warp_grid, 256, 0, st>>>(A);
  else gat_bwd_target_kernel<__nv_bfloat16><<<warp_grid, 256, 0, st>>>(A);
  if ((rc = check_launch("gat_bwd_target_kernel"))) return rc;
  gat_bwd_source_kernel<<<warp_grid, 256, 0, st>>>(A);
  if ((rc = check_launch("gat_bwd_source_kernel"))) return rc;
  if (x_dtype == MG_F32) gat_bwd_u_kernel<float><<<L.splits, 256, 0, st>>>(A);
  else gat_bwd_u_kernel<__nv_bfloat16><<<L.splits, 256, 0, st>>>(A);
  if ((rc = check_launch("gat_bwd_u_kernel"))) return rc;
  gat_bwd_finalize_kernel<<<1, 256, 0, st>>>(A, gu);
  if ((rc = check_launch("gat_bwd_finalize_kernel"))) return rc;
  const int rgrid = std::min(ceil_div(heads * out_dim * in_dim, 256), num_sms() * 4);
  gat_bwd_weight_reduce_kernel<<<rgrid, 256, 0, st>>>(A, gu);
  return check_launch("gat_bwd_weight_reduce_kernel");
}

}  // extern "C"
