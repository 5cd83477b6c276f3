// kNN graph build over node features (north_star kernel (2)).
//
// NOT part of the reference: MinGraph-UNet builds its graphs from grid adjacency
// (preprocessing/graph_construction/patch_graph_construction.py:78-97) and has no knn/cdist/topk code; the nearest
// reference arithmetic is the per-edge squared distance of model/graph_partition/mincut_refinement.py:43-46.  Parity is
// therefore pinned by the repo's own oracle only (oracle/restate.py::knn_graph), against which neighbour sets are
// BIT-EXACT: distances are accumulated in fp32 in feature order with separately rounded subtract / multiply / add
// (no FMA contraction), exactly like the oracle's numpy loop; ties go to the lower node id; self is excluded.
//
// Kernel: FP32 tiled pairwise distances feeding a warp-level top-k.
//   * block = 8 warps x 4 queries; candidate tiles of 128 nodes x 64 features are staged TRANSPOSED in shared memory
//     (conflict-free lane-per-candidate reads), query slices are broadcast reads;
//   * each lane accumulates 4 queries x 4 candidates per tile (16 independent sub/mul/add chains);
//   * the k best (distance, id) pairs of a query live one per lane, sorted; a candidate that beats the current k-th
//     is inserted with one ballot + shuffle shift.
// Graphs are independent per image: with nodes_per_graph > 0 a query only sees the nodes of its own graph.
#include "common.cuh"

namespace mg {

constexpr int kKnnWarps = 8;
constexpr int kKnnQW = 4;                       // queries per warp
constexpr int kKnnTQ = kKnnWarps * kKnnQW;      // queries per block
constexpr int kKnnTC = 128;                     // candidates per tile
constexpr int kKnnDC = 64;                      // features per staged slice
constexpr int kKnnCL = kKnnTC / 32;             // candidates per lane per tile

__device__ __forceinline__ bool knn_less(float d0, int i0, float d1, int i1) { return d0 < d1 || (d0 == d1 && i0 < i1); }

__global__ void __launch_bounds__(kKnnWarps * 32) knn_kernel(const float* __restrict__ x, int N, int D, int k,
                                                             int nodes_per_graph, int64_t* __restrict__ edge_index,
                                                             int32_t* __restrict__ col, float* __restrict__ dist_out) {
  __shared__ float xT[kKnnDC][kKnnTC + 1];      // candidate slice, transposed
  __shared__ float xq[kKnnTQ][kKnnDC];          // query slice
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npg = nodes_per_graph > 0 ? nodes_per_graph : N;
  const int qblocks_per_graph = ceil_div(npg, kKnnTQ);
  const int graph = blockIdx.x / qblocks_per_graph;
  const int g0 = graph * npg;                                   // first node of this graph
  const int q0 = g0 + (blockIdx.x - graph * qblocks_per_graph) * kKnnTQ;
  const int g1 = min(g0 + npg, N);

  // the k best of each of this warp's queries: lane i holds the i-th best (sorted ascending by (distance, id))
  float bd[kKnnQW];
  int bi[kKnnQW];
#pragma unroll
  for (int q = 0; q < kKnnQW; ++q) { bd[q] = INFINITY; bi[q] = 0x7fffffff; }

  for (int c0 = g0; c0 < g1; c0 += kKnnTC) {
    float acc[kKnnQW][kKnnCL];
#pragma unroll
    for (int q = 0; q < kKnnQW; ++q)
#pragma unroll
      for (int m = 0; m < kKnnCL; ++m) acc[q][m] = 0.f;
    for (int d0 = 0; d0 < D; d0 += kKnnDC) {
      const int dn = min(kKnnDC, D - d0);
      __syncthreads();
      // stage: candidate rows (coalesced along d) -> xT[d][c]; query rows -> xq[q][d]
      for (int idx = threadIdx.x; idx < kKnnTC * kKnnDC; idx += blockDim.x) {
        const int c = idx / kKnnDC, d = idx - c * kKnnDC;
        float v = 0.f;
        if (c0 + c < g1 && d < dn) v = __ldg(x + (size_t)(c0 + c) * D + d0 + d);
        xT[d][c] = v;
      }
      for (int idx = threadIdx.x; idx < kKnnTQ * kKnnDC; idx += blockDim.x) {
        const int q = idx / kKnnDC, d = idx - q * kKnnDC;
        float v = 0.f;
        if (q0 + q < g1 && d < dn) v = __ldg(x + (size_t)(q0 + q) * D + d0 + d);
        xq[q][d] = v;
      }
      __syncthreads();
      // acc += (xq - xc)^2 in feature order; rounded sub, mul, add (must not contract into FMA: oracle parity)
      for (int d = 0; d < dn; ++d) {
        float cv[kKnnCL];
#pragma unroll
        for (int m = 0; m < kKnnCL; ++m) cv[m] = xT[d][lane + 32 * m];
#pragma unroll
        for (int q = 0; q < kKnnQW; ++q) {
          const float qv = xq[warp * kKnnQW + q][d];
#pragma unroll
          for (int m = 0; m < kKnnCL; ++m) {
            const float t = __fsub_rn(qv, cv[m]);
            acc[q][m] = __fadd_rn(acc[q][m], __fmul_rn(t, t));
          }
        }
      }
    }
    // top-k update: candidates of this tile, per query
#pragma unroll
    for (int q = 0; q < kKnnQW; ++q) {
      const int qn = q0 + warp * kKnnQW + q;
#pragma unroll
      for (int m = 0; m < kKnnCL; ++m) {
        const int cn = c0 + lane + 32 * m;
        const bool cand_ok = cn < g1 && cn != qn && qn < g1;
        // current k-th best (lane k-1)
        const float kd = __shfl_sync(kFull, bd[q], k - 1);
        const int ki = __shfl_sync(kFull, bi[q], k - 1);
        unsigned todo = __ballot_sync(kFull, cand_ok && knn_less(acc[q][m], cn, kd, ki));
        while (todo) {
          const int src_lane = __ffs(todo) - 1;
          todo &= todo - 1;
          const float d = __shfl_sync(kFull, acc[q][m], src_lane);
          const int id = __shfl_sync(kFull, cn, src_lane);
          // re-check against the (possibly improved) k-th best
          const float kd2 = __shfl_sync(kFull, bd[q], k - 1);
          const int ki2 = __shfl_sync(kFull, bi[q], k - 1);
          if (!knn_less(d, id, kd2, ki2)) continue;              // warp-uniform
          const bool before = knn_less(d, id, bd[q], bi[q]);      // candidate sorts before this lane's element
          const float up_d = __shfl_up_sync(kFull, bd[q], 1);
          const int up_i = __shfl_up_sync(kFull, bi[q], 1);
          const bool prev_before = __shfl_up_sync(kFull, (int)before, 1) != 0 && lane > 0;
          if (before) {
            if (prev_before) { bd[q] = up_d; bi[q] = up_i; }       // shift right
            else { bd[q] = d; bi[q] = id; }                        // insertion point
          }
        }
      }
    }
  }
  // write: lane i < k holds neighbour i of each query
#pragma unroll
  for (int q = 0; q < kKnnQW; ++q) {
    const int qn = q0 + warp * kKnnQW + q;
    if (qn < g1 && lane < k) {
      const size_t o = (size_t)qn * k + lane;
      if (col) col[o] = bi[q];
      if (dist_out) dist_out[o] = bd[q];
      if (edge_index) {
        edge_index[o] = (int64_t)bi[q];                            // row 0: source (the neighbour)
        edge_index[(size_t)N * k + o] = (int64_t)qn;              // row 1: target (the query node)
      }
    }
  }
}

__global__ void knn_rowptr_kernel(int N, int k, int32_t* __restrict__ rowptr) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= N; i += gridDim.x * blockDim.x) rowptr[i] = i * k;
}

}  // namespace mg

using namespace mg;

extern "C" {

int mg_knn_graph(const float* x, int N, int D, int k, int nodes_per_graph, int64_t* edge_index, int32_t* rowptr, int32_t* col,
                 float* dist, mg_stream_t stream) {
  MG_REQUIRE(x && N > 0 && D > 0, MG_ERR_INVALID, "mg_knn_graph: bad arguments");
  MG_REQUIRE(k >= 1 && k <= 32, MG_ERR_UNSUPPORTED, "mg_knn_graph: k=%d (supported 1..32)", k);
  MG_REQUIRE(nodes_per_graph >= 0 && (nodes_per_graph == 0 || N % nodes_per_graph == 0), MG_ERR_INVALID,
             "mg_knn_graph: N=%d is not a multiple of nodes_per_graph=%d", N, nodes_per_graph);
  const int npg = nodes_per_graph > 0 ? nodes_per_graph : N;
  MG_REQUIRE(npg > k, MG_ERR_INVALID, "mg_knn_graph: a graph of %d nodes has fewer than k=%d other nodes", npg, k);
  MG_REQUIRE((int64_t)N * k < (int64_t)1 << 31, MG_ERR_UNSUPPORTED, "mg_knn_graph: N*k overflows int32 CSR");
  cudaStream_t st = (cudaStream_t)stream;
  const int G = N / npg;
  const int grid = G * ceil_div(npg, kKnnTQ);
  knn_kernel<<<grid, kKnnWarps * 32, 0, st>>>(x, N, D, k, nodes_per_graph, edge_index, col, dist);
  int rc = check_launch("knn_kernel");
  if (rc) return rc;
  if (rowptr) {
    knn_rowptr_kernel<<<std::min(ceil_div(N + 1, 256), num_sms() * 4), 256, 0, st>>>(N, k, rowptr);
    rc = check_launch("knn_rowptr_kernel");
  }
  return rc;
}

}  // extern "C"
