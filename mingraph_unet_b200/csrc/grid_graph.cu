// K2 — graph construction kernels.
//   * closed-form 4-connected patch grid: COO edge_index in the reference's emission order
//     (preprocessing/graph_construction/patch_graph_construction.py:78-97) and its CSR;
//   * complete region digraph (scripts/train_end_to_end.py:376-380);
//   * stable COO -> CSR for caller-supplied edge_index tensors.
// All integer work, bit-exact by construction; traffic is 16 B/edge (COO) or 4 B/edge (CSR).
#include "common.cuh"

namespace mg {

struct GridPos {
  int r, c, n;
  bool up, left, right, down;
  int off;      // COO slot of this node's first emitted edge
};

__device__ __forceinline__ int grid_off(int r, int c, int Hp, int Wp) {
  return r * (4 * Wp - 2) + (r < Hp - 1 ? 4 * c : 2 * c);
}

__device__ __forceinline__ GridPos grid_pos(int n, int Hp, int Wp) {
  GridPos p;
  p.n = n;
  p.r = n / Wp;
  p.c = n - p.r * Wp;
  p.up = p.r > 0;
  p.left = p.c > 0;
  p.right = p.c + 1 < Wp;
  p.down = p.r + 1 < Hp;
  p.off = grid_off(p.r, p.c, Hp, Wp);
  return p;
}

// one thread per (image, node): writes the <=4 edges that node emits
__global__ void grid_coo_kernel(int Hp, int Wp, int B, int offset_nodes, int64_t* __restrict__ ei) {
  const int N = Hp * Wp;
  const int64_t E = 2LL * (Hp * (Wp - 1) + Wp * (Hp - 1));
  const int64_t total = (int64_t)B * N;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / N);
    const GridPos p = grid_pos((int)(t - (int64_t)b * N), Hp, Wp);
    const int64_t base = offset_nodes ? (int64_t)b * N : 0;
    int64_t* src = ei + (int64_t)b * E + p.off;
    int64_t* tgt = src + (int64_t)B * E;
    const int64_t me = base + p.n;
    int k = 0;
    if (p.right) {
      src[k] = me;         tgt[k] = me + 1;  ++k;
      src[k] = me + 1;     tgt[k] = me;      ++k;
    }
    if (p.down) {
      src[k] = me;         tgt[k] = me + Wp; ++k;
      src[k] = me + Wp;    tgt[k] = me;      ++k;
    }
  }
}

// one thread per (image, node): closed-form rowptr + neighbour list (up,left,right,down)
__global__ void grid_csr_kernel(int Hp, int Wp, int B, int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                int32_t* __restrict__ eid_in, int32_t* __restrict__ eid_out) {
  const int N = Hp * Wp;
  const int E = 2 * (Hp * (Wp - 1) + Wp * (Hp - 1));
  const int64_t total = (int64_t)B * N;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / N);
    const GridPos p = grid_pos((int)(t - (int64_t)b * N), Hp, Wp);
    // in-degree prefix: full rows above, then the columns to the left in this row
    const int rows_up = p.r > 0 ? p.r - 1 : 0;                       // rows r'<r that have an up neighbour
    const int rows_dn = p.r < Hp - 1 ? p.r : Hp - 1;                 // rows r'<r that have a down neighbour
    int pre = Wp * (rows_up + rows_dn) + p.r * 2 * (Wp - 1);
    pre += p.c * ((p.up ? 1 : 0) + (p.down ? 1 : 0)) + (p.c > 0 ? p.c - 1 : 0) + (p.c < Wp - 1 ? p.c : Wp - 1);
    const int base_e = b * E;
    const int base_n = b * N;
    int k = base_e + pre;
    rowptr[base_n + p.n] = k;
    if (p.up) {
      const int q = p.n - Wp;                                         // emitted at node q: (q->q+Wp),(q+Wp->q)
      const int o = grid_off(p.r - 1, p.c, Hp, Wp) + (p.right ? 2 : 0);
      col[k] = base_n + q;
      if (eid_in) eid_in[k] = o;
      if (eid_out) eid_out[k] = o + 1;
      ++k;
    }
    if (p.left) {
      const int o = grid_off(p.r, p.c - 1, Hp, Wp);                   // node n-1: (n-1->n),(n->n-1)
      col[k] = base_n + p.n - 1;
      if (eid_in) eid_in[k] = o;
      if (eid_out) eid_out[k] = o + 1;
      ++k;
    }
    if (p.right) {
      col[k] = base_n + p.n + 1;
      if (eid_in) eid_in[k] = p.off + 1;
      if (eid_out) eid_out[k] = p.off;
      ++k;
    }
    if (p.down) {
      const int o = p.off + (p.right ? 2 : 0);
      col[k] = base_n + p.n + Wp;
      if (eid_in) eid_in[k] = o + 1;
      if (eid_out) eid_out[k] = o;
      ++k;
    }
    if (t == total - 1) rowptr[total] = B * E;
  }
}

__global__ void complete_coo_kernel(int K, int B, int offset_nodes, int64_t* __restrict__ ei) {
  const int64_t half = (int64_t)K * (K - 1) / 2;
  const int64_t E = 2 * half;
  const int64_t total = (int64_t)B * half;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / half);
    int64_t q = t - (int64_t)b * half;            // index into the row-major upper triangle
    int s = 0;
    while (q >= K - 1 - s) { q -= K - 1 - s; ++s; }
    const int tt = s + 1 + (int)q;
    const int64_t base = offset_nodes ? (int64_t)b * K : 0;
    const int64_t idx = t - (int64_t)b * half;
    int64_t* src = ei + (int64_t)b * E;
    int64_t* tgt = src + (int64_t)B * E;
    src[idx] = base + s;          tgt[idx] = base + tt;
    src[half + idx] = base + tt;  tgt[half + idx] = base + s;
  }
}

__global__ void complete_csr_kernel(int K, int B, int32_t* __restrict__ rowptr, int32_t* __restrict__ col) {
  const int total = B * K;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int b = t / K, j = t - b * K;
    int k = t * (K - 1);
    rowptr[t] = k;
    for (int s = 0; s < K; ++s)
      if (s != j) col[k++] = b * K + s;
    if (t == total - 1) rowptr[total] = total * (K - 1);
  }
}

// ---------------- generic stable COO -> CSR -------------------------------------------------
__global__ void csr_zero_kernel(int32_t* cnt, int n, int32_t* status) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) cnt[i] = 0;
  if (status && blockIdx.x == 0 && threadIdx.x == 0) *status = 0;
}

__global__ void csr_count_kernel(const int64_t* __restrict__ key, const int64_t* __restrict__ other, int64_t E, int N,
                                 int32_t* __restrict__ cnt, int32_t* status) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = key[e], o = other[e];
    if (k < 0 || k >= N || o < 0 || o >= N) {
      if (status) atomicExch(status, 1);
      continue;
    }
    atomicAdd(&cnt[k], 1);
  }
}

constexpr int kScanItems = 4;
constexpr int kScanThreads = 1024;
constexpr int kScanTile = kScanItems * kScanThreads;

__device__ int block_exclusive_scan(int v, int* total) {
  __shared__ int ws[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = ws[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(kFull, s, o);
      if (lane >= o) s += t;
    }
    ws[lane] = s;
  }
  __syncthreads();
  const int base = w > 0 ? ws[w - 1] : 0;
  if (total) *total = ws[31];
  __syncthreads();
  return base + inc - v;
}

// phase 1: tile-local exclusive scan in place, tile totals to sums[]
__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(int32_t* data, int n, int32_t* sums) {
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int v[kScanItems], s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = base + i < n ? data[base + i] : 0;
    s += v[i];
  }
  int total;
  int ex = block_exclusive_scan(s, &total);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) data[base + i] = ex;
    ex += v[i];
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
// phase 2: one block scans the tile totals (<= kScanTile tiles)
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(int32_t* sums, int nt) {
  const int base = threadIdx.x * kScanItems;
  int v[kScanItems], s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = base + i < nt ? sums[base + i] : 0;
    s += v[i];
  }
  int ex = block_exclusive_scan(s, nullptr);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < nt) sums[base + i] = ex;
    ex += v[i];
  }
}
// phase 3: add tile offsets; also seed the fill cursors
__global__ void scan_add_kernel(int32_t* data, int n, const int32_t* sums, int32_t* rowptr, int32_t* cursor) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int v = data[i] + sums[i / kScanTile];
    rowptr[i] = v;
    cursor[i] = v;
  }
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ key, const int64_t* __restrict__ other, int64_t E, int N,
                                int32_t* __restrict__ cursor, int32_t* __restrict__ eid_tmp) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = key[e], o = other[e];
    if (k < 0 || k >= N || o < 0 || o >= N) continue;
    eid_tmp[atomicAdd(&cursor[k], 1)] = (int)e;
  }
}

// warp per row: rank-sort the (distinct) edge ids of the row => ascending COO order (stable CSR)
__global__ void csr_sort_rows_kernel(const int32_t* __restrict__ rowptr, int N, const int32_t* __restrict__ eid_tmp,
                                     const int64_t* __restrict__ other, int32_t* __restrict__ col,
                                     int32_t* __restrict__ eid) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < N; row += warps) {
    const int b = rowptr[row], d = rowptr[row + 1] - b;
    if (d <= 32) {
      const int mine = lane < d ? eid_tmp[b + lane] : 0x7fffffff;
      int rank = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int o = __shfl_sync(kFull, mine, j);
        rank += (o < mine) ? 1 : 0;
      }
      if (lane < d) {
        col[b + rank] = (int)other[mine];
        if (eid) eid[b + rank] = mine;
      }
    } else {
      for (int i = lane; i < d; i += 32) {
        const int mine = eid_tmp[b + i];
        int rank = 0;
        for (int j = 0; j < d; ++j) rank += (eid_tmp[b + j] < mine) ? 1 : 0;
        col[b + rank] = (int)other[mine];
        if (eid) eid[b + rank] = mine;
      }
    }
  }
}

static inline int grid_for(int64_t work, int block, int cap_waves = 8) {
  int64_t g = ceil_div64(work, block);
  int64_t cap = (int64_t)num_sms() * cap_waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace mg

using namespace mg;

extern "C" {

int64_t mg_grid_num_edges(int Hp, int Wp) {
  if (Hp <= 0 || Wp <= 0) return 0;
  return 2LL * ((int64_t)Hp * (Wp - 1) + (int64_t)Wp * (Hp - 1));
}

int mg_grid_edge_index(int Hp, int Wp, int B, int offset_nodes, int64_t* ei, mg_stream_t stream) {
  MG_REQUIRE(Hp > 0 && Wp > 0 && B > 0, MG_ERR_INVALID, "mg_grid_edge_index: bad grid %dx%d B=%d", Hp, Wp, B);
  MG_REQUIRE((int64_t)B * mg_grid_num_edges(Hp, Wp) < (1LL << 31), MG_ERR_INVALID, "mg_grid_edge_index: too many edges");
  if (mg_grid_num_edges(Hp, Wp) == 0) return MG_OK;
  MG_REQUIRE(ei != nullptr, MG_ERR_INVALID, "mg_grid_edge_index: null output");
  const int64_t total = (int64_t)B * Hp * Wp;
  grid_coo_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(Hp, Wp, B, offset_nodes, ei);
  return check_launch("grid_coo_kernel");
}

int mg_grid_csr(int Hp, int Wp, int B, int32_t* rowptr, int32_t* col, int32_t* eid_in, int32_t* eid_out,
                mg_stream_t stream) {
  MG_REQUIRE(Hp > 0 && Wp > 0 && B > 0 && rowptr, MG_ERR_INVALID, "mg_grid_csr: bad arguments");
  MG_REQUIRE((int64_t)B * mg_grid_num_edges(Hp, Wp) < (1LL << 31) && (int64_t)B * Hp * Wp < (1LL << 31), MG_ERR_INVALID,
             "mg_grid_csr: graph too large for int32 CSR");
  const int64_t total = (int64_t)B * Hp * Wp;
  grid_csr_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(Hp, Wp, B, rowptr, col, eid_in, eid_out);
  return check_launch("grid_csr_kernel");
}

int mg_complete_edge_index(int K, int B, int offset_nodes, int64_t* ei, mg_stream_t stream) {
  MG_REQUIRE(K > 0 && B > 0, MG_ERR_INVALID, "mg_complete_edge_index: bad K=%d B=%d", K, B);
  if (K == 1) return MG_OK;
  const int64_t total = (int64_t)B * K * (K - 1) / 2;
  complete_coo_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(K, B, offset_nodes, ei);
  return check_launch("complete_coo_kernel");
}

int mg_complete_csr(int K, int B, int32_t* rowptr, int32_t* col, mg_stream_t stream) {
  MG_REQUIRE(K > 0 && B > 0 && rowptr, MG_ERR_INVALID, "mg_complete_csr: bad arguments");
  complete_csr_kernel<<<grid_for((int64_t)B * K, 256), 256, 0, (cudaStream_t)stream>>>(K, B, rowptr, col);
  return check_launch("complete_csr_kernel");
}

int64_t mg_csr_work_bytes(int N, int64_t E) {
  // counts/cursor (N+1) + tile sums (kScanTile) + unsorted edge ids (E)
  return 4 * ((int64_t)(N + 1) * 2 + kScanTile + E) + 256;
}

int mg_csr_from_coo(const int64_t* ei, int64_t E, int N, int by_target, int32_t* rowptr, int32_t* col, int32_t* eid,
                    void* work, int32_t* status, mg_stream_t stream) {
  MG_REQUIRE(N > 0 && E >= 0 && rowptr && work, MG_ERR_INVALID, "mg_csr_from_coo: bad arguments");
  MG_REQUIRE(E < (1LL << 31), MG_ERR_INVALID, "mg_csr_from_coo: E too large for int32 CSR");
  const int n1 = N + 1;
  const int tiles = ceil_div(n1, kScanTile);
  MG_REQUIRE(tiles <= kScanTile, MG_ERR_INVALID, "mg_csr_from_coo: N too large");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* cnt = reinterpret_cast<int32_t*>(work);
  int32_t* cursor = cnt + n1;
  int32_t* sums = cursor + n1;
  int32_t* eid_tmp = sums + kScanTile;
  const int64_t* key = by_target ? ei + E : ei;
  const int64_t* other = by_target ? ei : ei + E;
  int rc;
  csr_zero_kernel<<<grid_for(n1, 256), 256, 0, st>>>(cnt, n1, status);
  if ((rc = check_launch("csr_zero_kernel"))) return rc;
  if (E > 0) {
    csr_count_kernel<<<grid_for(E, 256), 256, 0, st>>>(key, other, E, N, cnt, status);
    if ((rc = check_launch("csr_count_kernel"))) return rc;
  }
  scan_tiles_kernel<<<tiles, kScanThreads, 0, st>>>(cnt, n1, sums);
  if ((rc = check_launch("scan_tiles_kernel"))) return rc;
  scan_sums_kernel<<<1, kScanThreads, 0, st>>>(sums, tiles);
  if ((rc = check_launch("scan_sums_kernel"))) return rc;
  scan_add_kernel<<<grid_for(n1, 256), 256, 0, st>>>(cnt, n1, sums, rowptr, cursor);
  if ((rc = check_launch("scan_add_kernel"))) return rc;
  if (E > 0) {
    csr_fill_kernel<<<grid_for(E, 256), 256, 0, st>>>(key, other, E, N, cursor, eid_tmp);
    if ((rc = check_launch("csr_fill_kernel"))) return rc;
    csr_sort_rows_kernel<<<grid_for((int64_t)N * 32, 256), 256, 0, st>>>(rowptr, N, eid_tmp, other, col, eid);
    if ((rc = check_launch("csr_sort_rows_kernel"))) return rc;
  }
  return MG_OK;
}

}  // extern "C"
