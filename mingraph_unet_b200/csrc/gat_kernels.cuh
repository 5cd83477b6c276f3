// Templates of the fused / unfused GAT aggregation kernels.  Instantiated per input dtype in
// gat_inst_f32.cu / gat_inst_bf16.cu so the translation units compile in parallel.
#pragma once
#include "gat_device.cuh"

namespace mg {

struct DimCfg { int V, T; };

int gat_scores_and_max(const void* x, int x_dtype, const int32_t* rowptr, const int32_t* col, int N, const float* W,
                       const float* a, int in_dim, int out_dim, int heads, int nodes_per_graph, float* s, float* gmax,
                       float* u, cudaStream_t st, int64_t E = -1);      // E >= 0: the caller knows the edge count (faster edge-max kernel)
static constexpr int64_t kSmemBudget = 200 * 1024;

// tensor-pipe node transform for spilled z (gat_tc_gemm.cu): passes = 1 (tf32, bf16-storage path) or 3 (3xTF32, fp32 path)
bool gat_transform_tc_supported(int N, int in_dim, int F, int heads, int passes);
int gat_transform_tc_launch(const float* z, const float* W, int N, int in_dim, int F, int heads, int concat, void* out,
                            int out_bf16, int passes, cudaStream_t st);

// bf16 score pre-pass on mma.sync (gat_tc.cu): u, s and the exact per-graph edge maximum for heads 1/2/4, in 32..256
bool gat_tc_prepass_supported(int N, int in_dim, int heads);
int gat_tc_prepass(const void* x, const int32_t* rowptr, const int32_t* col, int N, int64_t E, const float* W, const float* a, int in_dim,
                   int F, int heads, int nodes_per_graph, float* s, float* gmax, float* u, cudaStream_t st);
// exact per-graph edge maximum alone (gat_tc.cu), for heads 1/2/4 behind any score kernel that reset gmax
bool gat_tc_edge_max_supported(int heads);
int gat_tc_edge_max(const int32_t* rowptr, const int32_t* col, const float* s, int N, int64_t E, int heads, int nodes_per_graph, float* gmax,
                    cudaStream_t st);
// tensor-core aggregation with z spilled to HBM as bf16 (gat_tc.cu): heads 4, in a multiple of 64; needs s / gmax of the pre-pass
bool gat_agg_spill_supported(int N, int in_dim, int heads);
int gat_agg_spill_launch(const void* x, const int32_t* rowptr, const int32_t* col, const float* s, const float* gmax, void* z_bf16, int N,
                         int in_dim, float slope, int nodes_per_graph, cudaStream_t st);
// persistent TMA-fed bf16 tensor-pipe transform (gat_tma_gemm.cu): z spilled as bf16, W converted to bf16 into w_bf16
bool gat_transform_tma_supported(int N, int in_dim, int F, int heads);
int64_t gat_transform_tma_wbytes(int in_dim, int F, int heads);
int gat_transform_tma_launch(const void* z_bf16, const float* W, void* w_bf16, int N, int in_dim, int F, int heads, int concat,
                             void* out, int out_bf16, cudaStream_t st);

// tensor-pipe variant (gat_tc.cu): bf16 node features, transform on tcgen05 (tf32), inference only
bool gat_tc_supported(int N, int in_dim, int F, int heads, int concat, int out_bf16);
// (runs its own score / edge-max pre-pass into s (N, 2*heads) and gmax (G, heads))
int gat_tc_launch(const void* x, const int32_t* rowptr, const int32_t* col, float* s, float* gmax, float* u, const float* W,
                  const float* a, int N, int64_t E, int in_dim, int F, int heads, int concat, float slope, int nodes_per_graph, void* out,
                  int out_bf16, cudaStream_t st);

// ------------------------------------------------------------------------------------------
// fused kernel
// ------------------------------------------------------------------------------------------
struct GatFusedArgs {
  GatAggArgs agg;
  const float* W;        // (heads, F, in)
  void* out;             // (N, concat ? heads*F : F)
  float* save_den;       // (N, heads) or null
  float* save_z;         // (N, heads, in) or null
  int F, concat, out_bf16;
  int tile_nodes;        // TN (multiple of 8)
  int rn;                // nodes per thread in the transform phase (1,2,4,8)
  int in_pad, f_pad;     // in / F rounded up to multiples of 4
};

template <int RN>
__device__ __forceinline__ void transform_tile(const GatFusedArgs& A, const float* __restrict__ Wt,
                                               const float* __restrict__ Zs, int tile_base, int zs_stride) {
  const int FG = A.f_pad >> 2;
  const int items = (A.tile_nodes / RN) * FG;
  const int heads = A.agg.heads;
  const int out_w = A.concat ? heads * A.F : A.F;
  const float inv_h = 1.f / (float)heads;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int ng = it / FG, fg = it - ng * FG;
    const int n0 = ng * RN;
    if (tile_base + n0 >= A.agg.N) continue;
    float oacc[RN][4];
#pragma unroll
    for (int r = 0; r < RN; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) oacc[r][c] = 0.f;
    for (int h = 0; h < heads; ++h) {
      // acc[r][c] += z[r][i] * W[i][c] on packed fp32 pairs (FFMA2: the same IEEE fmaf per element, half the instructions)
      unsigned long long acc2[RN][2];
#pragma unroll
      for (int r = 0; r < RN; ++r) acc2[r][0] = acc2[r][1] = 0ull;
      const float* wh = Wt + (size_t)h * A.in_pad * A.f_pad + fg * 4;
      const float* zh = Zs + (size_t)n0 * zs_stride + h * A.in_pad;
      for (int i = 0; i < A.in_pad; i += 4) {
        float4 zv[RN];
#pragma unroll
        for (int r = 0; r < RN; ++r) zv[r] = *reinterpret_cast<const float4*>(zh + (size_t)r * zs_stride + i);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const float4 wv = *reinterpret_cast<const float4*>(wh + (size_t)(i + ii) * A.f_pad);
          const unsigned long long w01 = f32x2_pack(wv.x, wv.y), w23 = f32x2_pack(wv.z, wv.w);
#pragma unroll
          for (int r = 0; r < RN; ++r) {
            const float zz = ii == 0 ? zv[r].x : (ii == 1 ? zv[r].y : (ii == 2 ? zv[r].z : zv[r].w));
            const unsigned long long z2 = f32x2_pack(zz, zz);
            f32x2_fma(acc2[r][0], z2, w01);
            f32x2_fma(acc2[r][1], z2, w23);
          }
        }
      }
      float acc[RN][4];
#pragma unroll
      for (int r = 0; r < RN; ++r) {
        f32x2_unpack(acc2[r][0], acc[r][0], acc[r][1]);
        f32x2_unpack(acc2[r][1], acc[r][2], acc[r][3]);
      }
      // per-head epilogue: ELU, then concat store or running head mean
#pragma unroll
      for (int r = 0; r < RN; ++r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float v = elu1(acc[r][c]);
          if (A.concat) {
            const int n = tile_base + n0 + r, f = fg * 4 + c;
            if (n < A.agg.N && f < A.F) {
              const size_t o = (size_t)n * out_w + (size_t)h * A.F + f;
              if (A.out_bf16) reinterpret_cast<__nv_bfloat16*>(A.out)[o] = __float2bfloat16_rn(v);
              else reinterpret_cast<float*>(A.out)[o] = v;
            }
          } else {
            oacc[r][c] += v;
          }
        }
      }
    }
    if (!A.concat) {
#pragma unroll
      for (int r = 0; r < RN; ++r) {
        const int n = tile_base + n0 + r;
        if (n >= A.agg.N) continue;
        const int f0 = fg * 4;
        const size_t o = (size_t)n * out_w + f0;
        if (f0 + 3 < A.F && (out_w & 3) == 0) {
          // torch.mean(stack) = sum / heads (graph_attention.py:158)
          const float v0 = oacc[r][0] * inv_h, v1 = oacc[r][1] * inv_h, v2 = oacc[r][2] * inv_h, v3 = oacc[r][3] * inv_h;
          if (A.out_bf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v0, v1), hi = __floats2bfloat162_rn(v2, v3);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned*>(&lo);
            pk.y = *reinterpret_cast<unsigned*>(&hi);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(A.out) + o) = pk;
          } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(A.out) + o) = make_float4(v0, v1, v2, v3);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (f0 + c < A.F) {
              const float v = oacc[r][c] * inv_h;
              if (A.out_bf16) reinterpret_cast<__nv_bfloat16*>(A.out)[o + c] = __float2bfloat16_rn(v);
              else reinterpret_cast<float*>(A.out)[o + c] = v;
            }
          }
        }
      }
    }
  }
}

template <typename TX, int NH, int V, int T>
__global__ void __launch_bounds__(256) gat_fused_kernel(const GatFusedArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int heads = A.agg.heads, in_dim = A.agg.in_dim;
  const int zs_stride = heads * A.in_pad;
  float* Wt = reinterpret_cast<float*>(smem_raw);                          // [heads][in_pad][f_pad]
  float* Zs = Wt + (size_t)heads * A.in_pad * A.f_pad;                     // [TN][heads*in_pad]
  WarpScratch* scratch = reinterpret_cast<WarpScratch*>(Zs + (size_t)A.tile_nodes * zs_stride);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

  // one-time: weights -> shared memory, transposed to [h][i][f] (f fastest), zero padded
  const int wt_elems = heads * A.in_pad * A.f_pad;
  for (int idx = threadIdx.x; idx < wt_elems; idx += blockDim.x) Wt[idx] = 0.f;
  for (int idx = threadIdx.x; idx < A.tile_nodes * zs_stride; idx += blockDim.x) Zs[idx] = 0.f;
  __syncthreads();
  for (int idx = threadIdx.x; idx < heads * A.F * in_dim; idx += blockDim.x) {
    const int h = idx / (A.F * in_dim);
    const int rem = idx - h * A.F * in_dim;
    const int f = rem / in_dim, i = rem - f * in_dim;
    Wt[((size_t)h * A.in_pad + i) * A.f_pad + f] = __ldg(A.W + idx);
  }
  __syncthreads();

  const int ntiles = ceil_div(A.agg.N, A.tile_nodes);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tile_base = tile * A.tile_nodes;
    // ---- phase 1: gather + attention-weighted aggregation, one warp per destination ----
    for (int q = warp; q < A.tile_nodes; q += nwarps) {
      const int j = tile_base + q;
      if (j >= A.agg.N) break;
      float z[NH][V * T], den[NH];
      gat_aggregate_node<TX, NH, V, T>(A.agg, j, lane, scratch + warp, z, den);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        if (h < heads) {
          const float dn = den[h] + 1e-10f;                              // graph_attention.py:96
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const int d = LaneDims<V, T>::dim(lane, t);
            if (d < in_dim) {
#pragma unroll
              for (int v = 0; v < V; ++v) {
                const float zn = z[h][t * V + v] / dn;
                Zs[(size_t)q * zs_stride + h * A.in_pad + d + v] = zn;
                if (A.save_z) A.save_z[((size_t)j * heads + h) * in_dim + d + v] = zn;
              }
            }
          }
          if (A.save_den && lane == 0) A.save_den[(size_t)j * heads + h] = den[h];
        }
      }
    }
    __syncthreads();
    // ---- phase 2: per-head transform from shared memory + ELU + head mean / concat ----
    switch (A.rn) {
      case 8: transform_tile<8>(A, Wt, Zs, tile_base, zs_stride); break;
      case 4: transform_tile<4>(A, Wt, Zs, tile_base, zs_stride); break;
      case 2: transform_tile<2>(A, Wt, Zs, tile_base, zs_stride); break;
      default: transform_tile<1>(A, Wt, Zs, tile_base, zs_stride); break;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// unfused fallback: aggregate to global z, then tiled FP32 GEMM with the same epilogue
// ------------------------------------------------------------------------------------------
template <typename TX, int NH, int V, int T>
__global__ void __launch_bounds__(256) gat_aggregate_kernel(const GatAggArgs a, float* __restrict__ zout,
                                                            float* __restrict__ save_den) {
  __shared__ WarpScratch scratch[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int parts = a.dim_parts > 0 ? a.dim_parts : 1;
  const int64_t ntasks = (int64_t)a.N * parts;
  for (int64_t task = wg; task < ntasks; task += nw) {
    const int j = (int)(task / parts), part = (int)(task - (int64_t)j * parts);
    const int dim0 = part * 32 * V * T;              // wide rows: several warps per destination, 32*V*T dims each
    float z[NH][V * T], den[NH];
    gat_aggregate_node<TX, NH, V, T>(a, j, lane, scratch + warp, z, den, dim0);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      if (h < a.heads) {
        const float dn = den[h] + 1e-10f;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int d = dim0 + LaneDims<V, T>::dim(lane, t);
          if (d < a.in_dim) {
            const size_t o = ((size_t)j * a.heads + h) * a.in_dim + d;
            float zn[V];
#pragma unroll
            for (int v = 0; v < V; ++v) zn[v] = z[h][t * V + v] / dn;
            if (a.z_bf16) {
              __nv_bfloat16* zb = reinterpret_cast<__nv_bfloat16*>(zout) + o;
              if (V == 4) {                         // one 8-byte store per lane: 256 contiguous bytes per warp
                __nv_bfloat162 lo = __floats2bfloat162_rn(zn[0], zn[1]), hi = __floats2bfloat162_rn(zn[2 % V], zn[3 % V]);
                uint2 pk;
                pk.x = *reinterpret_cast<unsigned*>(&lo); pk.y = *reinterpret_cast<unsigned*>(&hi);
                *reinterpret_cast<uint2*>(zb) = pk;
              } else {
#pragma unroll
                for (int v = 0; v < V; ++v) zb[v] = __float2bfloat16_rn(zn[v]);
              }
            } else if (V == 4) {
              *reinterpret_cast<float4*>(zout + o) = make_float4(zn[0], zn[1 % V], zn[2 % V], zn[3 % V]);
            } else if (V == 2) {
              *reinterpret_cast<float2*>(zout + o) = make_float2(zn[0], zn[1 % V]);
            } else {
              zout[o] = zn[0];
            }
          }
        }
        if (save_den && lane == 0 && part == 0) save_den[(size_t)j * a.heads + h] = den[h];
      }
    }
  }
}

template <typename TX, int NH, int V, int T>
static int launch_fused(const GatFusedArgs& A, size_t smem, int grid, cudaStream_t st) {
  auto k = gat_fused_kernel<TX, NH, V, T>;
  if (smem > 48 * 1024) {          // per launch: function attributes are per device
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget) != cudaSuccess) {
      set_error("gat_fused_kernel: cannot raise dynamic shared memory to %d", (int)kSmemBudget);
      return MG_ERR_CUDA;
    }
  }
  k<<<grid, 256, smem, st>>>(A);
  return check_launch("gat_fused_kernel");
}

template <typename TX, int NH>
static int dispatch_fused_vt(const GatFusedArgs& A, DimCfg d, size_t smem, int grid, cudaStream_t st) {
  if (d.V == 1 && d.T == 1) return launch_fused<TX, NH, 1, 1>(A, smem, grid, st);
  if (d.V == 2 && d.T == 1) return launch_fused<TX, NH, 2, 1>(A, smem, grid, st);
  if (d.V == 4 && d.T == 1) return launch_fused<TX, NH, 4, 1>(A, smem, grid, st);
  set_error("gat_fused: no variant for V=%d T=%d", d.V, d.T);
  return MG_ERR_UNSUPPORTED;
}

template <typename TX>
static int dispatch_fused(const GatFusedArgs& A, int NH, DimCfg d, size_t smem, int grid, cudaStream_t st) {
  switch (NH) {
    case 1: return dispatch_fused_vt<TX, 1>(A, d, smem, grid, st);
    case 2: return dispatch_fused_vt<TX, 2>(A, d, smem, grid, st);
    case 4: return dispatch_fused_vt<TX, 4>(A, d, smem, grid, st);
    default: return dispatch_fused_vt<TX, 8>(A, d, smem, grid, st);
  }
}

template <typename TX, int NH>
static int dispatch_agg_vt(const GatAggArgs& a, DimCfg d, float* z, float* den, int grid, cudaStream_t st) {
#define MG_AGG(VV, TT)                                                              \
  if (d.V == VV && d.T == TT) {                                                     \
    gat_aggregate_kernel<TX, NH, VV, TT><<<grid, 256, 0, st>>>(a, z, den);          \
    return check_launch("gat_aggregate_kernel");                                    \
  }
  MG_AGG(1, 1) MG_AGG(2, 1) MG_AGG(4, 1) MG_AGG(4, 2) MG_AGG(4, 4)
#undef MG_AGG
  set_error("gat_aggregate: no variant for V=%d T=%d", d.V, d.T);
  return MG_ERR_UNSUPPORTED;
}

template <typename TX>
static int dispatch_agg(const GatAggArgs& a, int NH, DimCfg d, float* z, float* den, int grid, cudaStream_t st) {
  switch (NH) {
    case 1: return dispatch_agg_vt<TX, 1>(a, d, z, den, grid, st);
    case 2: return dispatch_agg_vt<TX, 2>(a, d, z, den, grid, st);
    case 4: return dispatch_agg_vt<TX, 4>(a, d, z, den, grid, st);
    default: return dispatch_agg_vt<TX, 8>(a, d, z, den, grid, st);
  }
}

}  // namespace mg
