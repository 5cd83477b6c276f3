// K1 patch mean pooling, region mean pooling (a11) and K7 nearest un-pool (a13).
// All three are HBM-streaming kernels: K1 reads the feature map once with 16-byte loads,
// K7 writes the dense (B,D,H,W) map once with 16-byte streaming stores.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

#include "pool_tma_kernels.cuh"

namespace mg {

// channels per block = warps per block, one channel strip per warp.  Measured at cfg 2 (B200, two shard branches):
// 32 / 16 / 8 / 4 / 2 channels per block -> 169.2 / 168.5 / 171.2 / 164.4 / 163.9 us per step: many short blocks ramp the
// HBM pipeline faster than few long ones.
static int pool_cc() {
  static const int v = getenv("MG_POOL_CC") ? atoi(getenv("MG_POOL_CC")) : 4;
  return v < 1 ? 1 : (v > 32 ? 32 : v);
}
constexpr int kPoolRows = 16; // rows whose 16-byte loads are issued together

// fast path: pw % VEC == 0, Wf % VEC == 0, lpp = pw/VEC a power of two <= 32.
// Block = (x chunk of 32 vectors, channel group) x patch row x image; warp w takes channels w, w+NW, ...
// of the group and, per channel, issues all ph row loads of its 512-byte strip before summing them.
template <typename TX, typename TO>
__global__ void __launch_bounds__(256) pool_patches_vec_kernel(const TX* __restrict__ x, int C, int Hf, int Wf, int ph,
                                                               int pw, int Hp, int Wp, int cg, int xchunks,
                                                               TO* __restrict__ out) {
  constexpr int VEC = Vec16<TX>::N;
  extern __shared__ float tile[];   // [32 / lpp][cg + 1]
  const int b = blockIdx.z, py = blockIdx.y;
  const int xc = blockIdx.x % xchunks, c0 = (blockIdx.x / xchunks) * cg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int lpp = pw / VEC;                               // lanes per patch
  const int pp = 32 / lpp;                                // patches per chunk
  const int y0 = py * ph, rows = min(Hf, y0 + ph) - y0;
  const float inv = 1.f / (float)(ph * pw);
  const int nvec = Wf / VEC;
  const int xv = xc * 32 + lane;
  const bool in_x = xv < nvec;
  const int ncc = min(cg, C - c0);
  for (int cc = warp; cc < ncc; cc += nw) {
    const TX* p = x + (((size_t)b * C + c0 + cc) * Hf + y0) * Wf + (size_t)xv * VEC;
    float acc = 0.f;
    for (int r0 = 0; r0 < rows; r0 += kPoolRows) {
      float part[kPoolRows];
#pragma unroll
      for (int r = 0; r < kPoolRows; ++r)
        part[r] = (in_x && r0 + r < rows) ? Vec16<TX>::sum(p + (size_t)(r0 + r) * Wf) : 0.f;
#pragma unroll
      for (int st = kPoolRows / 2; st > 0; st >>= 1)
#pragma unroll
        for (int r = 0; r < st; ++r) part[r] += part[r + st];
      acc += part[0];
    }
    for (int o = 1; o < lpp; o <<= 1) acc += __shfl_xor_sync(kFull, acc, o);
    if ((lane & (lpp - 1)) == 0) tile[(lane / lpp) * (cg + 1) + cc] = acc * inv;
  }
  __syncthreads();
  const int px0 = xc * pp;
  for (int idx = threadIdx.x; idx < pp * ncc; idx += blockDim.x) {
    const int pl = idx / ncc, cc = idx - pl * ncc;
    if (px0 + pl < Wp)
      out[((size_t)b * Hp * Wp + (size_t)py * Wp + px0 + pl) * C + c0 + cc] = from_f32<TO>(tile[pl * (cg + 1) + cc]);
  }
}


// generic path: any window; one thread per output element (px fastest so window reads share lines)
template <typename TX, typename TO>
__global__ void pool_patches_generic_kernel(const TX* __restrict__ x, int B, int C, int Hf, int Wf, int ph, int pw,
                                            int Hp, int Wp, TO* __restrict__ out) {
  const int64_t total = (int64_t)B * C * Hp * Wp;
  const float inv = 1.f / (float)(ph * pw);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int px = (int)(t % Wp);
    const int py = (int)((t / Wp) % Hp);
    const int c = (int)((t / ((int64_t)Wp * Hp)) % C);
    const int b = (int)(t / ((int64_t)Wp * Hp * C));
    const TX* plane = x + ((size_t)b * C + c) * Hf * Wf;
    const int y1 = min(Hf, (py + 1) * ph), x1 = min(Wf, (px + 1) * pw);
    float acc = 0.f;
    for (int y = py * ph; y < y1; ++y)
      for (int xx = px * pw; xx < x1; ++xx) acc += to_f32<TX>(plane[(size_t)y * Wf + xx]);
    out[((size_t)b * Hp * Wp + (size_t)py * Wp + px) * C + c] = from_f32<TO>(acc * inv);
  }
}

// ------------------------------------------------------------------------------------------
// a11: region mean pool.  Two deterministic stages (no atomics, fixed reduction order):
//   partial: block (chunk, image) reduces its chunk of nodes with warp-private shared accumulators;
//   finalize: thread per (image, k, d) sums the chunk partials in chunk order and divides by the count.
// ------------------------------------------------------------------------------------------
constexpr int kSegChunkNodes = 64;

__global__ void __launch_bounds__(256) segment_partial_kernel(const float* __restrict__ h, const int32_t* __restrict__ labels,
                                                              int N, int D, int K, float* __restrict__ part,
                                                              int32_t* __restrict__ part_cnt) {
  extern __shared__ float acc[];                       // [8][K][D] then int cnt[8][K]
  int* cnt = reinterpret_cast<int*>(acc + (size_t)8 * K * D);
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * K * D; i += blockDim.x) acc[i] = 0.f;
  for (int i = threadIdx.x; i < 8 * K; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  float* mine = acc + (size_t)warp * K * D;
  const int n0 = chunk * kSegChunkNodes, n1 = min(N, n0 + kSegChunkNodes);
  for (int n = n0 + warp; n < n1; n += 8) {
    const int k = labels ? __ldg(labels + (size_t)b * N + n) : n;
    if (k < 0 || k >= K) continue;
    const float* row = h + ((size_t)b * N + n) * D;
    for (int d = lane; d < D; d += 32) mine[k * D + d] += __ldg(row + d);
    if (lane == 0) cnt[warp * K + k] += 1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += acc[(size_t)w * K * D + i];
    part[((size_t)b * nchunks + chunk) * K * D + i] = s;
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    int c = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) c += cnt[w * K + k];
    part_cnt[((size_t)b * nchunks + chunk) * K + k] = c;
  }
}

// mean != 0: out = sum / count (0 for empty regions, train_end_to_end.py:372-373); else out = scale * sum
__global__ void segment_finalize_kernel(const float* __restrict__ part, const int32_t* __restrict__ part_cnt, int B, int nchunks,
                                        int K, int D, int mean, float scale, float* __restrict__ out,
                                        int32_t* __restrict__ counts) {
  const int total = B * K * D;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int d = t % D, k = (t / D) % K, b = t / (D * K);
    float s = 0.f;
    int c = 0;
    for (int ch = 0; ch < nchunks; ++ch) {
      s += part[((size_t)b * nchunks + ch) * K * D + k * D + d];
      c += part_cnt[((size_t)b * nchunks + ch) * K + k];
    }
    out[t] = mean ? (c > 0 ? s / (float)c : 0.f) : s * scale;
    if (counts && d == 0) counts[b * K + k] = c;
  }
}

// shared by mg_segment_mean and the un-pool backward (block_backward.cu).  work: segment_work_bytes().
int64_t segment_work_bytes(int B, int N, int D, int K) {
  const int nchunks = ceil_div(N, kSegChunkNodes);
  return (int64_t)B * nchunks * K * (D + 1) * 4 + 256;
}
int segment_reduce_launch(const float* h, const int32_t* labels, int B, int N, int D, int K, int mean, float scale, float* out,
                          int32_t* counts, void* work, cudaStream_t st) {
  const size_t smem = (size_t)8 * K * D * 4 + (size_t)8 * K * 4;
  MG_REQUIRE(smem <= 200 * 1024, MG_ERR_UNSUPPORTED, "segment reduce: K*D=%d too large for shared accumulators", K * D);
  const int nchunks = ceil_div(N, kSegChunkNodes);
  float* part = reinterpret_cast<float*>(work);
  int32_t* part_cnt = reinterpret_cast<int32_t*>(part + (size_t)B * nchunks * K * D);
  if (smem > 48 * 1024) cudaFuncSetAttribute(segment_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  dim3 grid(nchunks, B);
  MG_REQUIRE(B <= 65535, MG_ERR_INVALID, "segment reduce: batch too large");
  segment_partial_kernel<<<grid, 256, smem, st>>>(h, labels, N, D, K, part, part_cnt);
  int rc = check_launch("segment_partial_kernel");
  if (rc) return rc;
  segment_finalize_kernel<<<std::min(ceil_div(B * K * D, 256), num_sms() * 4), 256, 0, st>>>(part, part_cnt, B, nchunks, K, D, mean,
                                                                                       scale, out, counts);
  return check_launch("segment_finalize_kernel");
}

// ------------------------------------------------------------------------------------------
// K7: nearest un-pool.  out[b,d,y,x] = table[b, label[b, iy(y)*Wp + ix(x)], d]
// index rule = torch CPU upsample_nearest ('nearest'): identity, >>1, else
// min(floor(dst * float(in/out)), in-1) in fp32.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size, float scale) {
  if (out_size == in_size) return dst;
  if (out_size == 2 * in_size) return dst >> 1;
  const int s = (int)floorf(__fmul_rn((float)dst, scale));
  return min(s, in_size - 1);
}

template <typename TO>
struct Pack;
template <>
struct Pack<float> {
  static constexpr int VEC = 4;
  static __device__ __forceinline__ uint4 make(const float* v) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
};
template <>
struct Pack<__nv_bfloat16> {
  static constexpr int VEC = 8;
  static __device__ __forceinline__ uint4 make(const float* v) {
    uint4 r;
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]),
                   c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    r.x = *reinterpret_cast<unsigned*>(&a); r.y = *reinterpret_cast<unsigned*>(&b);
    r.z = *reinterpret_cast<unsigned*>(&c); r.w = *reinterpret_cast<unsigned*>(&d);
    return r;
  }
};

constexpr int kUnTX = 64, kUnTY = 4, kUnRY = 16, kUnDC = 8;

// vector path (W % VEC == 0, 16-byte aligned rows): grid (xblocks, yblocks, B * dchunks)
template <typename TO>
__global__ void __launch_bounds__(kUnTX* kUnTY) unpool_vec_kernel(const float* __restrict__ table,
                                                                 const int32_t* __restrict__ labels, int K, int D,
                                                                 int Hp, int Wp, int H, int W, TO* __restrict__ out,
                                                                 int64_t out_batch_stride, float sy, float sx) {
  constexpr int VEC = Pack<TO>::VEC;
  const int dchunks = ceil_div(D, kUnDC);
  const int b = blockIdx.z / dchunks, d0 = (blockIdx.z - b * dchunks) * kUnDC;
  const int nd = min(kUnDC, D - d0);
  const int xv = blockIdx.x * kUnTX + threadIdx.x;
  if (xv * VEC >= W) return;
  int px[VEC];
  bool same = true;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    px[v] = nearest_src(xv * VEC + v, Wp, W, sx);
    same = same && (px[v] == px[0]);
  }
  const int yend = min(H, (int)(blockIdx.y + 1) * kUnRY);
  const float* tb = table + (size_t)b * K * D;
  const int32_t* lb = labels ? labels + (size_t)b * Hp * Wp : nullptr;
  TO* ob = out + (size_t)b * out_batch_stride;
  int prev_py = -1;
  int lab[VEC];
  for (int y = blockIdx.y * kUnRY + threadIdx.y; y < yend; y += kUnTY) {
    const int py = nearest_src(y, Hp, H, sy);
    if (py != prev_py) {
      prev_py = py;
      if (same) {
        const int n = py * Wp + px[0];
        const int l = lb ? __ldg(lb + n) : n;
#pragma unroll
        for (int v = 0; v < VEC; ++v) lab[v] = l;
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const int n = py * Wp + px[v];
          lab[v] = lb ? __ldg(lb + n) : n;
        }
      }
    }
    TO* orow = ob + ((size_t)d0 * H + y) * W + (size_t)xv * VEC;
    if (same) {
      const float* tp = tb + (size_t)lab[0] * D + d0;
#pragma unroll
      for (int d = 0; d < kUnDC; ++d) {
        if (d < nd) {
          const float t = __ldg(tp + d);
          float vals[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) vals[v] = t;
          st_cs_v4(orow + (size_t)d * H * W, Pack<TO>::make(vals));
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < kUnDC; ++d) {
        if (d < nd) {
          float vals[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) vals[v] = __ldg(tb + (size_t)lab[v] * D + d0 + d);
          st_cs_v4(orow + (size_t)d * H * W, Pack<TO>::make(vals));
        }
      }
    }
  }
}

// scalar path: any W / alignment
template <typename TO>
__global__ void unpool_scalar_kernel(const float* __restrict__ table, const int32_t* __restrict__ labels, int B, int K,
                                     int D, int Hp, int Wp, int H, int W, TO* __restrict__ out,
                                     int64_t out_batch_stride, float sy, float sx) {
  const int64_t total = (int64_t)B * D * H * W;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(t % W);
    const int y = (int)((t / W) % H);
    const int d = (int)((t / ((int64_t)W * H)) % D);
    const int b = (int)(t / ((int64_t)W * H * D));
    const int n = nearest_src(y, Hp, H, sy) * Wp + nearest_src(xx, Wp, W, sx);
    const int l = labels ? __ldg(labels + (size_t)b * Hp * Wp + n) : n;
    out[(size_t)b * out_batch_stride + ((size_t)d * H + y) * W + xx] = from_f32<TO>(__ldg(table + ((size_t)b * K + l) * D + d));
  }
}

template <typename TX, typename TO>
static int launch_pool(const void* x, int B, int C, int Hf, int Wf, int ph, int pw, void* out, cudaStream_t st) {
  constexpr int VEC = Vec16<TX>::N;
  const int Hp = ceil_div(Hf, ph), Wp = ceil_div(Wf, pw);
  const int lpp = pw / VEC;
  const bool fast = (pw % VEC == 0) && (Wf % VEC == 0) && lpp >= 1 && lpp <= 32 && (lpp & (lpp - 1)) == 0 &&
                    ((uintptr_t)x % 16 == 0);
  // MG_POOL_VARIANT=-1 forces the LDG kernel below (A/B runs); default 0 = bulk-copy staged kernel where it applies
  static const int variant = getenv("MG_POOL_VARIANT") ? atoi(getenv("MG_POOL_VARIANT")) : 0;
  // default: TMA-staged persistent kernel (at least one whole row per ring buffer, <= 8 column passes per warp)
  const int row_bytes = Wf * (int)sizeof(TX);
  static const int stages_env = getenv("MG_POOL_STAGES") ? atoi(getenv("MG_POOL_STAGES")) : 3;
  static const int chunk_env = getenv("MG_POOL_CHUNK") ? atoi(getenv("MG_POOL_CHUNK")) : 8192;
  {
    int stages = std::max(2, std::min(8, stages_env));
    int stage_bytes = std::max(row_bytes, std::max(1024, chunk_env)) / 128 * 128;
    stage_bytes = std::max(stage_bytes, (row_bytes + 127) / 128 * 128);
    while (stages > 2 && (size_t)kPtWarps * stages * (stage_bytes + 8) > (size_t)kPtSmemBytes) --stages;
    const size_t smem = (size_t)kPtWarps * stages * (stage_bytes + 8);
    if (fast && variant == 0 && smem <= (size_t)kPtSmemBytes && ceil_div(Wf / VEC, 32) <= kPtMaxPasses && lpp <= 32) {
      PoolTmaArgs A;
      A.x = x; A.out = out; A.C = C; A.Hf = Hf; A.Wf = Wf; A.ph = ph; A.pw = pw; A.Hp = Hp; A.Wp = Wp;
      A.rpc = std::max(1, std::min(ph, stage_bytes / row_bytes));
      A.nstrips = B * Hp * C;
      A.stages = stages;
      A.stage_bytes = stage_bytes;
      auto kern = pool_patches_tma_kernel<TX, TO>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPtSmemBytes) != cudaSuccess) {
        set_error("mg_pool_patches: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
        return MG_ERR_CUDA;
      }
      const int grid = std::min(num_sms(), ceil_div(A.nstrips, kPtWarps));
      kern<<<grid, kPtWarps * 32, smem, st>>>(A);
      return check_launch("pool_patches_tma_kernel");
    }
  }
  if (fast) {
    const int cg = std::min(C, pool_cc());
    // warps per block: the count in 4..8 that wastes the fewest channel slots (ties -> more warps)
    int nw = std::min(cg, 8), best_waste = 1 << 30;
    for (int w = std::min(cg, 8); w >= std::min(cg, 4); --w) {
      const int waste = ceil_div(cg, w) * w - cg;
      if (waste < best_waste) { best_waste = waste; nw = w; }
    }
    const int xchunks = ceil_div(Wf / VEC, 32);
    const size_t smem = (size_t)(32 / lpp) * (cg + 1) * 4;
    dim3 grid(xchunks * ceil_div(C, cg), Hp, B);
    pool_patches_vec_kernel<TX, TO><<<grid, nw * 32, smem, st>>>(reinterpret_cast<const TX*>(x), C, Hf, Wf, ph, pw, Hp, Wp,
                                                               cg, xchunks, reinterpret_cast<TO*>(out));
    return check_launch("pool_patches_vec_kernel");
  }
  const int64_t total = (int64_t)B * C * Hp * Wp;
  const int grid = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 16);
  pool_patches_generic_kernel<TX, TO><<<grid, 256, 0, st>>>(reinterpret_cast<const TX*>(x), B, C, Hf, Wf, ph, pw, Hp, Wp,
                                                           reinterpret_cast<TO*>(out));
  return check_launch("pool_patches_generic_kernel");
}

template <typename TO>
static int launch_unpool(const float* table, const int32_t* labels, int B, int K, int D, int Hp, int Wp, int H, int W,
                         void* out, int64_t stride, cudaStream_t st) {
  constexpr int VEC = Pack<TO>::VEC;
  // torch: scale = float(in) / float(out)
  const float sy = (float)Hp / (float)H, sx = (float)Wp / (float)W;
  const bool vec_ok = (W % VEC == 0) && ((uintptr_t)out % 16 == 0) && ((stride * (int64_t)sizeof(TO)) % 16 == 0);
  if (vec_ok) {
    dim3 block(kUnTX, kUnTY);
    dim3 grid(ceil_div(W / VEC, kUnTX), ceil_div(H, kUnRY), B * ceil_div(D, kUnDC));
    MG_REQUIRE(grid.z <= 65535 && grid.y <= 65535, MG_ERR_INVALID, "mg_unpool_nearest: grid too large");
    unpool_vec_kernel<TO><<<grid, block, 0, st>>>(table, labels, K, D, Hp, Wp, H, W, reinterpret_cast<TO*>(out), stride, sy, sx);
    return check_launch("unpool_vec_kernel");
  }
  const int64_t total = (int64_t)B * D * H * W;
  const int grid = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 32);
  unpool_scalar_kernel<TO><<<grid, 256, 0, st>>>(table, labels, B, K, D, Hp, Wp, H, W, reinterpret_cast<TO*>(out), stride, sy, sx);
  return check_launch("unpool_scalar_kernel");
}

}  // namespace mg

using namespace mg;

extern "C" {

int mg_pool_patches(const void* x, int x_dtype, int B, int C, int Hf, int Wf, int ph, int pw, void* out, int out_dtype,
                    mg_stream_t stream) {
  MG_REQUIRE(x && out && B > 0 && C > 0 && Hf > 0 && Wf > 0 && ph > 0 && pw > 0, MG_ERR_INVALID,
             "mg_pool_patches: bad arguments");
  MG_REQUIRE(B <= 65535 && ceil_div(Hf, ph) <= 65535, MG_ERR_INVALID, "mg_pool_patches: grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == MG_F32 && out_dtype == MG_F32) return launch_pool<float, float>(x, B, C, Hf, Wf, ph, pw, out, st);
  if (x_dtype == MG_BF16 && out_dtype == MG_BF16)
    return launch_pool<__nv_bfloat16, __nv_bfloat16>(x, B, C, Hf, Wf, ph, pw, out, st);
  if (x_dtype == MG_BF16 && out_dtype == MG_F32) return launch_pool<__nv_bfloat16, float>(x, B, C, Hf, Wf, ph, pw, out, st);
  if (x_dtype == MG_F32 && out_dtype == MG_BF16) return launch_pool<float, __nv_bfloat16>(x, B, C, Hf, Wf, ph, pw, out, st);
  set_error("mg_pool_patches: unsupported dtypes %d -> %d", x_dtype, out_dtype);
  return MG_ERR_INVALID;
}

int64_t mg_segment_work_bytes(int B, int N, int D, int K) {
  if (B <= 0 || N <= 0 || D <= 0 || K <= 0) return 0;
  return segment_work_bytes(B, N, D, K);
}

int mg_segment_mean(const float* h, const int32_t* labels, int B, int N, int D, int K, float* out, int32_t* counts, void* work,
                    mg_stream_t stream) {
  MG_REQUIRE(h && labels && out && work && B > 0 && N > 0 && D > 0 && K > 0, MG_ERR_INVALID, "mg_segment_mean: bad arguments");
  return segment_reduce_launch(h, labels, B, N, D, K, 1, 1.f, out, counts, work, (cudaStream_t)stream);
}

int mg_unpool_nearest(const float* table, const int32_t* labels, int B, int K, int D, int Hp, int Wp, int H, int W,
                      void* out, int out_dtype, int64_t out_batch_stride, mg_stream_t stream) {
  MG_REQUIRE(table && out && B > 0 && K > 0 && D > 0 && Hp > 0 && Wp > 0 && H > 0 && W > 0, MG_ERR_INVALID,
             "mg_unpool_nearest: bad arguments");
  MG_REQUIRE(labels || K == Hp * Wp, MG_ERR_INVALID, "mg_unpool_nearest: labels==NULL requires K == Hp*Wp");
  MG_REQUIRE(out_batch_stride >= (int64_t)D * H * W, MG_ERR_INVALID, "mg_unpool_nearest: batch stride smaller than D*H*W");
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == MG_F32) return launch_unpool<float>(table, labels, B, K, D, Hp, Wp, H, W, out, out_batch_stride, st);
  if (out_dtype == MG_BF16)
    return launch_unpool<__nv_bfloat16>(table, labels, B, K, D, Hp, Wp, H, W, out, out_batch_stride, st);
  set_error("mg_unpool_nearest: unsupported out dtype %d", out_dtype);
  return MG_ERR_INVALID;
}

}  // extern "C"
