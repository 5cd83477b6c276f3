// Row f4 of the scope table: the two small losses that sit on tensors the graph block already holds.
//   FeatureConsistencyLoss  — model/unet/feature_loss.py:103-123 (per-patch contrastive distance, summed over patches,
//                             mean over the batch)
//   TVLoss                  — scripts/train_end_to_end.py:73-89 (squared forward differences along H and W)
// Both are HBM-streaming reads with deterministic two-stage reductions (per-warp partials in a caller-provided work
// buffer, fixed-order final sum in double) so a result is bitwise repeatable run to run; backward kernels are
// elementwise.  No atomics.
#include <algorithm>

#include "common.cuh"

namespace mg {

// ------------------------------------------------------------------------------------------
// label load: correspondence_map_y arrives as whatever the caller has (the reference calls
// .float() on it, feature_loss.py:106; train_end_to_end.py:342 passes torch.randint -> int64)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float load_y(const void* y, int y_dtype, size_t i) {
  switch (y_dtype) {
    case MG_F32: return __ldg(reinterpret_cast<const float*>(y) + i);
    case MG_BF16: return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(y) + i));
    case MG_I32: return (float)__ldg(reinterpret_cast<const int32_t*>(y) + i);
    default: return (float)__ldg(reinterpret_cast<const long long*>(y) + i);
  }
}

constexpr int kFlWarps = 8;          // patches per block pass (one warp per patch row)

// dist_sq of one patch row, all lanes return the full sum.  V elements per lane per step.
template <typename TU, typename TG, int V>
__device__ __forceinline__ float row_dist_sq(const TU* __restrict__ u, const TG* __restrict__ g, int D, int lane) {
  float acc = 0.f;
  for (int d = lane * V; d < D; d += 32 * V) {
    float a[V], b[V];
    VecLoad<TU, V>::ld(u + d, a);
    VecLoad<TG, V>::ld(g + d, b);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float df = a[i] - b[i];
      acc += df * df;
    }
  }
  return warp_sum(acc);
}

// grid = (chunks, B); block = kFlWarps warps; warp w of chunk c takes patches c*ppc + w, + kFlWarps, ...
// partial[b * chunks + c] = sum over the chunk's patches (fixed order: warp-serial, then warps 0..7)
template <typename TU, typename TG, int V>
__global__ void __launch_bounds__(kFlWarps * 32)
    feature_loss_partial_kernel(const TU* __restrict__ fu, const TG* __restrict__ fg, const void* __restrict__ y, int y_dtype,
                                int N, int D, float margin, int ppc, float* __restrict__ partial) {
  __shared__ float wsum[kFlWarps];
  const int b = blockIdx.y, c = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_end = min(N, (c + 1) * ppc);
  float acc = 0.f;
  for (int n = c * ppc + warp; n < n_end; n += kFlWarps) {
    const size_t row = (size_t)b * N + n;
    const float dsq = row_dist_sq<TU, TG, V>(fu + row * D, fg + row * D, D, lane);
    const float yv = load_y(y, y_dtype, row);
    const float dist = sqrtf(dsq + 1e-8f);                    // feature_loss.py:113
    const float hinge = fmaxf(margin - dist, 0.f);            // :115
    acc += yv * dsq + (1.f - yv) * (hinge * hinge);           // :107,116,118
  }
  if (lane == 0) wsum[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kFlWarps; ++w) s += wsum[w];
    partial[(size_t)b * gridDim.x + c] = s;
  }
}

// one block: per-image sums (fixed order) -> per_image (nullable) and their mean -> loss
__global__ void feature_loss_final_kernel(const float* __restrict__ partial, int B, int chunks, float* __restrict__ per_image,
                                          float* __restrict__ loss) {
  __shared__ double sh[256];
  double tot = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += (double)partial[(size_t)b * chunks + c];
    if (per_image) per_image[b] = (float)s;
    tot += (double)(float)s;                                  // torch: sum(dim=1) in fp32, then .mean()
  }
  sh[threadIdx.x] = tot;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)(sh[0] / (double)B);
}

// backward: gfu = c * (fu - fg), gfg = -gfu with c = (g/B) * (2y - (1-y) * 2*hinge/dist)
template <typename TU, typename TG, int V>
__global__ void __launch_bounds__(256)
    feature_loss_backward_kernel(const TU* __restrict__ fu, const TG* __restrict__ fg, const void* __restrict__ y, int y_dtype,
                                 int64_t rows, int D, float margin, float inv_b, const float* __restrict__ grad_loss,
                                 float* __restrict__ gfu, float* __restrict__ gfg) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TU* u = fu + row * D;
  const TG* g = fg + row * D;
  const float dsq = row_dist_sq<TU, TG, V>(u, g, D, lane);
  const float yv = load_y(y, y_dtype, row);
  const float dist = sqrtf(dsq + 1e-8f);
  const float hinge = fmaxf(margin - dist, 0.f);
  const float coef = __ldg(grad_loss) * inv_b * (2.f * yv - (1.f - yv) * 2.f * hinge / dist);
  for (int d = lane * V; d < D; d += 32 * V) {
    float a[V], b[V];
    VecLoad<TU, V>::ld(u + d, a);
    VecLoad<TG, V>::ld(g + d, b);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float v = coef * (a[i] - b[i]);
      if (gfu) gfu[row * D + d + i] = v;
      if (gfg) gfg[row * D + d + i] = -v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// TV loss.  Each warp owns 32*V consecutive columns of a strip of kTvRows rows of one plane and
// walks down it, so every element is loaded once (plus one halo row per strip); the right
// neighbour of a lane's last element comes from lane+1 by shuffle (lane 31: one scalar load).
// ------------------------------------------------------------------------------------------
constexpr int kTvRows = 16;
constexpr int kTvWarps = 4;

template <typename T, int V>
__device__ __forceinline__ void tv_load(const T* p, float* o);
template <>
__device__ __forceinline__ void tv_load<float, 4>(const float* p, float* o) { VecLoad<float, 4>::ld(p, o); }
template <>
__device__ __forceinline__ void tv_load<float, 1>(const float* p, float* o) { o[0] = __ldg(p); }
template <>
__device__ __forceinline__ void tv_load<__nv_bfloat16, 1>(const __nv_bfloat16* p, float* o) { o[0] = __bfloat162float(__ldg(p)); }
template <>
__device__ __forceinline__ void tv_load<__nv_bfloat16, 8>(const __nv_bfloat16* p, float* o) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __uint_as_float(w[i] << 16);
    o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// grid = (xchunks * strips, planes); block = kTvWarps warps, warp w -> x chunk (blockIdx.x % xcb) * kTvWarps + w
template <typename T, int V>
__global__ void __launch_bounds__(kTvWarps * 32)
    tv_partial_kernel(const T* __restrict__ x, int H, int W, int xcb, float2* __restrict__ partial) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int strip = blockIdx.x / xcb, xc = (blockIdx.x % xcb) * kTvWarps + warp;
  const T* plane = x + (size_t)blockIdx.y * H * W;
  const int x0 = (xc * 32 + lane) * V;
  const int y0 = strip * kTvRows, y1 = min(H, y0 + kTvRows);
  const bool in = x0 < W;                                       // W % V == 0 on the vector path
  float hs = 0.f, ws = 0.f;
  float cur[V], nxt[V];
#pragma unroll
  for (int i = 0; i < V; ++i) cur[i] = 0.f;
  if (in) tv_load<T, V>(plane + (size_t)y0 * W + x0, cur);
  for (int yy = y0; yy < y1; ++yy) {
    const bool has_down = yy + 1 < H;
    if (in && has_down) tv_load<T, V>(plane + (size_t)(yy + 1) * W + x0, nxt);
    // right neighbour of cur[V-1]
    float right = __shfl_down_sync(kFull, cur[0], 1);
    const bool has_right = in && (x0 + V < W);
    if (lane == 31 && has_right) {
      float t[1];
      tv_load<T, 1>(plane + (size_t)yy * W + x0 + V, t);
      right = t[0];
    }
    if (in) {
#pragma unroll
      for (int i = 0; i + 1 < V; ++i) {
        const float d = cur[i + 1] - cur[i];
        ws += d * d;
      }
      if (has_right) {
        const float d = right - cur[V - 1];
        ws += d * d;
      }
      if (has_down) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float d = nxt[i] - cur[i];
          hs += d * d;
          cur[i] = nxt[i];
        }
      }
    }
  }
  hs = warp_sum(hs);
  ws = warp_sum(ws);
  if (lane == 0)
    partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kTvWarps + warp] = make_float2(hs, ws);
}

// weight * (h_tv / count_h + w_tv / count_w) / batch in fp32, exactly the reference's expression (:89);
// a zero count divides 0 by 0 -> NaN, as torch does.
__global__ void tv_final_kernel(const float2* __restrict__ partial, int64_t n, float count_h, float count_w, float weight,
                                float batch, float* __restrict__ out) {
  __shared__ double sh[256], sw[256];
  double h = 0.0, w = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float2 p = partial[i];
    h += (double)p.x;
    w += (double)p.y;
  }
  sh[threadIdx.x] = h;
  sw[threadIdx.x] = w;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[threadIdx.x] += sh[threadIdx.x + o];
      sw[threadIdx.x] += sw[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float h_tv = (float)sh[0], w_tv = (float)sw[0];
    out[0] = weight * (h_tv / count_h + w_tv / count_w) / batch;
    out[1] = h_tv;
    out[2] = w_tv;
  }
}

// grad_x = g * weight / batch * ( 2/count_h * (2x - up - down restricted to existing neighbours) + same along W )
template <typename T>
__global__ void tv_backward_kernel(const T* __restrict__ x, int64_t total, int H, int W, float sh, float sw,
                                   const float* __restrict__ grad_loss, float* __restrict__ gx) {
  const float g = __ldg(grad_loss);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(t % W);
    const int yy = (int)((t / W) % H);
    const float c = to_f32<T>(x[t]);
    float gh = 0.f, gw = 0.f;
    if (yy > 0) gh += c - to_f32<T>(x[t - W]);
    if (yy + 1 < H) gh -= to_f32<T>(x[t + W]) - c;
    if (xx > 0) gw += c - to_f32<T>(x[t - 1]);
    if (xx + 1 < W) gw -= to_f32<T>(x[t + 1]) - c;
    gx[t] = g * (sh * gh + sw * gw);
  }
}

static int fl_chunks(int N) { return ceil_div(N, 8 * kFlWarps) < 64 ? ceil_div(N, 8 * kFlWarps) : 64; }

template <typename TU, typename TG>
static int launch_feature_loss(const void* fu, const void* fg, const void* y, int y_dtype, int B, int N, int D, float margin,
                               void* work, float* per_image, float* loss, cudaStream_t st) {
  const int chunks = fl_chunks(N);
  const int ppc = ceil_div(N, chunks);
  float* partial = reinterpret_cast<float*>(work);
  dim3 grid(chunks, B);
  const bool v4 = (D % 4 == 0) && ((uintptr_t)fu % 16 == 0) && ((uintptr_t)fg % 16 == 0);
  if (v4)
    feature_loss_partial_kernel<TU, TG, 4><<<grid, kFlWarps * 32, 0, st>>>(
        reinterpret_cast<const TU*>(fu), reinterpret_cast<const TG*>(fg), y, y_dtype, N, D, margin, ppc, partial);
  else
    feature_loss_partial_kernel<TU, TG, 1><<<grid, kFlWarps * 32, 0, st>>>(
        reinterpret_cast<const TU*>(fu), reinterpret_cast<const TG*>(fg), y, y_dtype, N, D, margin, ppc, partial);
  int rc = check_launch("feature_loss_partial_kernel");
  if (rc) return rc;
  feature_loss_final_kernel<<<1, 256, 0, st>>>(partial, B, chunks, per_image, loss);
  return check_launch("feature_loss_final_kernel");
}

template <typename TU, typename TG>
static int launch_feature_loss_bwd(const void* fu, const void* fg, const void* y, int y_dtype, int B, int N, int D, float margin,
                                   const float* grad_loss, float* gfu, float* gfg, cudaStream_t st) {
  const int64_t rows = (int64_t)B * N;
  const int grid = (int)ceil_div64(rows, 8);
  const bool v4 = (D % 4 == 0) && ((uintptr_t)fu % 16 == 0) && ((uintptr_t)fg % 16 == 0);
  if (v4)
    feature_loss_backward_kernel<TU, TG, 4><<<grid, 256, 0, st>>>(reinterpret_cast<const TU*>(fu),
                                                                 reinterpret_cast<const TG*>(fg), y, y_dtype, rows, D, margin,
                                                                 1.f / (float)B, grad_loss, gfu, gfg);
  else
    feature_loss_backward_kernel<TU, TG, 1><<<grid, 256, 0, st>>>(reinterpret_cast<const TU*>(fu),
                                                                 reinterpret_cast<const TG*>(fg), y, y_dtype, rows, D, margin,
                                                                 1.f / (float)B, grad_loss, gfu, gfg);
  return check_launch("feature_loss_backward_kernel");
}

struct TvShape {
  int V, xchunks, xcb, strips;
  int64_t partials;
};
static TvShape tv_shape(int dtype, const void* x, int planes, int H, int W) {
  TvShape s;
  const int vmax = dtype == MG_BF16 ? 8 : 4;
  const int esz = dtype == MG_BF16 ? 2 : 4;
  s.V = (W % vmax == 0 && ((uintptr_t)x % 16 == 0) && (((int64_t)W * esz) % 16 == 0)) ? vmax : 1;
  s.xchunks = ceil_div(W, 32 * s.V);
  s.xcb = ceil_div(s.xchunks, kTvWarps);
  s.strips = ceil_div(H, kTvRows);
  s.partials = (int64_t)planes * s.xcb * s.strips * kTvWarps;
  return s;
}

}  // namespace mg

using namespace mg;

extern "C" {

int64_t mg_feature_loss_work_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return (int64_t)B * fl_chunks(N) * (int64_t)sizeof(float);
}

int mg_feature_consistency_loss(const void* f_unet, int fu_dtype, const void* f_graph, int fg_dtype, const void* y, int y_dtype,
                                int B, int N, int D, float margin, void* work, float* per_image, float* loss,
                                mg_stream_t stream) {
  MG_REQUIRE(f_unet && f_graph && y && work && loss && B > 0 && N > 0 && D > 0, MG_ERR_INVALID,
             "mg_feature_consistency_loss: bad arguments");
  MG_REQUIRE(B <= 65535, MG_ERR_INVALID, "mg_feature_consistency_loss: batch too large");
  MG_REQUIRE(y_dtype >= MG_F32 && y_dtype <= MG_I64, MG_ERR_INVALID, "mg_feature_consistency_loss: bad label dtype %d", y_dtype);
  cudaStream_t st = (cudaStream_t)stream;
#define MG_FL(TU, TG) return launch_feature_loss<TU, TG>(f_unet, f_graph, y, y_dtype, B, N, D, margin, work, per_image, loss, st)
  if (fu_dtype == MG_F32 && fg_dtype == MG_F32) MG_FL(float, float);
  if (fu_dtype == MG_BF16 && fg_dtype == MG_BF16) MG_FL(__nv_bfloat16, __nv_bfloat16);
  if (fu_dtype == MG_BF16 && fg_dtype == MG_F32) MG_FL(__nv_bfloat16, float);
  if (fu_dtype == MG_F32 && fg_dtype == MG_BF16) MG_FL(float, __nv_bfloat16);
#undef MG_FL
  set_error("mg_feature_consistency_loss: unsupported dtypes %d, %d", fu_dtype, fg_dtype);
  return MG_ERR_INVALID;
}

int mg_feature_consistency_loss_backward(const void* f_unet, int fu_dtype, const void* f_graph, int fg_dtype, const void* y,
                                         int y_dtype, int B, int N, int D, float margin, const float* grad_loss,
                                         float* grad_f_unet, float* grad_f_graph, mg_stream_t stream) {
  MG_REQUIRE(f_unet && f_graph && y && grad_loss && (grad_f_unet || grad_f_graph) && B > 0 && N > 0 && D > 0, MG_ERR_INVALID,
             "mg_feature_consistency_loss_backward: bad arguments");
  MG_REQUIRE(y_dtype >= MG_F32 && y_dtype <= MG_I64, MG_ERR_INVALID, "mg_feature_consistency_loss_backward: bad label dtype %d",
             y_dtype);
  cudaStream_t st = (cudaStream_t)stream;
#define MG_FLB(TU, TG) \
  return launch_feature_loss_bwd<TU, TG>(f_unet, f_graph, y, y_dtype, B, N, D, margin, grad_loss, grad_f_unet, grad_f_graph, st)
  if (fu_dtype == MG_F32 && fg_dtype == MG_F32) MG_FLB(float, float);
  if (fu_dtype == MG_BF16 && fg_dtype == MG_BF16) MG_FLB(__nv_bfloat16, __nv_bfloat16);
  if (fu_dtype == MG_BF16 && fg_dtype == MG_F32) MG_FLB(__nv_bfloat16, float);
  if (fu_dtype == MG_F32 && fg_dtype == MG_BF16) MG_FLB(float, __nv_bfloat16);
#undef MG_FLB
  set_error("mg_feature_consistency_loss_backward: unsupported dtypes %d, %d", fu_dtype, fg_dtype);
  return MG_ERR_INVALID;
}

int64_t mg_tv_loss_work_bytes(int dtype, int B, int C, int H, int W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  // alignment of x is unknown here: size for the larger (scalar-path) partial count
  const int xchunks = ceil_div(W, 32);
  return (int64_t)B * C * ceil_div(xchunks, kTvWarps) * ceil_div(H, kTvRows) * kTvWarps * (int64_t)sizeof(float2);
}

int mg_tv_loss(const void* x, int dtype, int B, int C, int H, int W, float weight, void* work, float* out3, mg_stream_t stream) {
  MG_REQUIRE(x && work && out3 && B > 0 && C > 0 && H > 0 && W > 0, MG_ERR_INVALID, "mg_tv_loss: bad arguments");
  MG_REQUIRE(dtype == MG_F32 || dtype == MG_BF16, MG_ERR_INVALID, "mg_tv_loss: unsupported dtype %d", dtype);
  MG_REQUIRE((int64_t)B * C <= 65535, MG_ERR_INVALID, "mg_tv_loss: more than 65535 planes");
  cudaStream_t st = (cudaStream_t)stream;
  const int planes = B * C;
  const TvShape s = tv_shape(dtype, x, planes, H, W);
  dim3 grid(s.xcb * s.strips, planes);
  float2* partial = reinterpret_cast<float2*>(work);
  if (dtype == MG_F32) {
    if (s.V == 4)
      tv_partial_kernel<float, 4><<<grid, kTvWarps * 32, 0, st>>>(reinterpret_cast<const float*>(x), H, W, s.xcb, partial);
    else
      tv_partial_kernel<float, 1><<<grid, kTvWarps * 32, 0, st>>>(reinterpret_cast<const float*>(x), H, W, s.xcb, partial);
  } else {
    if (s.V == 8)
      tv_partial_kernel<__nv_bfloat16, 8><<<grid, kTvWarps * 32, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), H, W, s.xcb,
                                                                          partial);
    else
      tv_partial_kernel<__nv_bfloat16, 1><<<grid, kTvWarps * 32, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), H, W, s.xcb,
                                                                          partial);
  }
  int rc = check_launch("tv_partial_kernel");
  if (rc) return rc;
  // counts as the reference computes them (python ints -> fp32 scalars), train_end_to_end.py:85-86
  const float count_h = (float)((int64_t)(H - 1) * W), count_w = (float)((int64_t)H * (W - 1));
  tv_final_kernel<<<1, 256, 0, st>>>(partial, s.partials, count_h, count_w, weight, (float)B, out3);
  return check_launch("tv_final_kernel");
}

int mg_tv_loss_backward(const void* x, int dtype, int B, int C, int H, int W, float weight, const float* grad_loss, float* grad_x,
                        mg_stream_t stream) {
  MG_REQUIRE(x && grad_loss && grad_x && B > 0 && C > 0 && H > 0 && W > 0, MG_ERR_INVALID, "mg_tv_loss_backward: bad arguments");
  MG_REQUIRE(dtype == MG_F32 || dtype == MG_BF16, MG_ERR_INVALID, "mg_tv_loss_backward: unsupported dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)B * C * H * W;
  const float count_h = (float)((int64_t)(H - 1) * W), count_w = (float)((int64_t)H * (W - 1));
  // an empty direction has no elements to differentiate: its term contributes 0
  const float sh = H > 1 ? 2.f * weight / (count_h * (float)B) : 0.f;
  const float sw = W > 1 ? 2.f * weight / (count_w * (float)B) : 0.f;
  const int grid = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 32);
  if (dtype == MG_F32)
    tv_backward_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), total, H, W, sh, sw, grad_loss, grad_x);
  else
    tv_backward_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), total, H, W, sh, sw,
                                                            grad_loss, grad_x);
  return check_launch("tv_backward_kernel");
}

}  // extern "C"
