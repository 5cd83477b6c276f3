// Backward of the block's glue stages ("the matching backward scatter" of the forward's gathers):
//   * un-pool      f_g[b,d,y,x] = table[b, labels[b, src(y,x)], d]   (scripts/train_end_to_end.py:403-421)
//       grad_table[b,k,d] = sum over the pixels whose patch carries label k  — a segmented reduction of the
//       dense (B,D,H,W) gradient: per-patch window sums (HBM-streaming), then a per-image fixed-order
//       reduction of the N patch rows by label.  No atomics, deterministic.
//   * region pool  R[b,k,:] = mean(h[b, labels==k, :])                  (train_end_to_end.py:368-373)
//       grad_h[b,n,:] = grad_R[b, labels[n], :] / count[b, labels[n]]
//   * row softmax  S = softmax(logits, 1)                               (mincut_refinement.py:193)
//       grad_logits = S * (grad_S - sum_c grad_S_c S_c)
// The reference gets all three from autograd (IndexBackward / UpsampleNearest2DBackward / SoftmaxBackward).
#include "common.cuh"

namespace mg {

__device__ __forceinline__ int nearest_src_bw(int dst, int in_size, int out_size, float scale) {
  if (out_size == in_size) return dst;
  if (out_size == 2 * in_size) return dst >> 1;
  const int s = (int)floorf(__fmul_rn((float)dst, scale));
  return min(s, in_size - 1);
}

// first destination index whose source index is >= p (src is monotone non-decreasing in dst)
__device__ __forceinline__ int first_dst_of(int p, int in_size, int out_size, float scale) {
  if (p <= 0) return 0;
  if (p >= in_size) return out_size;
  int lo = 0, hi = out_size;                     // invariant: src(lo-1) < p <= src(hi)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (nearest_src_bw(mid, in_size, out_size, scale) >= p) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// generic window sums: thread per (b, n, d); d fastest so the (B,N,D) store is coalesced.
template <typename TG>
__global__ void unpool_bwd_window_kernel(const TG* __restrict__ g, int64_t batch_stride, int B, int D, int Hp, int Wp, int H,
                                         int W, float sy, float sx, float* __restrict__ gp) {
  const int64_t total = (int64_t)B * Hp * Wp * D;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(t % D);
    const int n = (int)((t / D) % (Hp * Wp));
    const int b = (int)(t / ((int64_t)D * Hp * Wp));
    const int py = n / Wp, px = n - py * Wp;
    const int y0 = first_dst_of(py, Hp, H, sy), y1 = first_dst_of(py + 1, Hp, H, sy);
    const int x0 = first_dst_of(px, Wp, W, sx), x1 = first_dst_of(px + 1, Wp, W, sx);
    const TG* plane = g + (size_t)b * batch_stride + (size_t)d * H * W;
    float acc = 0.f;
    for (int y = y0; y < y1; ++y) {
      float row = 0.f;
      for (int x = x0; x < x1; ++x) row += to_f32<TG>(plane[(size_t)y * W + x]);
      acc += row;
    }
    gp[t] = acc;
  }
}

// grad_h[b,n,:] (+)= grad_R[b, l, :] / count[b, l]
__global__ void segment_mean_bwd_kernel(const float* __restrict__ gR, const int32_t* __restrict__ labels,
                                        const int32_t* __restrict__ counts, int B, int N, int D, int K, int accumulate,
                                        float* __restrict__ gh) {
  const int64_t total = (int64_t)B * N * D;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(t % D);
    const int64_t bn = t / D;
    const int b = (int)(bn / N);
    const int l = __ldg(labels + bn);
    float v = 0.f;
    if (l >= 0 && l < K) {
      const int c = __ldg(counts + (size_t)b * K + l);
      if (c > 0) v = __ldg(gR + ((size_t)b * K + l) * D + d) / (float)c;
    }
    gh[t] = accumulate ? gh[t] + v : v;
  }
}

__global__ void softmax_bwd_kernel(const float* __restrict__ S, const float* __restrict__ gS, int N, int K,
                                   float* __restrict__ gl) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    const float* s = S + (size_t)n * K;
    const float* g = gS + (size_t)n * K;
    float dot = 0.f;
    for (int k = 0; k < K; ++k) dot = fmaf(__ldg(g + k), __ldg(s + k), dot);
    for (int k = 0; k < K; ++k) gl[(size_t)n * K + k] = __ldg(s + k) * (__ldg(g + k) - dot);
  }
}

}  // namespace mg

using namespace mg;

extern "C" {

int64_t mg_unpool_backward_work_bytes(int B, int D, int Hp, int Wp, int K) {
  if (B <= 0 || D <= 0 || Hp <= 0 || Wp <= 0 || K <= 0) return 0;
  return (((int64_t)B * Hp * Wp * D * 4 + 255) & ~(int64_t)255) + segment_work_bytes(B, Hp * Wp, D, K);
}

int mg_unpool_nearest_backward(const void* grad_out, int grad_dtype, int64_t grad_batch_stride, const int32_t* labels, int B,
                               int K, int D, int Hp, int Wp, int H, int W, void* work, float* grad_table,
                               mg_stream_t stream) {
  MG_REQUIRE(grad_out && work && grad_table && B > 0 && K > 0 && D > 0 && Hp > 0 && Wp > 0 && H > 0 && W > 0, MG_ERR_INVALID,
             "mg_unpool_nearest_backward: bad arguments");
  MG_REQUIRE(labels || K == Hp * Wp, MG_ERR_INVALID, "mg_unpool_nearest_backward: labels==NULL requires K == Hp*Wp");
  MG_REQUIRE(grad_dtype == MG_F32 || grad_dtype == MG_BF16, MG_ERR_INVALID, "mg_unpool_nearest_backward: dtype");
  MG_REQUIRE(grad_batch_stride >= (int64_t)D * H * W, MG_ERR_INVALID, "mg_unpool_nearest_backward: batch stride");
  const size_t smem = (size_t)8 * K * D * 4;
  MG_REQUIRE(labels == nullptr || smem <= 200 * 1024, MG_ERR_UNSUPPORTED, "mg_unpool_nearest_backward: K*D=%d too large", K * D);
  cudaStream_t st = (cudaStream_t)stream;
  float* gp = reinterpret_cast<float*>(work);
  const int N = Hp * Wp;
  float scale = 1.f;
  // windows are exact ph x pw tiles when the sizes divide by a power of two (the fp32 scale in/out of torch's
  // nearest rule is then exact and maps y -> y / ph); anything else takes the generic kernel, which evaluates
  // the same index rule as the forward
  auto pow2_ratio = [](int big, int small) { const int r = big / small; return big % small == 0 && (r & (r - 1)) == 0; };
  const bool tiles = labels && pow2_ratio(H, Hp) && pow2_ratio(W, Wp) && grad_batch_stride == (int64_t)D * H * W;
  if (tiles) {
    const int ph = H / Hp, pw = W / Wp;
    const int rc = mg_pool_patches(grad_out, grad_dtype, B, D, H, W, ph, pw, gp, MG_F32, stream);   // window MEANS (B,N,D)
    if (rc) return rc;
    scale = (float)(ph * pw);
  } else {
    const float sy = (float)Hp / (float)H, sx = (float)Wp / (float)W;
    const int64_t total = (int64_t)B * N * D;
    const int grid = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 16);
    if (grad_dtype == MG_F32)
      unpool_bwd_window_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(grad_out), grad_batch_stride, B, D,
                                                           Hp, Wp, H, W, sy, sx, gp);
    else
      unpool_bwd_window_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(grad_out),
                                                                   grad_batch_stride, B, D, Hp, Wp, H, W, sy, sx, gp);
    const int rc = check_launch("unpool_bwd_window_kernel");
    if (rc) return rc;
  }
  if (!labels) {          // identity labels: the table IS the per-patch matrix (window sums, scale 1)
    if (cudaMemcpyAsync(grad_table, gp, (size_t)B * N * D * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("mg_unpool_nearest_backward: copy failed");
      return MG_ERR_CUDA;
    }
    return MG_OK;
  }
  unsigned char* seg_work = reinterpret_cast<unsigned char*>(work) + (((size_t)B * N * D * 4 + 255) & ~(size_t)255);
  return segment_reduce_launch(gp, labels, B, N, D, K, 0, scale, grad_table, nullptr, seg_work, st);
}

int mg_segment_mean_backward(const float* grad_out, const int32_t* labels, const int32_t* counts, int B, int N, int D, int K,
                             int accumulate, float* grad_h, mg_stream_t stream) {
  MG_REQUIRE(grad_out && labels && counts && grad_h && B > 0 && N > 0 && D > 0 && K > 0, MG_ERR_INVALID,
             "mg_segment_mean_backward: bad arguments");
  const int64_t total = (int64_t)B * N * D;
  const int grid = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 16);
  segment_mean_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grad_out, labels, counts, B, N, D, K, accumulate ? 1 : 0,
                                                                 grad_h);
  return check_launch("segment_mean_bwd_kernel");
}

int mg_softmax_backward(const float* S, const float* grad_S, int N, int K, float* grad_logits, mg_stream_t stream) {
  MG_REQUIRE(S && grad_S && grad_logits && N > 0 && K > 0, MG_ERR_INVALID, "mg_softmax_backward: bad arguments");
  const int grid = std::min(ceil_div(N, 256), num_sms() * 8);
  softmax_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(S, grad_S, N, K, grad_logits);
  return check_launch("softmax_bwd_kernel");
}

}  // extern "C"
