// FeatureFusion.forward, per-region branch — model/fusion_detection/feature_fusion.py:81-140 of the reference:
//   f_g_pixel[b, :, y, x] = f_g[region_to_pixel_map[b, y, x]]   if 0 <= map < R   (valid_mask, :119)
//                         = 0                                     otherwise         (:85, :140)
// i.e. an embedding gather from a small (R, D) table to a dense (B, D, H, W) map — the same HBM-write-bound shape as
// the nearest un-pool (pool_unpool.cu), with one label per PIXEL instead of one per patch.  The result is written
// straight into a channel slice of the caller's fusion buffer (out_batch_stride), so `Concat(F_u, F_g)` (:143) costs no
// extra pass over the block's dominant tensor.
//
// Vector kernel: a thread owns VEC consecutive pixels of a row (VEC * sizeof(out) = 16 bytes), reads their labels with
// 16-byte loads once, and walks the channels: table rows come in as float4 (4 channels per load, L1/L2 resident — the
// table is R*D*4 bytes), every channel plane gets one 16-byte streaming store.  When the VEC labels agree (the common
// case inside a region) one table read serves all VEC pixels.
//
// (device code: kernels only, so that tests/emu can compile it for the host; launch code in fusion.cu)
//
// STATUS: written after the round-1 GPU budget was spent — NOT yet run on hardware; its GPU tests are opt-in
// (MG_TEST_UNVERIFIED=1).  Checked on the CPU two ways: a Python model of the index arithmetic
// (tests/test_oracle_fusion.py) and THIS source compiled for the host and executed block by block, thread by thread,
// with alignment-checking loads and stores (tests/emu/, tests/test_fusion_emulation.py).
#pragma once
#include "common.cuh"

namespace mg {

template <typename TO>
struct FuPack;
template <>
struct FuPack<float> {
  static constexpr int VEC = 4;
  static __device__ __forceinline__ uint4 make(const float* v) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
};
template <>
struct FuPack<__nv_bfloat16> {
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

// VEC consecutive labels with 16-byte loads; a label outside [0, R) becomes -1 (feature_fusion.py:119)
template <typename TM, int VEC>
struct LabelLoad;
template <int VEC>
struct LabelLoad<int32_t, VEC> {
  static __device__ __forceinline__ void ld(const int32_t* p, int R, int* lab) {
#pragma unroll
    for (int i = 0; i < VEC / 4; ++i) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(p) + i);
      const int t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) lab[4 * i + j] = (t[j] >= 0 && t[j] < R) ? t[j] : -1;
    }
  }
};
template <int VEC>
struct LabelLoad<long long, VEC> {
  static __device__ __forceinline__ void ld(const long long* p, int R, int* lab) {
#pragma unroll
    for (int i = 0; i < VEC / 2; ++i) {
      const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(p) + i);
      lab[2 * i] = (v.x >= 0 && v.x < (long long)R) ? (int)v.x : -1;
      lab[2 * i + 1] = (v.y >= 0 && v.y < (long long)R) ? (int)v.y : -1;
    }
  }
};

__device__ __forceinline__ float f4_get(const float4& t, int i) { return i == 0 ? t.x : i == 1 ? t.y : i == 2 ? t.z : t.w; }

constexpr int kFuTX = 64, kFuTY = 4, kFuRY = 16;

// grid (ceil(W/VEC/kFuTX), ceil(H/kFuRY), B * dchunks); requires W % VEC == 0, D % 4 == 0, dchunk % 4 == 0,
// 16-byte aligned table / map / out and (out_batch_stride * sizeof(TO)) % 16 == 0
template <typename TO, typename TM>
__global__ void __launch_bounds__(kFuTX* kFuTY) region_map_gather_vec_kernel(const float* __restrict__ table, int R, int D,
                                                                           const TM* __restrict__ map, int H, int W,
                                                                           TO* __restrict__ out, int64_t out_batch_stride,
                                                                           int dchunk) {
  constexpr int VEC = FuPack<TO>::VEC;
  const int dchunks = ceil_div(D, dchunk);
  const int b = blockIdx.z / dchunks, d0 = (blockIdx.z - b * dchunks) * dchunk;
  const int nd4 = min(dchunk, D - d0) >> 2;                   // float4 groups of channels in this chunk
  const int xv = blockIdx.x * kFuTX + threadIdx.x;
  if (xv * VEC >= W) return;
  const int yend = min(H, (int)(blockIdx.y + 1) * kFuRY);
  const TM* mb = map + (size_t)b * H * W;
  TO* ob = out + (size_t)b * out_batch_stride;
  const size_t plane = (size_t)H * W;
  for (int y = blockIdx.y * kFuRY + threadIdx.y; y < yend; y += kFuTY) {
    int lab[VEC];
    LabelLoad<TM, VEC>::ld(mb + (size_t)y * W + (size_t)xv * VEC, R, lab);
    bool same = true;
#pragma unroll
    for (int v = 1; v < VEC; ++v) same = same && (lab[v] == lab[0]);
    TO* orow = ob + ((size_t)d0 * H + y) * W + (size_t)xv * VEC;
    if (same) {
      const float4* tp = reinterpret_cast<const float4*>(table + (size_t)max(lab[0], 0) * D + d0);
      const bool ok = lab[0] >= 0;
      for (int d4 = 0; d4 < nd4; ++d4) {
        const float4 t = ok ? __ldg(tp + d4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
          float vals[VEC];
          const float s = f4_get(t, dd);
#pragma unroll
          for (int v = 0; v < VEC; ++v) vals[v] = s;
          st_cs_v4(orow + (size_t)(4 * d4 + dd) * plane, FuPack<TO>::make(vals));
        }
      }
    } else {
      for (int d4 = 0; d4 < nd4; ++d4) {
        float4 t[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          t[v] = lab[v] >= 0 ? __ldg(reinterpret_cast<const float4*>(table + (size_t)lab[v] * D + d0) + d4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
          float vals[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) vals[v] = f4_get(t[v], dd);
          st_cs_v4(orow + (size_t)(4 * d4 + dd) * plane, FuPack<TO>::make(vals));
        }
      }
    }
  }
}

// any shape / alignment: one output element per thread step
template <typename TO, typename TM>
__global__ void __launch_bounds__(256) region_map_gather_scalar_kernel(const float* __restrict__ table, int R, int D,
                                                                        const TM* __restrict__ map, int B, int H, int W,
                                                                        TO* __restrict__ out, int64_t out_batch_stride) {
  const int64_t plane = (int64_t)H * W, total = (int64_t)B * D * plane;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = t % plane;
    const int d = (int)((t / plane) % D);
    const int64_t b = t / (plane * D);
    const long long l = (long long)__ldg(map + b * plane + pix);
    const float v = (l >= 0 && l < (long long)R) ? __ldg(table + (size_t)l * D + d) : 0.f;
    out[b * out_batch_stride + (int64_t)d * plane + pix] = from_f32<TO>(v);
  }
}

// launch shape of the vector kernel (shared with the host emulation in tests/emu): all channels of a pixel in one block
// when that still fills the machine (labels are then read once), otherwise channel chunks of 32 across blockIdx.z
inline int fusion_vec_dchunk(int B, int D, int H, int W, int vec, int sms) {
  const int64_t blocks_full = (int64_t)ceil_div(W / vec, kFuTX) * ceil_div(H, kFuRY) * B;
  return (blocks_full >= 4 * (int64_t)sms || D <= 32) ? D : 32;
}

}  // namespace mg
