// Launch code and C ABI of the per-region fusion gather; the kernels live in fusion_kernels.cuh (see there).
#include <algorithm>

#include "fusion_kernels.cuh"

namespace mg {

template <typename TO, typename TM>
static int launch_region_gather(const float* table, int R, int D, const void* map, int B, int H, int W, void* out,
                                int64_t stride, cudaStream_t st) {
  constexpr int VEC = FuPack<TO>::VEC;
  const bool vec_ok = (W % VEC == 0) && (D % 4 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)map % 16 == 0) &&
                      ((uintptr_t)table % 16 == 0) && ((stride * (int64_t)sizeof(TO)) % 16 == 0);
  if (vec_ok) {
    const int dchunk = fusion_vec_dchunk(B, D, H, W, VEC, num_sms());
    dim3 block(kFuTX, kFuTY);
    dim3 grid(ceil_div(W / VEC, kFuTX), ceil_div(H, kFuRY), B * ceil_div(D, dchunk));
    MG_REQUIRE(grid.z <= 65535 && grid.y <= 65535, MG_ERR_INVALID, "mg_region_map_gather: grid too large");
    region_map_gather_vec_kernel<TO, TM><<<grid, block, 0, st>>>(table, R, D, reinterpret_cast<const TM*>(map), H, W,
                                                                 reinterpret_cast<TO*>(out), stride, dchunk);
    return check_launch("region_map_gather_vec_kernel");
  }
  const int64_t total = (int64_t)B * D * H * W;
  const int grid = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 32);
  region_map_gather_scalar_kernel<TO, TM><<<grid, 256, 0, st>>>(table, R, D, reinterpret_cast<const TM*>(map), B, H, W,
                                                                reinterpret_cast<TO*>(out), stride);
  return check_launch("region_map_gather_scalar_kernel");
}

}  // namespace mg

using namespace mg;

extern "C" {

int mg_region_map_gather(const float* table, int R, int D, const void* map, int map_dtype, int B, int H, int W, void* out,
                         int out_dtype, int64_t out_batch_stride, mg_stream_t stream) {
  MG_REQUIRE(table && map && out && R > 0 && D > 0 && B > 0 && H > 0 && W > 0, MG_ERR_INVALID,
             "mg_region_map_gather: bad arguments");
  MG_REQUIRE(out_batch_stride >= (int64_t)D * H * W, MG_ERR_INVALID, "mg_region_map_gather: batch stride smaller than D*H*W");
  MG_REQUIRE(map_dtype == MG_I32 || map_dtype == MG_I64, MG_ERR_INVALID, "mg_region_map_gather: map must be int32 or int64");
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == MG_F32 && map_dtype == MG_I32)
    return launch_region_gather<float, int32_t>(table, R, D, map, B, H, W, out, out_batch_stride, st);
  if (out_dtype == MG_F32 && map_dtype == MG_I64)
    return launch_region_gather<float, long long>(table, R, D, map, B, H, W, out, out_batch_stride, st);
  if (out_dtype == MG_BF16 && map_dtype == MG_I32)
    return launch_region_gather<__nv_bfloat16, int32_t>(table, R, D, map, B, H, W, out, out_batch_stride, st);
  if (out_dtype == MG_BF16 && map_dtype == MG_I64)
    return launch_region_gather<__nv_bfloat16, long long>(table, R, D, map, B, H, W, out, out_batch_stride, st);
  set_error("mg_region_map_gather: unsupported out dtype %d", out_dtype);
  return MG_ERR_INVALID;
}

}  // extern "C"
