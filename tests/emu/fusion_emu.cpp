// TEST INFRASTRUCTURE — host execution of csrc/fusion_kernels.cuh (see cuda_host_shim.h): the harness reproduces
// launch_region_gather's choice of kernel and launch shape (csrc/fusion.cu) and then calls the kernel function once per
// (block, thread).  Built by tests/test_fusion_emulation.py with g++.
#include "cuda_host_shim.h"

#include "../../mingraph_unet_b200/csrc/fusion_kernels.cuh"

using namespace mg;

template <typename TO, typename TM>
static int run(const float* table, int R, int D, const void* map, int B, int H, int W, void* out, int64_t stride, int sms,
               int force_scalar) {
  constexpr int VEC = FuPack<TO>::VEC;
  const bool vec_ok = (W % VEC == 0) && (D % 4 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)map % 16 == 0) &&
                      ((uintptr_t)table % 16 == 0) && ((stride * (int64_t)sizeof(TO)) % 16 == 0);
  if (vec_ok && !force_scalar) {
    const int dchunk = fusion_vec_dchunk(B, D, H, W, VEC, sms);
    blockDim = dim3(kFuTX, kFuTY, 1);
    gridDim = dim3(ceil_div(W / VEC, kFuTX), ceil_div(H, kFuRY), B * ceil_div(D, dchunk));
    for (unsigned bz = 0; bz < gridDim.z; ++bz)
      for (unsigned by = 0; by < gridDim.y; ++by)
        for (unsigned bx = 0; bx < gridDim.x; ++bx)
          for (unsigned ty = 0; ty < blockDim.y; ++ty)
            for (unsigned tx = 0; tx < blockDim.x; ++tx) {
              blockIdx = {bx, by, bz};
              threadIdx = {tx, ty, 0};
              region_map_gather_vec_kernel<TO, TM>(table, R, D, reinterpret_cast<const TM*>(map), H, W,
                                                   reinterpret_cast<TO*>(out), stride, dchunk);
            }
    return 1;
  }
  const int64_t total = (int64_t)B * D * H * W;
  blockDim = dim3(256, 1, 1);
  gridDim = dim3((unsigned)std::min<int64_t>(ceil_div64(total, 256), (int64_t)sms * 32), 1, 1);
  for (unsigned bx = 0; bx < gridDim.x; ++bx)
    for (unsigned tx = 0; tx < blockDim.x; ++tx) {
      blockIdx = {bx, 0, 0};
      threadIdx = {tx, 0, 0};
      region_map_gather_scalar_kernel<TO, TM>(table, R, D, reinterpret_cast<const TM*>(map), B, H, W,
                                              reinterpret_cast<TO*>(out), stride);
    }
  return 0;
}

// returns 1 if the vector kernel ran, 0 for the scalar kernel, -1 on a detected fault (misaligned / out-of-range access)
extern "C" int emu_region_map_gather(const float* table, int R, int D, const void* map, int map_is_i64, int B, int H, int W,
                                     void* out, int out_is_bf16, int64_t stride, const void* out_lo, const void* out_hi,
                                     int sms, int force_scalar) {
  emu_lo = (const char*)out_lo;
  emu_hi = (const char*)out_hi;
  emu_faults = 0;
  int r;
  if (!out_is_bf16 && !map_is_i64) r = run<float, int32_t>(table, R, D, map, B, H, W, out, stride, sms, force_scalar);
  else if (!out_is_bf16) r = run<float, long long>(table, R, D, map, B, H, W, out, stride, sms, force_scalar);
  else if (!map_is_i64) r = run<__nv_bfloat16, int32_t>(table, R, D, map, B, H, W, out, stride, sms, force_scalar);
  else r = run<__nv_bfloat16, long long>(table, R, D, map, B, H, W, out, stride, sms, force_scalar);
  return emu_faults ? -1 : r;
}
