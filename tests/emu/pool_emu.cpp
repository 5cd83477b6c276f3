// TEST INFRASTRUCTURE — host execution of csrc/pool_tma_kernels.cuh (see cuda_warp_shim.h): static and dynamic strip
// scheduling of pool_patches_tma_kernel and the tensor-core variant pool_patches_mma_kernel, with the launch parameters
// launch_pool (csrc/pool_unpool.cu) computes.  Built by tests/test_pool_emulation.py with g++.
#include "cuda_warp_shim.h"

#include "../../mingraph_unet_b200/csrc/pool_tma_kernels.cuh"

using namespace mg;

template <typename TX, typename TO>
static int run(const void* x, int B, int C, int Hf, int Wf, int ph, int pw, void* out, int sms, int variant, int stages_req,
               int chunk_req, int* counters) {
  const int Hp = ceil_div(Hf, ph), Wp = ceil_div(Wf, pw);
  const int row_bytes = Wf * (int)sizeof(TX);
  int stages = std::max(2, std::min(8, stages_req));
  int stage_bytes = std::max(row_bytes, std::max(1024, chunk_req)) / 128 * 128;
  stage_bytes = std::max(stage_bytes, (row_bytes + 127) / 128 * 128);
  while (stages > 2 && (size_t)kPtWarps * stages * (stage_bytes + 8) + kPtWarps * kPtFifo * 4 > (size_t)kPtSmemBytes) --stages;
  const size_t smem = (size_t)kPtWarps * stages * (stage_bytes + 8) + kPtWarps * kPtFifo * 4;
  if (smem > (size_t)kPtSmemBytes) return -2;
  PoolTmaArgs A;
  A.x = x; A.out = out; A.C = C; A.Hf = Hf; A.Wf = Wf; A.ph = ph; A.pw = pw; A.Hp = Hp; A.Wp = Wp;
  A.rpc = std::max(1, std::min(ph, stage_bytes / row_bytes));
  A.nstrips = B * Hp * C;
  A.stages = stages;
  A.stage_bytes = stage_bytes;
  A.counters = variant == 1 ? counters : nullptr;
  const int grid = std::min(sms, ceil_div(A.nstrips, kPtWarps));
  for (int bx = 0; bx < grid; ++bx) {
    if (variant == 0) emu_run_block(bx, grid, kPtWarps * 32, [&]() { pool_patches_tma_kernel<TX, TO, false>(A); });
    else if (variant == 1) emu_run_block(bx, grid, kPtWarps * 32, [&]() { pool_patches_tma_kernel<TX, TO, true>(A); });
    else if constexpr (std::is_same<TX, __nv_bfloat16>::value)
      emu_run_block(bx, grid, kPtWarps * 32, [&]() { pool_patches_mma_kernel<TO>(A); });
    else return -3;
  }
  return grid;
}

// variant: 0 static strips, 1 dynamic strips (counters: 2 ints, zero before the first launch), 2 tensor-core summation.
// returns the grid size, or -1 on a detected fault (misaligned / out-of-range access, mbarrier misuse)
extern "C" int emu_pool(const void* x, int x_is_bf16, long long x_bytes, int B, int C, int Hf, int Wf, int ph, int pw, void* out,
                        int out_is_bf16, int sms, int variant, int stages, int chunk, int* counters, long long* copied) {
  emu_src_lo = (const char*)x;
  emu_src_hi = (const char*)x + x_bytes;
  emu_faults = 0;
  emu_copied_bytes = 0;
  int r;
  if (x_is_bf16 && out_is_bf16) r = run<__nv_bfloat16, __nv_bfloat16>(x, B, C, Hf, Wf, ph, pw, out, sms, variant, stages, chunk, counters);
  else if (x_is_bf16) r = run<__nv_bfloat16, float>(x, B, C, Hf, Wf, ph, pw, out, sms, variant, stages, chunk, counters);
  else if (out_is_bf16) r = run<float, __nv_bfloat16>(x, B, C, Hf, Wf, ph, pw, out, sms, variant, stages, chunk, counters);
  else r = run<float, float>(x, B, C, Hf, Wf, ph, pw, out, sms, variant, stages, chunk, counters);
  if (copied) *copied = emu_copied_bytes.load();
  return emu_faults ? -1 : r;
}
