// Hardware probe (standalone; nvcc -gencode arch=compute_100a,code=sm_100a -o gather4_probe gather4_probe.cu):
//  A. layout / out-of-range behaviour of cp.async.bulk.tensor.2d ... tile::gather4 with a 128-byte-swizzled map;
//  B. row-gather rate per GPU of (i) TMA gather4 and (ii) cp.async 16-byte copies, same loop shape as the aggregation
//     kernel's warps (12 warps per SM, 32 random 128-byte rows per item, 2 stages per warp).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
  return (EncodeTiledFn)p;
}
static bool make_map(CUtensorMap* m, void* base, uint64_t rows, uint32_t box_rows, CUtensorMapSwizzle sw) {
  cuuint64_t dims[2] = {64, rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("  encode(box_rows=%u, swizzle=%d) failed: %d\n", box_rows, (int)sw, (int)r);
  return r == CUDA_SUCCESS;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {      // bounded: a lost completion traps instead of hanging the GPU
  for (int spin = 0; spin < (1 << 22); ++spin) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, int c, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(c), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

// ---- A: one gather4, dump where each (row, 16-byte chunk) landed ----
__global__ void layout_kernel(const __grid_constant__ CUtensorMap map, int r0, int r1, int r2, int r3, float* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  const uint32_t b = s32(&bar);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 0xff;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(b) : "memory");
    gather4(s32(sm), &map, 0, r0, r1, r2, r3, b);
    mbar_wait(b, 0);
  }
  __syncthreads();
  // 32 chunks of 16 bytes: first bf16 of each
  if (threadIdx.x < 32) out[threadIdx.x] = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(sm + threadIdx.x * 16));
}

// ---- B: gather-rate loops ----
constexpr int kWarps = 12, kStage = 4096, kStages = 2;
template <int MODE>   // 0: TMA gather4, 1: cp.async
__global__ void __launch_bounds__(kWarps * 32, 1) rate_kernel(const __grid_constant__ CUtensorMap map, const __nv_bfloat16* x,
                                                              const int* __restrict__ idx, int items_total, unsigned* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[kWarps][kStages];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ring = s32(sm + warp * kStages * kStage);
  if (lane == 0)
    for (int s = 0; s < kStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[warp][s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int gw = blockIdx.x * kWarps + warp, nw = gridDim.x * kWarps;
  unsigned acc = 0;
  const int r4 = lane >> 3, ch8 = lane & 7;
  auto issue = [&](int item, int stg) {
    const int s = __ldg(idx + (size_t)item * 32 + lane);
    if (MODE == 0) {
      const int s1 = __shfl_down_sync(0xffffffffu, s, 1), s2 = __shfl_down_sync(0xffffffffu, s, 2), s3 = __shfl_down_sync(0xffffffffu, s, 3);
      const uint32_t bar = s32(&bars[warp][stg]);
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4096;" ::"r"(bar) : "memory");
      __syncwarp();
      if ((lane & 3) == 0) gather4(ring + stg * kStage + lane * 128, &map, 0, s, s1, s2, s3, bar);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int sn = __shfl_sync(0xffffffffu, s, i * 4 + r4);
        const uint32_t dst = ring + stg * kStage + (i * 4 + r4) * 128 + ((ch8 ^ ((i * 4 + r4) & 7)) << 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"((const char*)x + (size_t)sn * 128 + ch8 * 16) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  int n = 0;
  int item = gw;
  if (item < items_total) issue(item, 0);
  for (; item < items_total; item += nw, ++n) {
    const int stg = n & 1;
    if (item + nw < items_total) issue(item + nw, stg ^ 1);
    else if (MODE == 1) asm volatile("cp.async.commit_group;" ::: "memory");
    if (MODE == 0) mbar_wait(s32(&bars[warp][stg]), (uint32_t)((n >> 1) & 1));
    else asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    const uint4* st = reinterpret_cast<const uint4*>(sm + warp * kStages * kStage + stg * kStage);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 v = st[i * 32 + lane];
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    __syncwarp();
  }
  if (acc == 0x12345678u) out[0] = acc;
}

int main() {
  const int N = 262144;
  std::vector<__nv_bfloat16> hx((size_t)N * 64);
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < 64; ++c) hx[(size_t)r * 64 + c] = __float2bfloat16((float)((r % 8) * 8 + c / 8));   // row digit, chunk digit (exact in bf16)
  __nv_bfloat16* dx; cudaMalloc(&dx, hx.size() * 2); cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  float* dout; cudaMalloc(&dout, 4096);
  float hout[32];
  for (int box_rows : {1}) {
    for (int swi = 0; swi < 2; ++swi) {
      CUtensorMapSwizzle sw = swi ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
      alignas(64) CUtensorMap map;
      printf("== box_rows %d swizzle %s\n", box_rows, swi ? "128B" : "none");
      if (!make_map(&map, dx, N, box_rows, sw)) continue;
      int rows[4] = {5, 1001, 2, 7};
      layout_kernel<<<1, 32, 4096>>>(map, rows[0], rows[1], rows[2], rows[3], dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  layout kernel: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hout, dout, 128, cudaMemcpyDeviceToHost);
      for (int r = 0; r < 4; ++r) {
        printf("  smem row %d:", r);
        for (int c = 0; c < 8; ++c) printf(" %8.3f", hout[r * 8 + c]);
        printf("\n");
      }
    }
  }
  // B
  const int items = N * 8 / 32;            // 2.1 M rows
  std::vector<int> hidx((size_t)items * 32);
  srand(1);
  for (auto& v : hidx) v = (int)(((unsigned)rand() * 32768u + (unsigned)rand()) % (unsigned)N);
  int* didx; cudaMalloc(&didx, hidx.size() * 4); cudaMemcpy(didx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice);
  unsigned* dacc; cudaMalloc(&dacc, 4);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int box_rows : {1}) {
    alignas(64) CUtensorMap map;
    if (!make_map(&map, dx, N, box_rows, CU_TENSOR_MAP_SWIZZLE_128B)) continue;
    const size_t smem = kWarps * kStages * kStage + 1024;
    cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 2; ++mode) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        if (mode == 0) rate_kernel<0><<<sms, kWarps * 32, smem>>>(map, dx, didx, items, dacc);
        else rate_kernel<1><<<sms, kWarps * 32, smem>>>(map, dx, didx, items, dacc);
        cudaEventRecord(b);
        cudaError_t e = cudaEventSynchronize(b);
        if (e != cudaSuccess) { printf("rate kernel mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
      }
      printf("rate box_rows %d %s: %.1f us for %d rows of 128 B = %.0f GB/s\n", box_rows, mode == 0 ? "TMA gather4" : "cp.async   ", best * 1e3,
             items * 32, (double)items * 32 * 128 / (best * 1e-3) / 1e9);
    }
  }
  return 0;
}
