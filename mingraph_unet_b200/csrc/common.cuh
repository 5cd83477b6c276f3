// Shared device/host helpers for libmingraph_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mingraph_b200.h"

namespace mg {

// ---- error plumbing (C ABI returns int; message kept per host thread) -----------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError() -> MG_ERR_CUDA

#define MG_REQUIRE(cond, code, ...)              \
  do {                                           \
    if (!(cond)) {                               \
      mg::set_error(__VA_ARGS__);                \
      return (code);                             \
    }                                            \
  } while (0)

int num_sms();                        // cached cudaDevAttrMultiProcessorCount (148 on B200)

// deterministic two-stage per-label row reduction (pool_unpool.cu): out (B,K,D) = mean or scale*sum of h (B,N,D) rows by label
int64_t segment_work_bytes(int B, int N, int D, int K);
int segment_reduce_launch(const float* h, const int32_t* labels, int B, int N, int D, int K, int mean, float scale, float* out,
                          int32_t* counts, void* work, cudaStream_t st);

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- element load/store with fp32 math ---------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Load V consecutive elements (V in {1,2,4}) as fp32 through the read-only path.
template <typename T, int V>
struct VecLoad;
template <>
struct VecLoad<float, 1> {
  static __device__ __forceinline__ void ld(const float* p, float* o) { o[0] = __ldg(p); }
};
template <>
struct VecLoad<float, 2> {
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    float2 v = __ldg(reinterpret_cast<const float2*>(p));
    o[0] = v.x; o[1] = v.y;
  }
};
template <>
struct VecLoad<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <>
struct VecLoad<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* o) {
    o[0] = __bfloat162float(__ldg(p));
  }
};
template <>
struct VecLoad<__nv_bfloat16, 2> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* o) {
    unsigned u = __ldg(reinterpret_cast<const unsigned*>(p));
    o[0] = __uint_as_float(u << 16);
    o[1] = __uint_as_float(u & 0xffff0000u);
  }
};
template <>
struct VecLoad<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* o) {
    uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    o[0] = __uint_as_float(u.x << 16);
    o[1] = __uint_as_float(u.x & 0xffff0000u);
    o[2] = __uint_as_float(u.y << 16);
    o[3] = __uint_as_float(u.y & 0xffff0000u);
  }
};

// ---- warp reductions ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// Order-preserving float atomic max (deterministic: max is order independent).
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (!(__float_as_uint(v) >> 31))        // branch on the sign bit: -0.0f must take the unsigned-min branch
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

__device__ __forceinline__ float leaky_relu(float v, float slope) { return v > 0.f ? v : v * slope; }
__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }

// Packed fp32x2 FMA (sm_100 FFMA2): two independent IEEE fmaf results per instruction, bit-identical to the scalar form.
// A pair built from the same scalar twice is folded by ptxas into the instruction's scalar-broadcast operand.
#ifndef MG_HOST_EMULATION
__device__ __forceinline__ unsigned long long f32x2_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
  return r;
}
__device__ __forceinline__ void f32x2_unpack(unsigned long long v, float& lo, float& hi) {
  unsigned a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
  lo = __uint_as_float(a); hi = __uint_as_float(b);
}
__device__ __forceinline__ void f32x2_fma(unsigned long long& d, unsigned long long a, unsigned long long b) {      // d = a * b + d, per lane
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
#endif

// Streaming (evict-first) 16-byte store for write-once outputs.
#ifndef MG_HOST_EMULATION
__device__ __forceinline__ void st_cs_v4(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
#else   // tests/emu: the kernels compiled for the host; the shim header provides an alignment-checking st_cs_v4
#endif

}  // namespace mg
