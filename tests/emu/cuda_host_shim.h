// TEST INFRASTRUCTURE — lets g++ compile device code of libmingraph_b200 (kernels without barriers / shuffles on their
// path) and run it on the host: one call of the kernel function per (block, thread) with the built-in index variables
// set by the harness.  Vector loads and stores CHECK THEIR ALIGNMENT (a misaligned 16-byte access is a fault on the
// GPU but silently works on x86), and every store is bounds-checked against the registered output range.
#pragma once
#define MG_HOST_EMULATION 1
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <cuda_runtime.h>      // vector types, dim3, host-side declarations only (g++ defines no __CUDACC__)
#include <cuda_bf16.h>
#undef __launch_bounds__
#define __launch_bounds__(...)

struct EmuIdx { unsigned x, y, z; };
static thread_local EmuIdx threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

static const char* emu_lo = nullptr;          // registered writable range
static const char* emu_hi = nullptr;
static long long emu_faults = 0;

static inline void emu_check(const void* p, size_t align, const char* what) {
  if (((uintptr_t)p) % align != 0) {
    fprintf(stderr, "emu: misaligned %s of %zu bytes at %p\n", what, align, p);
    ++emu_faults;
  }
}

template <typename T>
static inline T __ldg(const T* p) {
  emu_check(p, sizeof(T) > 16 ? 16 : sizeof(T), "load");
  return *p;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
#ifndef MG_EMU_WARP   // (cuda_warp_shim.h provides real lock-step versions)
// common.cuh helpers that are compiled but not executed by the emulated kernels
static inline float __shfl_xor_sync(unsigned, float v, int, int = 32) { abort(); return v; }
#endif
static inline int atomicMax(int* a, int v) { int o = *a; if (v > o) *a = v; return o; }
static inline unsigned atomicMin(unsigned* a, unsigned v) { unsigned o = *a; if (v < o) *a = v; return o; }

namespace mg {
static inline void st_cs_v4(void* p, uint4 v) {
  emu_check(p, 16, "store");
  if ((const char*)p < emu_lo || (const char*)p + 16 > emu_hi) {
    fprintf(stderr, "emu: store outside the output buffer at %p\n", p);
    ++emu_faults;
    return;
  }
  memcpy(p, &v, 16);
}
}  // namespace mg
