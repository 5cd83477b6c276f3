// Node transform of the GAT layer on the tensor pipe, for the shapes whose weights / z tile do not fit the fused
// kernels (gat_fused_kernel, gat_tc_kernel):  out[n, f] = mean_h / concat_h  ELU( sum_i z[n, h, i] * W[h, f, i] )
// (model/gat/graph_attention.py:53 re-associated: W_h (sum_i alpha x_i), then :118 ELU and :155-158 concat / mean).
//
// z (N, heads, in) fp32 comes from gat_aggregate_kernel.  One CTA computes a 128-row x BN-column output tile for ALL
// heads: per (head, 32-wide K slice) step, 4 producer warps load the z slice and the W slice with 16-byte loads and
// store them as K-major, 128B-swizzled UMMA operands in a 3-stage shared-memory ring; one thread issues
// tcgen05.mma kind::tf32 into the head's own TMEM accumulator block (heads * BN <= 256 columns, so two CTAs share an
// SM); 4 epilogue warps then read the accumulators, apply ELU, sum / concatenate the heads and store.
//
// PASSES = 1: operands used as tf32 (bf16-storage path, tolerance 2e-2).
// PASSES = 3: 3xTF32 — each fp32 operand is split on the fly into hi = tf32(v) and lo = tf32(v - hi) and the tile
//             accumulates hi*hi + lo*hi + hi*lo in fp32: products carry ~21 mantissa bits, which holds the fp32
//             tolerance (1e-5) of the reference while running on the tensor cores.
#include <stdlib.h>

#include "gat_kernels.cuh"

namespace mg {

constexpr int kGmRows = 128;
constexpr int kGmProducerWarps = 4;
constexpr int kGmThreads = (kGmProducerWarps + 1 + 4) * 32;      // producers | MMA issuer | epilogue
constexpr int kGmHeader = 1024;

struct GemmTcArgs {
  const float* z;      // (N, heads, in)
  const float* W;      // (heads, F, in)
  void* out;           // (N, concat ? heads*F : F)
  int N, in_dim, F, heads, concat, out_bf16, BN, stages, tmem_cols;
};

__device__ __forceinline__ uint32_t gm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gm_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void gm_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gm_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "GM_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra GM_WAIT_DONE;\n"
      "bra GM_WAIT_LOOP;\n"
      "GM_WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void gm_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gm_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ uint64_t gm_desc_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void gm_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float gm_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ float gm_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// bytes of one pipeline stage: A (128 rows x 128 B) and B (BN rows x 128 B), times 2 when hi/lo copies are kept
__host__ __device__ inline int gm_stage_bytes(int BN, int passes) { return (kGmRows * 128 + BN * 128) * (passes == 3 ? 2 : 1); }

template <int PASSES>
__global__ void __launch_bounds__(kGmThreads) gat_transform_tc_kernel(const GemmTcArgs A) {
  extern __shared__ unsigned char gm_smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>(((uintptr_t)gm_smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int BN = A.BN, S = A.stages;
  const int stage_bytes = gm_stage_bytes(BN, PASSES);
  const int a_bytes = kGmRows * 128, b_bytes = BN * 128;
  const uint32_t bar_full0 = gm_smem_u32(sm), bar_empty0 = gm_smem_u32(sm + 64), bar_acc = gm_smem_u32(sm + 128);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(sm + 136);
  unsigned char* ring = sm + kGmHeader;
  const int row0 = blockIdx.y * kGmRows, f0 = blockIdx.x * BN;   // column tiles of one row tile run back to back: z stays in L2
  const int KB = A.in_dim / 32;
  const int nsteps = A.heads * KB;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      gm_mbar_init(bar_full0 + 8 * s, kGmProducerWarps);
      gm_mbar_init(bar_empty0 + 8 * s, 1);
    }
    gm_mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kGmProducerWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gm_smem_u32(tmem_ptr_s)),
                 "r"(A.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp < kGmProducerWarps) {
    // =========================== producers: z / W slices -> swizzled operand blocks ===========================
    const int pt = tid;                                       // 0..127
    const int a_chunks = kGmRows * 8, b_chunks = BN * 8;      // 16-byte chunks per block
    for (int step = 0; step < nsteps; ++step) {
      const int h = step / KB, kb = step - h * KB;
      const int s = step % S, n = step / S;
      if (n > 0) gm_mbar_wait(bar_empty0 + 8 * s, (uint32_t)((n - 1) & 1));
      unsigned char* Ab = ring + (size_t)s * stage_bytes;
      unsigned char* Bb = Ab + a_bytes * (PASSES == 3 ? 2 : 1);
      // A: rows row0 .. row0+127 of z[:, h, kb*32 .. +32]
      for (int c = pt; c < a_chunks; c += kGmProducerWarps * 32 * 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cc = c + u * kGmProducerWarps * 32;
          const int r = cc >> 3, ch = cc & 7;
          const int node = row0 + r;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cc < a_chunks && node < A.N)
            v[u] = __ldg(reinterpret_cast<const float4*>(A.z + ((size_t)node * A.heads + h) * A.in_dim + kb * 32) + ch);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cc = c + u * kGmProducerWarps * 32;
          if (cc >= a_chunks) break;
          const int r = cc >> 3, ch = cc & 7;
          unsigned char* dst = Ab + r * 128 + ((ch ^ (r & 7)) << 4);
          if (PASSES == 3) {
            const float4 hi = make_float4(gm_tf32(v[u].x), gm_tf32(v[u].y), gm_tf32(v[u].z), gm_tf32(v[u].w));
            *reinterpret_cast<float4*>(dst) = hi;
            *reinterpret_cast<float4*>(dst + a_bytes) =
                make_float4(gm_tf32(v[u].x - hi.x), gm_tf32(v[u].y - hi.y), gm_tf32(v[u].z - hi.z), gm_tf32(v[u].w - hi.w));
          } else {
            *reinterpret_cast<float4*>(dst) = v[u];
          }
        }
      }
      // B: rows f0 .. f0+BN-1 of W[h, :, kb*32 .. +32]
      for (int c = pt; c < b_chunks; c += kGmProducerWarps * 32 * 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cc = c + u * kGmProducerWarps * 32;
          const int r = cc >> 3, ch = cc & 7;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cc < b_chunks) v[u] = __ldg(reinterpret_cast<const float4*>(A.W + ((size_t)h * A.F + f0 + r) * A.in_dim + kb * 32) + ch);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cc = c + u * kGmProducerWarps * 32;
          if (cc >= b_chunks) break;
          const int r = cc >> 3, ch = cc & 7;
          unsigned char* dst = Bb + r * 128 + ((ch ^ (r & 7)) << 4);
          if (PASSES == 3) {
            const float4 hi = make_float4(gm_tf32(v[u].x), gm_tf32(v[u].y), gm_tf32(v[u].z), gm_tf32(v[u].w));
            *reinterpret_cast<float4*>(dst) = hi;
            *reinterpret_cast<float4*>(dst + b_bytes) =
                make_float4(gm_tf32(v[u].x - hi.x), gm_tf32(v[u].y - hi.y), gm_tf32(v[u].z - hi.z), gm_tf32(v[u].w - hi.w));
          } else {
            *reinterpret_cast<float4*>(dst) = v[u];
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) gm_mbar_arrive(bar_full0 + 8 * s);
    }
  } else if (warp == kGmProducerWarps) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kGmRows >> 4) << 24);
      for (int step = 0; step < nsteps; ++step) {
        const int h = step / KB, kb = step - h * KB;
        const int s = step % S, n = step / S;
        gm_mbar_wait(bar_full0 + 8 * s, (uint32_t)(n & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = gm_smem_u32(ring + (size_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + a_bytes * (PASSES == 3 ? 2 : 1);
        const uint64_t a_hi = gm_desc_sw128(a_addr), b_hi = gm_desc_sw128(b_addr);
        const uint32_t d_tmem = tmem_base + (uint32_t)(h * BN);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          gm_mma_tf32(d_tmem, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          if (PASSES == 3) {
            const uint64_t a_lo = gm_desc_sw128(a_addr + a_bytes), b_lo = gm_desc_sw128(b_addr + b_bytes);
            gm_mma_tf32(d_tmem, a_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, 1u);
            gm_mma_tf32(d_tmem, a_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), idesc, 1u);
          }
        }
        gm_commit(bar_empty0 + 8 * s);
      }
      gm_commit(bar_acc);
    }
    __syncwarp();
  } else {
    // =========================== epilogue: TMEM -> ELU -> mean / concat -> global ===========================
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const int node = row0 + row;
    const float inv_h = 1.f / (float)A.heads;
    const int out_w = A.concat ? A.heads * A.F : A.F;
    gm_mbar_wait(bar_acc, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16);
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float oacc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) oacc[i] = 0.f;
      for (int h = 0; h < A.heads; ++h) {
        uint32_t v[16];
        gm_tmem_ld16(t_row + (uint32_t)(h * BN + c0), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float e16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float y = __uint_as_float(v[i]);
          // ELU (graph_attention.py:118): accurate expm1 on the fp32 path, hardware ex2 on the bf16 path
          e16[i] = PASSES == 3 ? elu1(y) : fmaxf(y, 0.f) + (gm_ex2(fminf(y, 0.f) * 1.4426950408889634f) - 1.f);
        }
        if (A.concat) {
          if (node < A.N) {
            const size_t o = (size_t)node * out_w + (size_t)h * A.F + f0 + c0;
            if (A.out_bf16) {
              uint4 pk[2];
              unsigned* pw = reinterpret_cast<unsigned*>(pk);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(e16[2 * i], e16[2 * i + 1]);
                pw[i] = *reinterpret_cast<unsigned*>(&b2);
              }
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(A.out) + o);
              op[0] = pk[0]; op[1] = pk[1];
            } else {
              float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(A.out) + o);
#pragma unroll
              for (int i = 0; i < 4; ++i) op[i] = make_float4(e16[4 * i], e16[4 * i + 1], e16[4 * i + 2], e16[4 * i + 3]);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) oacc[i] += e16[i];
        }
      }
      if (!A.concat && node < A.N) {
        const size_t o = (size_t)node * out_w + f0 + c0;
        if (A.out_bf16) {
          uint4 pk[2];
          unsigned* pw = reinterpret_cast<unsigned*>(pk);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            __nv_bfloat162 b2 = __floats2bfloat162_rn(oacc[2 * i] * inv_h, oacc[2 * i + 1] * inv_h);
            pw[i] = *reinterpret_cast<unsigned*>(&b2);
          }
          uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(A.out) + o);
          op[0] = pk[0]; op[1] = pk[1];
        } else {
          float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(A.out) + o);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            op[i] = make_float4(oacc[4 * i] * inv_h, oacc[4 * i + 1] * inv_h, oacc[4 * i + 2] * inv_h, oacc[4 * i + 3] * inv_h);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == kGmProducerWarps) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(A.tmem_cols) : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
struct GemmPlan { bool ok; int BN, stages, tmem_cols; size_t smem; };

static GemmPlan plan_gemm(int in_dim, int F, int heads, int passes) {
  GemmPlan p{false, 0, 0, 0, 0};
  if (in_dim % 32 != 0 || F % 16 != 0 || heads < 1 || heads > 8) return p;
  int BN = 0;
  for (int cand : {128, 64, 32, 16})
    if (F % cand == 0 && heads * cand <= 256) { BN = cand; break; }     // <= 256 TMEM columns: two CTAs per SM
  if (!BN) return p;
  int cols = 32;
  while (cols < heads * BN) cols <<= 1;
  const int stage = gm_stage_bytes(BN, passes);
  int stages = std::min(4, (100 * 1024 - kGmHeader) / stage);
  if (stages < 2) stages = std::min(4, (200 * 1024 - kGmHeader) / stage);
  if (stages < 2) return p;
  p.ok = true; p.BN = BN; p.stages = stages; p.tmem_cols = cols;
  p.smem = (size_t)kGmHeader + (size_t)stages * stage + 1024;
  return p;
}

bool gat_transform_tc_supported(int N, int in_dim, int F, int heads, int passes) {
  static const int enabled = getenv("MG_GAT_TC_GEMM") ? atoi(getenv("MG_GAT_TC_GEMM")) : 1;
  if (!enabled) return false;
  if (2.0 * N * in_dim * (double)F * heads < 1.0e9) return false;        // small problems: launch-latency bound anyway
  // 3xTF32: the tensor core's fp32 accumulation error grows ~linearly with K (measured max-abs vs the CPU oracle at
  // |y| ~ 2.5: 3.2e-6 / 5.1e-6 / 9.5e-6 for K = 128 / 256 / 512; 2.6e-5 at K = 512 without head averaging,
  // profiles/r1_tc_gemm_check.md).  Keep a 2x margin under the 1e-5 fp32 tolerance: K <= 256 (<= 128 for one head).
  if (passes == 3 && in_dim * (heads == 1 ? 2 : 1) > 256) return false;
  return plan_gemm(in_dim, F, heads, passes).ok;
}

int gat_transform_tc_launch(const float* z, const float* W, int N, int in_dim, int F, int heads, int concat, void* out,
                            int out_bf16, int passes, cudaStream_t st) {
  const GemmPlan p = plan_gemm(in_dim, F, heads, passes);
  if (!p.ok) {
    set_error("gat_transform_tc: unsupported shape in=%d F=%d heads=%d", in_dim, F, heads);
    return MG_ERR_UNSUPPORTED;
  }
  GemmTcArgs A;
  A.z = z; A.W = W; A.out = out; A.N = N; A.in_dim = in_dim; A.F = F; A.heads = heads; A.concat = concat;
  A.out_bf16 = out_bf16; A.BN = p.BN; A.stages = p.stages; A.tmem_cols = p.tmem_cols;
  dim3 grid(F / p.BN, ceil_div(N, kGmRows));
  if (grid.y > 65535) {
    set_error("gat_transform_tc: N=%d too large", N);
    return MG_ERR_UNSUPPORTED;
  }
  if (cudaFuncSetAttribute(passes == 3 ? (const void*)gat_transform_tc_kernel<3> : (const void*)gat_transform_tc_kernel<1>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
    set_error("gat_transform_tc_kernel: cannot raise dynamic shared memory");
    return MG_ERR_CUDA;
  }
  if (passes == 3) gat_transform_tc_kernel<3><<<grid, kGmThreads, p.smem, st>>>(A);
  else gat_transform_tc_kernel<1><<<grid, kGmThreads, p.smem, st>>>(A);
  return check_launch("gat_transform_tc_kernel");
}

}  // namespace mg
