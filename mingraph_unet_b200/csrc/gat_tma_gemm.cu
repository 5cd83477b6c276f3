// Node transform of the GAT layer for bf16 storage at large in*F: a persistent, TMA-fed, warp-specialised tcgen05 GEMM.
//   out[n, f] = mean_h / concat_h  ELU( sum_i z[n, h, i] * W[h, f, i] )     (model/gat/graph_attention.py:53,118,155-158)
//
// z (N, heads, in) is spilled by gat_aggregate_kernel as bf16 and W is converted to bf16 once per call; both are read
// by the TMA engine (cp.async.bulk.tensor.2d, 128-byte swizzle) straight into K-major UMMA operand tiles:
//   warp 0   one thread: TMA producer, 4-stage ring of (A 128 x 64, B BN x 64) bf16 tiles, mbarrier expect_tx
//   warp 1   one thread: tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), M128 x N=BN x K16, accumulators in TMEM,
//            double-buffered per HEAD so the epilogue of head h overlaps the MMAs of head h+1
//   warps 2-5 epilogue: tcgen05.ld -> ELU -> running head sum in registers -> (last head) bf16 tile through swizzled
//            shared memory -> coalesced 16-byte stores
// CTAs are persistent (one per SM) and walk the (row tile, column tile) grid column-fastest, so the CTAs working at
// the same time share the same rows of z in L2.
// Precision: z and W rounded to bf16 (2^-9 relative), fp32 accumulation: measured max-abs vs the fp32 oracle ~2e-3 on
// unit-scale inputs, inside the 2e-2 bf16 budget.  The fp32-storage path never takes this kernel.
#include <cuda.h>
#include <stdlib.h>

#include "gat_kernels.cuh"

namespace mg {

constexpr int kTmRows = 128;
constexpr int kTmKB = 64;                 // bf16 elements per K block = 128 bytes
constexpr int kTmStages = 4;
constexpr int kTmThreads = 6 * 32;
constexpr int kTmHeader = 1024;

struct TmaGemmArgs {
  void* out;
  int N, in_dim, F, heads, concat, out_bf16;
};

__device__ __forceinline__ uint32_t tm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tm_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tm_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tm_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tm_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TM_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TM_WAIT_DONE;\n"
      "bra TM_WAIT_LOOP;\n"
      "TM_WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tm_tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tm_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tm_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ uint64_t tm_desc_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tm_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float tm_elu(float v) {
  float e;
  const float t = fminf(v, 0.f) * 1.4426950408889634f;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return fmaxf(v, 0.f) + (e - 1.f);
}

template <int BN>
__global__ void __launch_bounds__(kTmThreads, 1)
gat_transform_tma_kernel(const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_w, const TmaGemmArgs A) {
  constexpr int A_BYTES = kTmRows * 128, B_BYTES = BN * 128, STAGE = A_BYTES + B_BYTES;
  extern __shared__ unsigned char tm_smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>(((uintptr_t)tm_smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar_full0 = tm_smem_u32(sm), bar_empty0 = tm_smem_u32(sm + 64);
  const uint32_t bar_accf0 = tm_smem_u32(sm + 128), bar_acce0 = tm_smem_u32(sm + 160);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(sm + 192);
  unsigned char* ring = sm + kTmHeader;
  unsigned char* Ss = ring + kTmStages * STAGE;                 // staging for the bf16 mean output tile (128 x BN x 2 B)
  const int KB = A.in_dim / kTmKB;
  const int col_tiles = A.F / BN, row_tiles = ceil_div(A.N, kTmRows);
  const int ntiles = col_tiles * row_tiles;

  if (tid == 0) {
    for (int s = 0; s < kTmStages; ++s) {
      tm_mbar_init(bar_full0 + 8 * s, 1);
      tm_mbar_init(bar_empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      tm_mbar_init(bar_accf0 + 8 * s, 1);
      tm_mbar_init(bar_acce0 + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_z) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tm_smem_u32(tmem_ptr_s)), "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = (tile / col_tiles) * kTmRows, f0 = (tile % col_tiles) * BN;
        for (int h = 0; h < A.heads; ++h)
          for (int kb = 0; kb < KB; ++kb, ++it) {
            const int s = it % kTmStages, n = it / kTmStages;
            if (n > 0) tm_mbar_wait(bar_empty0 + 8 * s, (uint32_t)((n - 1) & 1));
            const uint32_t a_dst = tm_smem_u32(ring + (size_t)s * STAGE), b_dst = a_dst + A_BYTES;
            tm_mbar_expect_tx(bar_full0 + 8 * s, (uint32_t)STAGE);
            tm_tma_load_2d(a_dst, &map_z, h * A.in_dim + kb * kTmKB, row0, bar_full0 + 8 * s);
            tm_tma_load_2d(b_dst, &map_w, kb * kTmKB, h * A.F + f0, bar_full0 + 8 * s);
          }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTmRows >> 4) << 24);
      int it = 0, ac = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int h = 0; h < A.heads; ++h, ++ac) {
          const int as = ac & 1, an = ac >> 1;
          if (an > 0) tm_mbar_wait(bar_acce0 + 8 * as, (uint32_t)((an - 1) & 1));       // epilogue drained this stage
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
          for (int kb = 0; kb < KB; ++kb, ++it) {
            const int s = it % kTmStages, n = it / kTmStages;
            tm_mbar_wait(bar_full0 + 8 * s, (uint32_t)(n & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_addr = tm_smem_u32(ring + (size_t)s * STAGE);
            const uint64_t ad = tm_desc_sw128(a_addr), bd = tm_desc_sw128(a_addr + A_BYTES);
#pragma unroll
            for (int k = 0; k < kTmKB / 16; ++k)                 // UMMA_K = 16 bf16 = 32 bytes = 2 descriptor units
              tm_mma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            tm_commit(bar_empty0 + 8 * s);
          }
          tm_commit(bar_accf0 + 8 * as);
        }
    }
    __syncwarp();
  } else {
    // =========================== epilogue ===========================
    const int ew = warp & 3;                                    // TMEM lane quarter (warps 2,3,4,5 -> 2,3,0,1)
    const int row = ew * 32 + lane;
    const int et = (warp - 2) * 32 + lane;
    const float inv_h = 1.f / (float)A.heads;
    const int out_w = A.concat ? A.heads * A.F : A.F;
    const bool staged = A.out_bf16 && !A.concat;
    constexpr int CPR = BN / 8;                                 // 16-byte chunks per staged row
    int ac = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int row0 = (tile / col_tiles) * kTmRows, f0 = (tile % col_tiles) * BN;
      const int node = row0 + row;
      float oacc[BN];
#pragma unroll
      for (int i = 0; i < BN; ++i) oacc[i] = 0.f;
      for (int h = 0; h < A.heads; ++h, ++ac) {
        const int as = ac & 1, an = ac >> 1;
        tm_mbar_wait(bar_accf0 + 8 * as, (uint32_t)(an & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v0[16], v1[16];
          tm_tmem_ld16(t_row + (uint32_t)c0, v0);
          tm_tmem_ld16(t_row + (uint32_t)(c0 + 16), v1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float e[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) { e[i] = tm_elu(__uint_as_float(v0[i])); e[16 + i] = tm_elu(__uint_as_float(v1[i])); }
          if (A.concat) {
            if (node < A.N) {
              const size_t o = (size_t)node * out_w + (size_t)h * A.F + f0 + c0;
              if (A.out_bf16) {
                uint4 pk[4];
                unsigned* pw = reinterpret_cast<unsigned*>(pk);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  __nv_bfloat162 b2 = __floats2bfloat162_rn(e[2 * i], e[2 * i + 1]);
                  pw[i] = *reinterpret_cast<unsigned*>(&b2);
                }
                uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(A.out) + o);
#pragma unroll
                for (int i = 0; i < 4; ++i) op[i] = pk[i];
              } else {
                float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(A.out) + o);
#pragma unroll
                for (int i = 0; i < 8; ++i) op[i] = make_float4(e[4 * i], e[4 * i + 1], e[4 * i + 2], e[4 * i + 3]);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) oacc[c0 + i] += e[i];
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tm_mbar_arrive(bar_acce0 + 8 * as);
      }
      if (!A.concat) {
        if (staged) {
#pragma unroll
          for (int c = 0; c < CPR; ++c) {
            uint4 pk;
            unsigned* pw = reinterpret_cast<unsigned*>(&pk);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __nv_bfloat162 b2 = __floats2bfloat162_rn(oacc[c * 8 + 2 * i] * inv_h, oacc[c * 8 + 2 * i + 1] * inv_h);
              pw[i] = *reinterpret_cast<unsigned*>(&b2);
            }
            *reinterpret_cast<uint4*>(Ss + (size_t)row * BN * 2 + ((c ^ (row & 7)) << 4)) = pk;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int valid_rows = min(kTmRows, A.N - row0);
          for (int idx = et; idx < valid_rows * CPR; idx += 128) {
            const int r = idx / CPR, c = idx - r * CPR;
            const uint4 val = *reinterpret_cast<const uint4*>(Ss + (size_t)r * BN * 2 + ((c ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(A.out) + (size_t)(row0 + r) * out_w + f0 + c * 8) = val;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        } else if (node < A.N) {
          if (A.out_bf16) {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(A.out) + (size_t)node * out_w + f0);
#pragma unroll
            for (int c = 0; c < CPR; ++c) {
              uint4 pk;
              unsigned* pw = reinterpret_cast<unsigned*>(&pk);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(oacc[c * 8 + 2 * i] * inv_h, oacc[c * 8 + 2 * i + 1] * inv_h);
                pw[i] = *reinterpret_cast<unsigned*>(&b2);
              }
              op[c] = pk;
            }
          } else {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(A.out) + (size_t)node * out_w + f0);
#pragma unroll
            for (int i = 0; i < BN / 4; ++i)
              op[i] = make_float4(oacc[4 * i] * inv_h, oacc[4 * i + 1] * inv_h, oacc[4 * i + 2] * inv_h, oacc[4 * i + 3] * inv_h);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

// W (heads, F, in) fp32 -> bf16
__global__ void tm_convert_w_kernel(const float* __restrict__ W, int64_t n, __nv_bfloat16* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(__ldg(W + i));
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static bool make_map_bf16_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};                     // bytes between rows
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int tma_bn(int F) { return F % 128 == 0 ? 128 : (F % 64 == 0 ? 64 : 0); }

static size_t tma_smem(int BN) {
  return (size_t)kTmHeader + (size_t)kTmStages * (kTmRows * 128 + BN * 128) + (size_t)kTmRows * BN * 2 + 1024;
}

bool gat_transform_tma_supported(int N, int in_dim, int F, int heads) {
  static const int enabled = getenv("MG_GAT_TMA_GEMM") ? atoi(getenv("MG_GAT_TMA_GEMM")) : 1;
  if (!enabled || encode_tiled_fn() == nullptr) return false;
  if (2.0 * N * in_dim * (double)F * heads < 1.0e9) return false;
  if (in_dim % kTmKB != 0 || heads < 1 || heads > 8 || tma_bn(F) == 0) return false;
  if ((int64_t)heads * F > 65535 * 16 || N < kTmRows) return false;
  return true;
}

int64_t gat_transform_tma_wbytes(int in_dim, int F, int heads) { return (((int64_t)heads * F * in_dim * 2) + 255) & ~(int64_t)255; }

int gat_transform_tma_launch(const void* z_bf16, const float* W, void* w_bf16, int N, int in_dim, int F, int heads, int concat,
                             void* out, int out_bf16, cudaStream_t st) {
  const int BN = tma_bn(F);
  const int64_t nw = (int64_t)heads * F * in_dim;
  tm_convert_w_kernel<<<(int)std::min<int64_t>(ceil_div64(nw, 256), (int64_t)num_sms() * 8), 256, 0, st>>>(
      W, nw, reinterpret_cast<__nv_bfloat16*>(w_bf16));
  int rc = check_launch("tm_convert_w_kernel");
  if (rc) return rc;
  alignas(64) CUtensorMap map_z, map_w;
  if (!make_map_bf16_2d(&map_z, z_bf16, (uint64_t)heads * in_dim, (uint64_t)N, kTmKB, kTmRows) ||
      !make_map_bf16_2d(&map_w, w_bf16, (uint64_t)in_dim, (uint64_t)heads * F, kTmKB, (uint32_t)BN)) {
    set_error("gat_transform_tma: cuTensorMapEncodeTiled failed");
    return MG_ERR_CUDA;
  }
  TmaGemmArgs A;
  A.out = out; A.N = N; A.in_dim = in_dim; A.F = F; A.heads = heads; A.concat = concat; A.out_bf16 = out_bf16;
  const int ntiles = ceil_div(N, kTmRows) * (F / BN);
  const int grid = std::min(ntiles, num_sms());
  const size_t smem = tma_smem(BN);
  if (BN == 128) {
    if (cudaFuncSetAttribute(gat_transform_tma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("gat_transform_tma_kernel: cannot raise dynamic shared memory");
      return MG_ERR_CUDA;
    }
    gat_transform_tma_kernel<128><<<grid, kTmThreads, smem, st>>>(map_z, map_w, A);
  } else {
    if (cudaFuncSetAttribute(gat_transform_tma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("gat_transform_tma_kernel: cannot raise dynamic shared memory");
      return MG_ERR_CUDA;
    }
    gat_transform_tma_kernel<64><<<grid, kTmThreads, smem, st>>>(map_z, map_w, A);
  }
  return check_launch("gat_transform_tma_kernel");
}

}  // namespace mg
