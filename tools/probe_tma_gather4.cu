// Compile probe (no GPU needed): ptxas for sm_100a accepts the TMA row-gather form and emits UTMALDG.2D.GATHER4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -c tools/probe_tma_gather4.cu -o /tmp/g4.o && cuobjdump -sass /tmp/g4.o | grep GATHER4
// Candidate for the CSR neighbour-row gather of the aggregation kernels: four arbitrary rows of x per
// instruction land in shared memory with no address arithmetic or scoreboard stalls in the warps.
#include <cuda.h>
#include <stdint.h>
__global__ void k(const __grid_constant__ CUtensorMap map, int* idx, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(b));
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(d), "l"(&map), "r"(0), "r"(idx[0]), "r"(idx[1]), "r"(idx[2]), "r"(idx[3]), "r"(b) : "memory");
  }
  out[threadIdx.x] = smem[threadIdx.x];
}
