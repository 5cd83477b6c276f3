// __nv_bfloat16 instantiations of the GAT aggregation kernels (see gat_kernels.cuh).
#include "gat_kernels.cuh"

namespace mg {

int gat_launch_fused_bf16(const GatFusedArgs& A, int NH, DimCfg d, size_t smem, int grid, cudaStream_t st) {
  return dispatch_fused<__nv_bfloat16>(A, NH, d, smem, grid, st);
}

int gat_launch_agg_bf16(const GatAggArgs& a, int NH, DimCfg d, float* z, float* den, int grid, cudaStream_t st) {
  return dispatch_agg<__nv_bfloat16>(a, NH, d, z, den, grid, st);
}

}  // namespace mg
