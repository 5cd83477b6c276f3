// float instantiations of the GAT aggregation kernels (see gat_kernels.cuh).
#include "gat_kernels.cuh"

namespace mg {

int gat_launch_fused_f32(const GatFusedArgs& A, int NH, DimCfg d, size_t smem, int grid, cudaStream_t st) {
  return dispatch_fused<float>(A, NH, d, smem, grid, st);
}

int gat_launch_agg_f32(const GatAggArgs& a, int NH, DimCfg d, float* z, float* den, int grid, cudaStream_t st) {
  return dispatch_agg<float>(a, NH, d, z, den, grid, st);
}

}  // namespace mg
