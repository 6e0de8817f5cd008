// tcgen05 implementation of the fused SDF path (placeholder until the kernel lands).
#include "sdf_mlp.cuh"
namespace recmv {
int tc_sdf_forward(const PointSource&, const void*, const PeWeights&, float*, float*, int64_t, int,
                   cudaStream_t) {
  return RECMV_E_UNSUPPORTED;
}
}  // namespace recmv
