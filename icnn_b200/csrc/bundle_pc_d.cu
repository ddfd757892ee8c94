// K2 predictor-corrector path: instantiations for n_y % 4 != 0 (scalar row loads; see bundle_pc.cu).
#include "bundle_pc_kernel.cuh"
namespace icnn {
cudaError_t launch_pc_101_1(const PcArgs& a, const PcConfig& c, int B, cudaStream_t st) { return launch_pc<1, 1, false>(a, c, B, st); }
cudaError_t launch_pc_101_2(const PcArgs& a, const PcConfig& c, int B, cudaStream_t st) { return launch_pc<1, 2, false>(a, c, B, st); }
cudaError_t launch_pc_102_1(const PcArgs& a, const PcConfig& c, int B, cudaStream_t st) { return launch_pc<2, 1, false>(a, c, B, st); }
cudaError_t launch_pc_102_2(const PcArgs& a, const PcConfig& c, int B, cudaStream_t st) { return launch_pc<2, 2, false>(a, c, B, st); }
}  // namespace icnn
