// K2 predictor-corrector path: instantiations that keep the per-sample n-vectors in the caller's scratch
// (icnn_bundle_bufs::vec_ws) instead of shared memory (see bundle_pc.cu / bundle_pc_kernel.cuh, GV).
#include "bundle_pc_kernel.cuh"
namespace icnn {
cudaError_t launch_pc_201_4(const PcArgs& a, const PcConfig& c, int B, cudaStream_t st) { return launch_pc<1, 4, true, true>(a, c, B, st); }
cudaError_t launch_pc_202_2(const PcArgs& a, const PcConfig& c, int B, cudaStream_t st) { return launch_pc<2, 2, true, true>(a, c, B, st); }
cudaError_t launch_pc_204_2(const PcArgs& a, const PcConfig& c, int B, cudaStream_t st) { return launch_pc<4, 2, true, true>(a, c, B, st); }
cudaError_t launch_pc_208_2(const PcArgs& a, const PcConfig& c, int B, cudaStream_t st) { return launch_pc<8, 2, true, true>(a, c, B, st); }
cudaError_t launch_pc_208_4(const PcArgs& a, const PcConfig& c, int B, cudaStream_t st) { return launch_pc<8, 4, true, true>(a, c, B, st); }
}  // namespace icnn
