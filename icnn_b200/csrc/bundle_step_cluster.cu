// Cluster-split instantiations of the K2 kernel (sample split over 2/4/8 CTAs by columns, DSMEM
// exchanges): the measured-and-rejected variant kept for ICNN_K2_CS / ICNN_K2_RESIDENT and its
// parity test.  Separate translation unit to keep the build parallel.
#include "bundle_step_kernel.cuh"

namespace icnn {

cudaError_t bundle_step_cluster_launch(const StepArgs& a, const K2Config& c, int B, cudaStream_t st) {
  if (c.cs == 8) return launch_k2<8, 8>(a, c, B, st);
  if (c.cs == 4) return launch_k2<8, 4>(a, c, B, st);
  return launch_k2<8, 2>(a, c, B, st);
}

}  // namespace icnn
