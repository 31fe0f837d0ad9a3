// tcgen05 engine (placeholder until the tensor-core conditioner lands).
#include "pzb_common.cuh"
namespace pzb {
bool tc_plan_supported(const PlanDev&) { return false; }
size_t tc_pack_bytes(const NscDev&) { return 0; }
void tc_pack_stage(const NscDev&, const float* const*, void*) {}
cudaError_t launch_tc(const PlanDev*, const PlanDev&, const LaunchArgs&, int, void*, cudaStream_t) {
  return cudaErrorNotSupported;
}
}  // namespace pzb
