// sweep_fused.cuh — fused persistent plane-sweep (placeholder until the kernel lands).
#pragma once
namespace tbrm {
cudaError_t sweep_pass_fused(tbrm_resources&, const SweepUniforms&, bool, int*, bool* handled) {
    *handled = false;
    return cudaSuccess;
}
}  // namespace tbrm
