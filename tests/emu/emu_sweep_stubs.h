// emu_sweep_stubs.h — TEST INFRASTRUCTURE (tests/emu): stands in for sweep_tma.cuh (TMA, mbarriers, tensor maps from the driver) when the
// product sources are compiled for the CPU. The TMA-staged sweep reports "not handled", exactly what it reports where the driver entry
// point for tensor maps is missing, so the generic fused sweep (cooperative; emulated) or the per-slice schedule takes every pass;
// sharding a volume over GPUs is refused.
#pragma once

namespace tbrm {

cudaError_t sweep_pass_tma(tbrm_resources&, const SweepUniforms&, bool, int*, bool* handled) {
    *handled = false;
    return cudaSuccess;
}
int slab_pass_order(const SweepUniforms&) { return 2; }
size_t slab_arena_bytes(const int32_t ldims[3]) { return 64 + (size_t) ldims[0] * ldims[1] * 8; }
cudaError_t slab_ensure_arena(tbrm_resources&) {
    set_last_error("slab exchange is not emulated on the CPU");
    return cudaErrorNotSupported;
}

}  // namespace tbrm
