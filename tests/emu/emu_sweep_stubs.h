// emu_sweep_stubs.h — TEST INFRASTRUCTURE (tests/emu): stands in for sweep_fused.cuh / sweep_tma.cuh, whose kernels are inline PTX
// (relaxed / system-scope accesses, mbarriers, TMA) and cooperative launches, when the product sources are compiled for the CPU. Both
// fused schedules report "not handled", exactly what they report on a device without cooperative launch, so the per-slice schedule
// (sweep_slice_kernel, the reference's own schedule) takes every pass; sharding a volume over GPUs is refused.
#pragma once

namespace tbrm {

cudaError_t sweep_pass_fused(tbrm_resources&, const SweepUniforms&, bool, int*, bool* handled) {
    *handled = false;
    return cudaSuccess;
}
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
