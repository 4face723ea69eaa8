// sweep.cu — illumination sweep kernels (AddDirLight / ChangeDirLight), sm_100a.
//
//  * sweep_slice_kernel: the reference's schedule — one launch per slice, ping-pong propagation buffers in global
//    memory (LightingShaders.cpp:132-158, 289-318). Handles every format / axis / half-res combination; it is the
//    fallback for configurations the fused kernel does not cover and the cross-check for it.
//  * the fused persistent plane-sweep kernel lives in sweep_fused.cuh (included below).
#include "tbrm_internal.hpp"

namespace tbrm {

// bilinear SampleLevel with AM_Border on a propagation buffer (LightingShaderUtils.cpp:190-195)
template <typename LightT>
__device__ __forceinline__ float sample_buffer_border(const LightT* __restrict__ buf, int W, int H, float u, float v, float border) {
    int i0, j0;
    float fx, fy;
    axis_taps(u, W, i0, fx);
    axis_taps(v, H, j0, fy);
    const bool x0 = (unsigned) i0 < (unsigned) W, x1 = (unsigned) (i0 + 1) < (unsigned) W;
    const bool y0 = (unsigned) j0 < (unsigned) H, y1 = (unsigned) (j0 + 1) < (unsigned) H;
    const float t00 = (x0 && y0) ? light_load(buf, (size_t) i0 + (size_t) W * j0) : border;
    const float t10 = (x1 && y0) ? light_load(buf, (size_t) i0 + 1 + (size_t) W * j0) : border;
    const float t01 = (x0 && y1) ? light_load(buf, (size_t) i0 + (size_t) W * (j0 + 1)) : border;
    const float t11 = (x1 && y1) ? light_load(buf, (size_t) i0 + 1 + (size_t) W * (j0 + 1)) : border;
    return lerpf(lerpf(t00, t10, fx), lerpf(t01, t11, fx), fy);
}

template <typename DataT, typename LightT, bool CHANGE>
__global__ void __launch_bounds__(256) sweep_slice_kernel(const SweepUniforms U, const int loop, const DataT* __restrict__ data,
                                                          const float4* __restrict__ tf, LightT* __restrict__ light,
                                                          const LightT* __restrict__ rd_a, LightT* __restrict__ wr_a,
                                                          const LightT* __restrict__ rd_r, LightT* __restrict__ wr_r) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y * blockDim.y + threadIdx.y;
    const int tx = U.td[0], ty = U.td[1];
    if (px >= tx || py >= ty) return;  // out-of-range UAV writes are dropped by D3D
    int x, y, z;
    permute(U.axis, px, py, loop, x, y, z);
    const float ub = ((float) px + 0.5f) / (float) tx, vb = ((float) py + 0.5f) / (float) ty;
    const size_t bi = (size_t) px + (size_t) tx * py;
    const size_t li = (size_t) x + (size_t) U.ldims[0] * ((size_t) y + (size_t) U.ldims[1] * (size_t) z);

    const float aprev = sample_buffer_border(rd_a, tx, ty, ub + U.a.uv_off[0], vb + U.a.uv_off[1], U.a.border);
    const float acs = occlusion_sample<DataT>(U, U.a, data, tf, x, y, z);
    const float acur = aprev * (1.0f - acs);
    light_store(wr_a, bi, acur);
    if (!CHANGE) {
        if (fabsf(acur) > 1e-3f) light_store(light, li, light_load(light, li) + (acur * U.sign));
    } else {
        const float rprev = sample_buffer_border(rd_r, tx, ty, ub + U.r.uv_off[0], vb + U.r.uv_off[1], U.r.border);
        const float rcs = occlusion_sample<DataT>(U, U.r, data, tf, x, y, z);
        const float rcur = rprev * (1.0f - rcs);
        light_store(wr_r, bi, rcur);
        if (fabsf(acur - rcur) > 1e-3f) light_store(light, li, light_load(light, li) + acur - rcur);
    }
}

// ---- same-face passes of several lights joined into one sweep ------------------------------------------------
// The reference's Readme (:165-166, 186-187) names this as the optimisation of the paper it does not implement. In the per-slice
// schedule — one launch per slice, launch-latency bound — a group of K lights that propagate from the same cube face costs the launches
// of one: every thread evaluates the K members in turn (own propagation buffers, the arithmetic of AddDirLightShader.usf unchanged) and
// adds each to the light volume through the volume's pixel format, exactly as K consecutive dispatches would at that voxel. A group of
// one is bit-identical to sweep_slice_kernel; several members differ from consecutive AddDirLight calls only in summation order.
struct JoinedPasses {
    int n;
    LightPass p[kMaxJoined];
};
__device__ __forceinline__ float light_roundtrip(const float*, float v) { return v; }
__device__ __forceinline__ float light_roundtrip(const uint8_t*, float v) { return (float) quant8(v) / 255.0f; }

template <typename DataT, typename LightT>
__global__ void __launch_bounds__(256) sweep_slice_joined_kernel(const SweepUniforms U, const JoinedPasses J, const int loop, const int parity,
                                                                 const DataT* __restrict__ data, const float4* __restrict__ tf,
                                                                 LightT* __restrict__ light, LightT* __restrict__ bufs, const size_t plane) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y * blockDim.y + threadIdx.y;
    const int tx = U.td[0], ty = U.td[1];
    if (px >= tx || py >= ty) return;
    int x, y, z;
    permute(U.axis, px, py, loop, x, y, z);
    const float ub = ((float) px + 0.5f) / (float) tx, vb = ((float) py + 0.5f) / (float) ty;
    const size_t bi = (size_t) px + (size_t) tx * py;
    const size_t li = (size_t) x + (size_t) U.ldims[0] * ((size_t) y + (size_t) U.ldims[1] * (size_t) z);
    float lv = light_load(light, li);
    bool changed = false;
    for (int m = 0; m < J.n; ++m) {
        const LightPass& L = J.p[m];
        const LightT* rd = bufs + (size_t) (2 * m + parity) * plane;
        LightT* wr = bufs + (size_t) (2 * m + (parity ^ 1)) * plane;
        const float prev = sample_buffer_border(rd, tx, ty, ub + L.uv_off[0], vb + L.uv_off[1], L.border);
        const float cs = occlusion_sample<DataT>(U, L, data, tf, x, y, z);
        const float cur = prev * (1.0f - cs);
        light_store(wr, bi, cur);
        if (fabsf(cur) > 1e-3f) {
            lv = light_roundtrip(light, lv + (cur * U.sign));
            changed = true;
        }
    }
    if (changed) light_store(light, li, lv);
}

template <typename T>
__global__ void fill_kernel(T* p, size_t n, T v) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
__global__ void fill_quant8_kernel(uint8_t* p, size_t n, float v) {
    const uint8_t q = quant8(v);
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = q;
}
__global__ void fill_float4_kernel(float4* p, size_t n4, float v) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    const float4 q = make_float4(v, v, v, v);
    for (; i < n4; i += stride) p[i] = q;
}

// Clear2DTexture_RenderThread / ClearVolumeTexture_RenderThread (UtilityShaders.cpp:27-75)
cudaError_t sweep_fill_buffer(tbrm_resources& r, void* buf, size_t count, float value) {
    if (count == 0) return cudaSuccess;
    if (r.light_fmt == TBRM_FMT_G8) {
        const int blocks = (int) std::min<size_t>((count + 255) / 256, 148 * 8);
        fill_quant8_kernel<<<blocks, 256, 0, r.stream>>>((uint8_t*) buf, count, value);
    } else if (count % 4 == 0) {
        const size_t n4 = count / 4;
        const int blocks = (int) std::min<size_t>((n4 + 255) / 256, 148 * 16);
        fill_float4_kernel<<<blocks, 256, 0, r.stream>>>((float4*) buf, n4, value);
    } else {
        const int blocks = (int) std::min<size_t>((count + 255) / 256, 148 * 8);
        fill_kernel<float><<<blocks, 256, 0, r.stream>>>((float*) buf, count, value);
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t clear_light(tbrm_resources& r, float value) { return sweep_fill_buffer(r, r.light, r.light_voxels(), value); }

template <typename DataT, typename LightT>
static cudaError_t per_slice_typed(tbrm_resources& r, const SweepUniforms& u, bool change, int* launches) {
    const int tx = u.td[0], ty = u.td[1], n = u.td[2];
    const size_t count = (size_t) tx * ty;
    void** B = r.rw[u.axis];
    cudaError_t e;
    // LightingShaders.cpp:74-79 (Add: buffers 0,1 <- LightAlpha) / :213-222 (Change: 0,1 <- removed, 2,3 <- added)
    if (!change) {
        if ((e = sweep_fill_buffer(r, B[0], count, u.a.light_alpha)) != cudaSuccess) return e;
        if ((e = sweep_fill_buffer(r, B[1], count, u.a.light_alpha)) != cudaSuccess) return e;
        *launches += 2;
    } else {
        if ((e = sweep_fill_buffer(r, B[0], count, u.r.light_alpha)) != cudaSuccess) return e;
        if ((e = sweep_fill_buffer(r, B[1], count, u.r.light_alpha)) != cudaSuccess) return e;
        if ((e = sweep_fill_buffer(r, B[2], count, u.a.light_alpha)) != cudaSuccess) return e;
        if ((e = sweep_fill_buffer(r, B[3], count, u.a.light_alpha)) != cudaSuccess) return e;
        *launches += 4;
    }
    const dim3 block(32, 8);
    const dim3 grid((tx + block.x - 1) / block.x, (ty + block.y - 1) / block.y);
    for (int k = 0; k < n; ++k) {
        const int j = u.start + u.dirn * k;
        const bool even = (j % 2 == 0);
        if (!change) {
            const LightT* rd = (const LightT*) (even ? B[0] : B[1]);
            LightT* wr = (LightT*) (even ? B[1] : B[0]);
            sweep_slice_kernel<DataT, LightT, false><<<grid, block, 0, r.stream>>>(u, j, (const DataT*) r.data, r.tf, (LightT*) r.light,
                                                                                   rd, wr, nullptr, nullptr);
        } else {
            const LightT* rrd = (const LightT*) (even ? B[0] : B[1]);
            LightT* rwr = (LightT*) (even ? B[1] : B[0]);
            const LightT* ard = (const LightT*) (even ? B[2] : B[3]);
            LightT* awr = (LightT*) (even ? B[3] : B[2]);
            sweep_slice_kernel<DataT, LightT, true><<<grid, block, 0, r.stream>>>(u, j, (const DataT*) r.data, r.tf, (LightT*) r.light,
                                                                                  ard, awr, rrd, rwr);
        }
    }
    count_launch(n);
    *launches += n;
    return cudaGetLastError();
}

template <typename DataT, typename LightT>
static cudaError_t joined_typed(tbrm_resources& r, const SweepUniforms& u, const JoinedPasses& J, int* launches) {
    const int tx = u.td[0], ty = u.td[1], n = u.td[2];
    const size_t plane = (size_t) tx * ty;
    const size_t need = 2 * (size_t) kMaxJoined * plane * r.light_elem();
    cudaError_t e;
    if (r.joined_bytes < need) {
        if (r.joined_buf) {
            cudaStreamSynchronize(r.stream);
            cudaFree(r.joined_buf);
            r.joined_buf = nullptr, r.joined_bytes = 0;
        }
        if ((e = cudaMalloc(&r.joined_buf, need)) != cudaSuccess) return e;
        r.joined_bytes = need;
    }
    LightT* bufs = (LightT*) r.joined_buf;
    for (int m = 0; m < J.n; ++m) {  // both buffers of a member start as its LightAlpha (LightingShaders.cpp:74-79)
        if ((e = sweep_fill_buffer(r, bufs + (size_t) 2 * m * plane, 2 * plane, J.p[m].light_alpha)) != cudaSuccess) return e;
    }
    *launches += J.n;
    const dim3 block(32, 8);
    const dim3 grid((tx + block.x - 1) / block.x, (ty + block.y - 1) / block.y);
    for (int k = 0; k < n; ++k) {
        const int j = u.start + u.dirn * k;
        sweep_slice_joined_kernel<DataT, LightT><<<grid, block, 0, r.stream>>>(u, J, j, (j % 2 == 0) ? 0 : 1, (const DataT*) r.data, r.tf,
                                                                                (LightT*) r.light, bufs, plane);
    }
    count_launch(n);
    *launches += n;
    return cudaGetLastError();
}

cudaError_t sweep_pass_joined(tbrm_resources& r, const SweepUniforms& u, const LightPass* members, int n_members, int* launches) {
    JoinedPasses J;
    J.n = std::min(n_members, kMaxJoined);
    for (int m = 0; m < J.n; ++m) J.p[m] = members[m];
    const bool l8 = r.light_fmt == TBRM_FMT_G8;
    switch (r.data_fmt) {
        case TBRM_FMT_G8: return l8 ? joined_typed<uint8_t, uint8_t>(r, u, J, launches) : joined_typed<uint8_t, float>(r, u, J, launches);
        case TBRM_FMT_G16: return l8 ? joined_typed<uint16_t, uint8_t>(r, u, J, launches) : joined_typed<uint16_t, float>(r, u, J, launches);
        default: return l8 ? joined_typed<float, uint8_t>(r, u, J, launches) : joined_typed<float, float>(r, u, J, launches);
    }
}

cudaError_t sweep_pass_per_slice(tbrm_resources& r, const SweepUniforms& u, bool change, int* launches) {
    const bool l8 = r.light_fmt == TBRM_FMT_G8;
    switch (r.data_fmt) {
        case TBRM_FMT_G8:
            return l8 ? per_slice_typed<uint8_t, uint8_t>(r, u, change, launches) : per_slice_typed<uint8_t, float>(r, u, change, launches);
        case TBRM_FMT_G16:
            return l8 ? per_slice_typed<uint16_t, uint8_t>(r, u, change, launches)
                      : per_slice_typed<uint16_t, float>(r, u, change, launches);
        default:
            return l8 ? per_slice_typed<float, uint8_t>(r, u, change, launches) : per_slice_typed<float, float>(r, u, change, launches);
    }
}

}  // namespace tbrm

#include "sweep_fused.cuh"
#include "sweep_tma.cuh"
