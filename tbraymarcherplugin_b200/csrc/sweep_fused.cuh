// sweep_fused.cuh — fused persistent plane-sweep: ONE launch per axis pass instead of one per slice.
//
// The reference runs TD.Z serial dispatches per pass and ping-pongs the propagated light through two VRAM
// textures (LightingShaders.cpp:132-158). Its single-dispatch "GPUSync" variant is disabled because a group-level
// barrier cannot order work across thread groups (AddDirLightShader_GPUSync.usf, Readme.md:178). Here the whole pass
// is one cooperative (all-CTAs-resident) kernel:
//   * the buffer plane (TD.X x TD.Y) is cut into tiles, one CTA per tile, each CTA walks all TD.Z slices;
//   * the propagated light of the last DEPTH slices lives in a small ring (DEPTH x TD.X x TD.Y fp32) that stays
//     resident in the 126 MB L2 — it never needs to reach HBM;
//   * a tile may start slice k as soon as the tiles its bilinear footprint touches have published slice k-1
//     (per-tile progress flags, release/acquire at gpu scope). The dependency is one-directional (toward the light),
//     so tiles form a skewed pipeline and flag latency only adds pipeline fill, not per-slice cost;
//   * back-pressure: a ring slot is reused only after every tile that reads it has moved on.
// Per-voxel arithmetic is sweep_common.cuh, i.e. bit-identical to the per-slice kernel and the oracle.
#pragma once
#include <cooperative_groups.h>

namespace tbrm {

constexpr int kFusedThreads = 256;
constexpr int kFusedMaxDeps = 32;
constexpr int kFlagStride = 8;  // one 32-byte sector per flag

struct FusedParams {
    SweepUniforms U;
    int tile_w, tile_h, ntx, nty;
    int depth;      // ring slots
    float* ring;    // [depth][CHANGE ? 2 : 1][td1][td0]
    unsigned int* flags;
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// bilinear AM_Border sample of ring slice `buf` (L2-coherent loads: other SMs wrote it)
__device__ __forceinline__ float sample_ring_border(const float* __restrict__ buf, int W, int H, float u, float v, float border) {
    int i0, j0;
    float fx, fy;
    axis_taps(u, W, i0, fx);
    axis_taps(v, H, j0, fy);
    const bool x0 = (unsigned) i0 < (unsigned) W, x1 = (unsigned) (i0 + 1) < (unsigned) W;
    const bool y0 = (unsigned) j0 < (unsigned) H, y1 = (unsigned) (j0 + 1) < (unsigned) H;
    const float t00 = (x0 && y0) ? __ldcg(buf + (size_t) i0 + (size_t) W * j0) : border;
    const float t10 = (x1 && y0) ? __ldcg(buf + (size_t) i0 + 1 + (size_t) W * j0) : border;
    const float t01 = (x0 && y1) ? __ldcg(buf + (size_t) i0 + (size_t) W * (j0 + 1)) : border;
    const float t11 = (x1 && y1) ? __ldcg(buf + (size_t) i0 + 1 + (size_t) W * (j0 + 1)) : border;
    return lerpf(lerpf(t00, t10, fx), lerpf(t01, t11, fx), fy);
}

// value a propagation buffer in the light pixel format would hold after storing v
template <typename LightT>
__device__ __forceinline__ float buffer_roundtrip(float v);
template <>
__device__ __forceinline__ float buffer_roundtrip<float>(float v) {
    return v;
}
template <>
__device__ __forceinline__ float buffer_roundtrip<uint8_t>(float v) {
    return (float) quant8(v) / 255.0f;
}

// conservative range of tiles touched by the bilinear footprint of pixels [p0, p1) shifted by `off` pixels
__device__ __forceinline__ void tile_range(int p0, int p1, float off_lo, float off_hi, int tile, int ntiles, int& lo, int& hi) {
    const int a = (int) floorf((float) p0 + off_lo) - 2;
    const int b = (int) floorf((float) (p1 - 1) + off_hi) + 3;
    lo = max(0, a) / tile;
    hi = min(ntiles - 1, max(0, b) / tile);
    if (a >= ntiles * tile) lo = ntiles;  // entirely outside: empty range
    if (b < 0) hi = -1;
}

template <typename DataT, typename LightT, bool CHANGE>
__global__ void __launch_bounds__(kFusedThreads) sweep_fused_kernel(const FusedParams P, const DataT* __restrict__ data,
                                                                    const float4* __restrict__ tf, LightT* __restrict__ light) {
    const SweepUniforms& U = P.U;
    const int tx = U.td[0], ty = U.td[1], ns = U.td[2];
    const int tile = blockIdx.x;
    const int tix = tile % P.ntx, tiy = tile / P.ntx;
    const int x0 = tix * P.tile_w, y0 = tiy * P.tile_h;
    const int x1 = min(x0 + P.tile_w, tx), y1 = min(y0 + P.tile_h, ty);
    const int tw = x1 - x0, npix = tw * (y1 - y0);
    const size_t plane = (size_t) tx * ty;
    const size_t slot_stride = plane * (CHANGE ? 2 : 1);

    __shared__ int s_up[kFusedMaxDeps], s_down[kFusedMaxDeps];
    __shared__ int s_nup, s_ndown;
    if (threadIdx.x == 0) {
        // pixel offsets of the read-buffer footprint (both lights for Change)
        float ox_lo = U.a.uv_off[0] * (float) tx, ox_hi = ox_lo, oy_lo = U.a.uv_off[1] * (float) ty, oy_hi = oy_lo;
        if (CHANGE) {
            const float rx = U.r.uv_off[0] * (float) tx, ry = U.r.uv_off[1] * (float) ty;
            ox_lo = fminf(ox_lo, rx), ox_hi = fmaxf(ox_hi, rx), oy_lo = fminf(oy_lo, ry), oy_hi = fmaxf(oy_hi, ry);
        }
        int nup = 0, ndown = 0, ax, bx, ay, by;
        tile_range(x0, x1, ox_lo, ox_hi, P.tile_w, P.ntx, ax, bx);
        tile_range(y0, y1, oy_lo, oy_hi, P.tile_h, P.nty, ay, by);
        for (int j = ay; j <= by; ++j)
            for (int i = ax; i <= bx; ++i)
                if ((i != tix || j != tiy) && nup < kFusedMaxDeps) s_up[nup++] = j * P.ntx + i;
        tile_range(x0, x1, -ox_hi, -ox_lo, P.tile_w, P.ntx, ax, bx);
        tile_range(y0, y1, -oy_hi, -oy_lo, P.tile_h, P.nty, ay, by);
        for (int j = ay; j <= by; ++j)
            for (int i = ax; i <= bx; ++i)
                if ((i != tix || j != tiy) && ndown < kFusedMaxDeps) s_down[ndown++] = j * P.ntx + i;
        s_nup = nup, s_ndown = ndown;
    }
    // prologue: the buffers start out cleared to LightAlpha (LightingShaders.cpp:74-79 / :213-222) = "slice -1"
    {
        float* slot = P.ring + (size_t) (P.depth - 1) * slot_stride;
        const float ia = buffer_roundtrip<LightT>(U.a.light_alpha), ir = buffer_roundtrip<LightT>(U.r.light_alpha);
        for (int p = threadIdx.x; p < npix; p += kFusedThreads) {
            const size_t bi = (size_t) (x0 + p % tw) + (size_t) tx * (y0 + p / tw);
            __stcg(slot + bi, ia);
            if (CHANGE) __stcg(slot + plane + bi, ir);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        st_release(P.flags + (size_t) tile * kFlagStride, 1u);
    }
    const int nup = s_nup, ndown = s_ndown;

    for (int k = 0; k < ns; ++k) {
        const int loop = U.start + U.dirn * k;
        // wait: upstream tiles published slice k-1; downstream tiles are done with the slot we are about to overwrite
        if ((int) threadIdx.x < nup) {
            const unsigned int* f = P.flags + (size_t) s_up[threadIdx.x] * kFlagStride;
            while (ld_acquire(f) < (unsigned) (k + 1)) {
            }
        } else if ((int) threadIdx.x >= 32 && (int) threadIdx.x - 32 < ndown && k - P.depth + 3 >= 2) {
            const unsigned int* f = P.flags + (size_t) s_down[threadIdx.x - 32] * kFlagStride;
            while (ld_acquire(f) < (unsigned) (k - P.depth + 3)) {
            }
        }
        __syncthreads();
        const float* rd = P.ring + (size_t) ((k + P.depth - 1) % P.depth) * slot_stride;
        float* wr = P.ring + (size_t) (k % P.depth) * slot_stride;
        for (int p = threadIdx.x; p < npix; p += kFusedThreads) {
            const int px = x0 + p % tw, py = y0 + p / tw;
            int x, y, z;
            permute(U.axis, px, py, loop, x, y, z);
            const float ub = ((float) px + 0.5f) / (float) tx, vb = ((float) py + 0.5f) / (float) ty;
            const size_t bi = (size_t) px + (size_t) tx * py;
            const size_t li = (size_t) x + (size_t) U.ldims[0] * ((size_t) y + (size_t) U.ldims[1] * (size_t) z);
            const float aprev = sample_ring_border(rd, tx, ty, ub + U.a.uv_off[0], vb + U.a.uv_off[1], U.a.border);
            const float acs = occlusion_sample<DataT>(U, U.a, data, tf, x, y, z);
            const float acur = aprev * (1.0f - acs);
            __stcg(wr + bi, buffer_roundtrip<LightT>(acur));
            if (!CHANGE) {
                if (fabsf(acur) > 1e-3f) light_store(light, li, light_load(light, li) + (acur * U.sign));
            } else {
                const float rprev = sample_ring_border(rd + plane, tx, ty, ub + U.r.uv_off[0], vb + U.r.uv_off[1], U.r.border);
                const float rcs = occlusion_sample<DataT>(U, U.r, data, tf, x, y, z);
                const float rcur = rprev * (1.0f - rcs);
                __stcg(wr + plane + bi, buffer_roundtrip<LightT>(rcur));
                if (fabsf(acur - rcur) > 1e-3f) light_store(light, li, light_load(light, li) + acur - rcur);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            st_release(P.flags + (size_t) tile * kFlagStride, (unsigned) (k + 2));
        }
    }
}

template <typename DataT, typename LightT, bool CHANGE>
static cudaError_t fused_typed(tbrm_resources& r, const SweepUniforms& u, int* launches, bool* handled) {
    auto kernel = sweep_fused_kernel<DataT, LightT, CHANGE>;
    int dev = r.device, sms = 0, per_sm = 0, coop = 0;
    cudaError_t e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kFusedThreads, 0)) != cudaSuccess) return e;
    const int capacity = sms * per_sm;
    if (!coop || capacity <= 0) return cudaSuccess;  // not handled -> per-slice path
    const int tx = u.td[0], ty = u.td[1];
    // smallest tiles (most parallelism, 1 pixel per thread minimum) whose count fits the co-resident capacity
    int tw = 32, th = 8;
    auto count = [&](int w, int h) { return (long long) ((tx + w - 1) / w) * ((ty + h - 1) / h); };
    while (count(tw, th) > capacity) {
        if (th < tw)
            th *= 2;
        else
            tw *= 2;
        if (tw > 4096) return cudaSuccess;
    }
    FusedParams P;
    P.U = u;
    P.tile_w = tw, P.tile_h = th;
    P.ntx = (tx + tw - 1) / tw, P.nty = (ty + th - 1) / th;
    P.depth = 8;
    const int ntiles = P.ntx * P.nty;
    // the footprint of a tile must fit the dependency lists
    {
        const float ox = fmaxf(fabsf(u.a.uv_off[0]), CHANGE ? fabsf(u.r.uv_off[0]) : 0.0f) * (float) tx;
        const float oy = fmaxf(fabsf(u.a.uv_off[1]), CHANGE ? fabsf(u.r.uv_off[1]) : 0.0f) * (float) ty;
        if (!(ox < 1e6f) || !(oy < 1e6f)) return cudaSuccess;
        const long long nx = (long long) ((ox + 6.0f) / tw) + 2, ny = (long long) ((oy + 6.0f) / th) + 2;
        if (nx * ny - 1 > kFusedMaxDeps) return cudaSuccess;
    }
    const size_t ring_bytes = (size_t) P.depth * (CHANGE ? 2 : 1) * tx * ty * sizeof(float);
    if (r.ring_bytes < ring_bytes) {
        if (r.ring) {
            cudaStreamSynchronize(r.stream);
            cudaFree(r.ring);
            r.ring = nullptr, r.ring_bytes = 0;
        }
        if ((e = cudaMalloc(&r.ring, ring_bytes)) != cudaSuccess) return e;
        r.ring_bytes = ring_bytes;
    }
    if (r.flags_count < (size_t) ntiles * kFlagStride) {
        if (r.flags) {
            cudaStreamSynchronize(r.stream);
            cudaFree(r.flags);
            r.flags = nullptr, r.flags_count = 0;
        }
        if ((e = cudaMalloc((void**) &r.flags, (size_t) ntiles * kFlagStride * sizeof(unsigned int))) != cudaSuccess) return e;
        r.flags_count = (size_t) ntiles * kFlagStride;
    }
    P.ring = (float*) r.ring;
    P.flags = r.flags;
    r.ring_epoch = 0;  // plain floats go into the ring: the TMA sweep must clear it before trusting tags again
    if ((e = cudaMemsetAsync(r.flags, 0, (size_t) ntiles * kFlagStride * sizeof(unsigned int), r.stream)) != cudaSuccess) return e;
    const DataT* d = (const DataT*) r.data;
    const float4* tf = r.tf;
    LightT* l = (LightT*) r.light;
    void* args[] = {(void*) &P, (void*) &d, (void*) &tf, (void*) &l};
    if ((e = cudaLaunchCooperativeKernel((const void*) kernel, dim3(ntiles), dim3(kFusedThreads), args, 0, r.stream)) != cudaSuccess) return e;
    count_launch();
    *launches += 1;
    *handled = true;
    return cudaSuccess;
}

cudaError_t sweep_pass_fused(tbrm_resources& r, const SweepUniforms& u, bool change, int* launches, bool* handled) {
    *handled = false;
    const bool l8 = r.light_fmt == TBRM_FMT_G8;
#define TBRM_FUSED(D, L) (change ? fused_typed<D, L, true>(r, u, launches, handled) : fused_typed<D, L, false>(r, u, launches, handled))
    switch (r.data_fmt) {
        case TBRM_FMT_G8: return l8 ? TBRM_FUSED(uint8_t, uint8_t) : TBRM_FUSED(uint8_t, float);
        case TBRM_FMT_G16: return l8 ? TBRM_FUSED(uint16_t, uint8_t) : TBRM_FUSED(uint16_t, float);
        default: return l8 ? TBRM_FUSED(float, uint8_t) : TBRM_FUSED(float, float);
    }
#undef TBRM_FUSED
}

}  // namespace tbrm
