// mandelbulb.cu — procedural Mandelbulb SDF march (config 5), sm_100a.
// Reference: Source/FractalMarcher/Shaders/Private/SDFMarcher.usf:24-58 (Mandelbulb_SDF, GetActualPosition),
//            :61-112 (PerformMandelbulbRaymarchReturnDistance); entry from PerformRaymarchCubeSetup.
// Pure FP32 + SFU work: no textures, the only memory traffic is the output image.
#include <cstdlib>

#include "tbrm_internal.hpp"

namespace tbrm {

struct MbCam {
    float eye[3], fwd[3], rt[3], ut[3];
    float inv_w2, inv_h2;
    float m[4][3];
    float depth;
    int width, height, frame_mod8, jitter;
};

struct MbUniforms {
    MbCam cam;
    tbrm_mandelbulb mb;
    int row_begin, row_end;
    int p8;  // Power == 8: the transcendental-free iteration (mandelbulb_sdf_p8)
};

// TBRM_MANDELBULB_TRIG=1 forces the reference's transcendental formulation for every power
static bool mandelbulb_use_p8(float power) {
    static const bool force_trig = [] { const char* e = getenv("TBRM_MANDELBULB_TRIG"); return e && e[0] == '1'; }();
    return power == 8.0f && !force_trig;
}

__device__ __forceinline__ void mb_normalize(float& x, float& y, float& z) {
    const float l = sqrtf(dot3(x, y, z, x, y, z));
    x = x / l, y = y / l, z = z / l;
}
__device__ __forceinline__ void mb_mul3x3(float vx, float vy, float vz, const float m[4][3], float& ox, float& oy, float& oz) {
    ox = ((vx * m[0][0]) + (vy * m[1][0])) + (vz * m[2][0]);
    oy = ((vx * m[0][1]) + (vy * m[1][1])) + (vz * m[2][1]);
    oz = ((vx * m[0][2]) + (vy * m[1][2])) + (vz * m[2][2]);
}

// Mandelbulb_SDF — SDFMarcher.usf:24-52, the reference's formulation (acos / atan2 / pow / sin / cos per iteration)
__device__ __forceinline__ float mandelbulb_sdf_trig(float px, float py, float pz, float bailout, float power, int iterations,
                                                     unsigned int& iters) {
    float zx = px, zy = py, zz = pz;
    float dr = 1.0f, r = 0.0f;
    for (int i = 0; i < iterations; i++) {
        r = sqrtf(dot3(zx, zy, zz, zx, zy, zz));
        if (r > bailout) break;
        ++iters;
        float theta = acosf(zz / r);
        float phi = atan2f(zy, zx);
        dr = powf(r, power - 1.0f) * power * dr + 1.0f;
        const float zr = powf(r, power);
        theta = theta * power;
        phi = phi * power;
        float st, ct, sp, cp;
        sincosf(theta, &st, &ct);
        sincosf(phi, &sp, &cp);
        zx = zr * (st * cp) + px;
        zy = zr * (sp * st) + py;
        zz = zr * ct + pz;
    }
    return 0.5f * logf(r) * r / dr;
}

// The same iteration for Power == 8 (the plugin's default, FractalVolume.h:84-94) without transcendentals: with cos(theta) = z/r,
// sin(theta) = rho/r >= 0, cos(phi) = x/rho, sin(phi) = y/rho (rho = |(x,y)|; phi = 0 where rho = 0, as atan2(0,0)), three angle
// doublings give the sines and cosines of 8*theta and 8*phi, and r^8, r^7 come from squaring. Only +, -, *, /, sqrt (correctly
// rounded, --fmad=false) and the final log: the CPU oracle's twin (tbo_set_mandelbulb_variant(1)) produces the same bits up to
// that log. Against the reference's formulation the two differ by rounding only, which the iteration amplifies next to the surface
// exactly as libm-vs-CUDA transcendentals do: 0.12 % of the pixels of the 1080p frame differ by more than 1e-4 (measured on the CPU).
__host__ __device__ __forceinline__ float mandelbulb_sdf_p8(float px, float py, float pz, float bailout, int iterations, unsigned int& iters) {
    float zx = px, zy = py, zz = pz;
    float dr = 1.0f, r = 0.0f;
    for (int i = 0; i < iterations; i++) {
        const float r2 = ((zx * zx) + (zy * zy)) + (zz * zz);
        r = sqrtf(r2);
        if (r > bailout) break;
        ++iters;
        const float r4 = r2 * r2, r8 = r4 * r4;
        const float r7 = (r4 * r2) * r;
        dr = r7 * 8.0f * dr + 1.0f;
        const float rho2 = (zx * zx) + (zy * zy);
        const float rho = sqrtf(rho2);
        const float ir = 1.0f / r, irho = 1.0f / rho;  // two reciprocals and four products instead of four quotients
        float ct = zz * ir, st = rho * ir;
        float cp = rho > 0.0f ? zx * irho : 1.0f, sp = rho > 0.0f ? zy * irho : 0.0f;
#pragma unroll
        for (int d = 0; d < 3; ++d) {  // 8 * angle
            const float c2 = (ct * ct) - (st * st), s2 = 2.0f * (ct * st);
            ct = c2, st = s2;
            const float c3 = (cp * cp) - (sp * sp), s3 = 2.0f * (cp * sp);
            cp = c3, sp = s3;
        }
        zx = r8 * (st * cp) + px;
        zy = r8 * (sp * st) + py;
        zz = r8 * ct + pz;
    }
    return 0.5f * logf(r) * r / dr;
}

// the same source compiled for the host (the function is __host__ __device__; the host side is built with -ffp-contract=off): lets a machine
// without a GPU check this iteration against the oracle's twin (tbrm_debug_mandelbulb_sdf_p8, tests/test_host_cpu.py)
float mandelbulb_sdf_p8_host(float px, float py, float pz, float bailout, int iterations, unsigned int* iters) {
    unsigned int it = 0;
    const float d = mandelbulb_sdf_p8(px, py, pz, bailout, iterations, it);
    if (iters) *iters = it;
    return d;
}

// p8: Power == 8 and the doubling variant is enabled (uniform over the launch)
__device__ __forceinline__ float mandelbulb_sdf(float px, float py, float pz, float bailout, float power, int iterations, unsigned int& iters,
                                                bool p8) {
    return p8 ? mandelbulb_sdf_p8(px, py, pz, bailout, iterations, iters) : mandelbulb_sdf_trig(px, py, pz, bailout, power, iterations, iters);
}

// CameraVector + PerformRaymarchCubeSetup for pixel (ix, iy) (same arithmetic as raymarch.cu): entry position (cx,cy,cz) in UVW, local
// camera vector (lx,ly,lz), thickness
__device__ __forceinline__ void mb_pixel_setup(const MbCam& c, int ix, int iy, float& cx, float& cy, float& cz, float& lx, float& ly, float& lz,
                                               float& thick) {
    const float sx = ((float) ix + 0.5f) * c.inv_w2 - 1.0f;
    const float sy = 1.0f - ((float) iy + 0.5f) * c.inv_h2;
    float dx = (c.fwd[0] + c.rt[0] * sx) + c.ut[0] * sy, dy = (c.fwd[1] + c.rt[1] * sx) + c.ut[1] * sy,
          dz = (c.fwd[2] + c.rt[2] * sx) + c.ut[2] * sy;
    mb_normalize(dx, dy, dz);
    const float Vx = -dx, Vy = -dy, Vz = -dz;
    float nx = Vx, ny = Vy, nz = Vz;
    mb_normalize(nx, ny, nz);
    float wx, wy, wz;
    mb_mul3x3(nx * c.depth, ny * c.depth, nz * c.depth, c.m, wx, wy, wz);
    float depth = sqrtf(dot3(wx, wy, wz, wx, wy, wz));
    depth = depth / fabsf(dot3(c.fwd[0], c.fwd[1], c.fwd[2], Vx, Vy, Vz));
    float ox, oy, oz;
    mb_mul3x3(c.eye[0], c.eye[1], c.eye[2], c.m, ox, oy, oz);
    ox = ox + c.m[3][0], oy = oy + c.m[3][1], oz = oz + c.m[3][2];
    mb_mul3x3(Vx, Vy, Vz, c.m, lx, ly, lz);
    mb_normalize(lx, ly, lz);
    lx = -lx, ly = -ly, lz = -lz;
    ox = ox + 0.5f, oy = oy + 0.5f, oz = oz + 0.5f;
    const float ivx = 1.0f / lx, ivy = 1.0f / ly, ivz = 1.0f / lz;
    const float tminx = (0.0f - ox) * ivx, tminy = (0.0f - oy) * ivy, tminz = (0.0f - oz) * ivz;
    const float tmaxx = (1.0f - ox) * ivx, tmaxy = (1.0f - oy) * ivy, tmaxz = (1.0f - oz) * ivz;
    float t0 = fmaxf(fminf(tmaxx, tminx), fmaxf(fminf(tmaxy, tminy), fminf(tmaxz, tminz)));
    float t1 = fminf(fmaxf(tmaxx, tminx), fminf(fmaxf(tmaxy, tminy), fmaxf(tmaxz, tminz)));
    t0 = fmaxf(0.0f, t0);
    t1 = fminf(depth, t1);
    thick = fmaxf(0.0f, t1 - t0);
    cx = ox + (t0 * lx), cy = oy + (t0 * ly), cz = oz + (t0 * lz);
}

__global__ void __launch_bounds__(256) mandelbulb_kernel(const MbUniforms U, float2* __restrict__ out,
                                                         unsigned long long* __restrict__ iters_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ix = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int iy = U.row_begin + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    unsigned int iters = 0;
    if (ix < U.cam.width && iy < U.row_end) {
        const MbCam& c = U.cam;
        float cx, cy, cz, lx, ly, lz, thick;
        mb_pixel_setup(c, ix, iy, cx, cy, cz, lx, ly, lz, thick);

        float rx = 0.0f, ry = 0.0f;
        if (thick > 0.0f) {
            const tbrm_mandelbulb& mb = U.mb;
            const float stx = lx / mb.extent, sty = ly / mb.extent, stz = lz / mb.extent;  // SDFMarcher.usf:76
            const int max_iter = (int) mb.max_iterations;
            float dist = 0.0f;
            bool done = false;
            for (int s = 0; (float) s < mb.max_steps; s++) {
                const float apx = mb.center[0] + ((cx - 0.5f) * mb.extent), apy = mb.center[1] + ((cy - 0.5f) * mb.extent),
                            apz = mb.center[2] + ((cz - 0.5f) * mb.extent);
                dist = mandelbulb_sdf(apx, apy, apz, mb.bailout, mb.power, max_iter, iters, U.p8 != 0);
                if (dist < mb.high_precision_eps) {
                    float ratio = (float) s / (float) mb.max_steps;
                    ratio = ratio * 10.0f;
                    rx = 1.0f - ratio, ry = 1.0f;
                    done = true;
                    break;
                }
                cx = cx + (dist * stx), cy = cy + (dist * sty), cz = cz + (dist * stz);
                if (saturatef(cx) != cx || saturatef(cy) != cy || saturatef(cz) != cz) {
                    rx = 0.0f, ry = 0.0f;
                    done = true;
                    break;
                }
            }
            if (!done && dist < mb.low_precision_eps) rx = 0.0f, ry = 1.0f;
        }
        out[(size_t) (iy - U.row_begin) * c.width + ix] = make_float2(rx, ry);
    }
    if (iters_out) {
        unsigned int s = iters;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && s) atomicAdd(iters_out, (unsigned long long) s);
    }
}

// PerformMandelbulbRaymarchReturnNormal — SDFMarcher.usf:117-188: the same sphere tracing; a hit returns the normalised vector of three SDF
// evaluations at positions offset BACKWARDS by DerivationDistance / Extent along each axis (:156-165), alpha 1
__global__ void __launch_bounds__(256) mandelbulb_normal_kernel(const MbUniforms U, const float derivation_distance, float4* __restrict__ out,
                                                                unsigned long long* __restrict__ iters_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ix = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int iy = U.row_begin + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    unsigned int iters = 0;
    if (ix < U.cam.width && iy < U.row_end) {
        const MbCam& c = U.cam;
        float cx, cy, cz, lx, ly, lz, thick;
        mb_pixel_setup(c, ix, iy, cx, cy, cz, lx, ly, lz, thick);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (thick > 0.0f) {
            const tbrm_mandelbulb& mb = U.mb;
            const float stx = lx / mb.extent, sty = ly / mb.extent, stz = lz / mb.extent;  // :134
            const float dd = derivation_distance / mb.extent;                               // :138
            const int max_iter = (int) mb.max_iterations;
            float dist = 0.0f;
            bool done = false;
            for (int s = 0; (float) s < mb.max_steps; s++) {  // :142
                dist = mandelbulb_sdf(mb.center[0] + ((cx - 0.5f) * mb.extent), mb.center[1] + ((cy - 0.5f) * mb.extent),
                                      mb.center[2] + ((cz - 0.5f) * mb.extent), mb.bailout, mb.power, max_iter, iters, U.p8 != 0);
                if (dist < mb.high_precision_eps) {  // :147
                    float nx = mandelbulb_sdf(mb.center[0] + (((cx - dd) - 0.5f) * mb.extent), mb.center[1] + (((cy - 0.0f) - 0.5f) * mb.extent),
                                              mb.center[2] + (((cz - 0.0f) - 0.5f) * mb.extent), mb.bailout, mb.power, max_iter, iters, U.p8 != 0);
                    float ny = mandelbulb_sdf(mb.center[0] + (((cx - 0.0f) - 0.5f) * mb.extent), mb.center[1] + (((cy - dd) - 0.5f) * mb.extent),
                                              mb.center[2] + (((cz - 0.0f) - 0.5f) * mb.extent), mb.bailout, mb.power, max_iter, iters, U.p8 != 0);
                    float nz = mandelbulb_sdf(mb.center[0] + (((cx - 0.0f) - 0.5f) * mb.extent), mb.center[1] + (((cy - 0.0f) - 0.5f) * mb.extent),
                                              mb.center[2] + (((cz - dd) - 0.5f) * mb.extent), mb.bailout, mb.power, max_iter, iters, U.p8 != 0);
                    mb_normalize(nx, ny, nz);  // :166
                    o = make_float4(nx, ny, nz, 1.0f);
                    done = true;
                    break;
                }
                cx = cx + (dist * stx), cy = cy + (dist * sty), cz = cz + (dist * stz);  // :171
                if (saturatef(cx) != cx || saturatef(cy) != cy || saturatef(cz) != cz) {  // :174-177
                    done = true;
                    break;
                }
            }
            if (!done && dist < mb.low_precision_eps) o.w = 1.0f;  // :182-186 "return black normal"
        }
        out[(size_t) (iy - U.row_begin) * c.width + ix] = o;
    }
    if (iters_out) {
        unsigned int s = iters;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && s) atomicAdd(iters_out, (unsigned long long) s);
    }
}

// CalculateMandelbulbSDF.usf:24-65: per voxel the distance estimate (50 iterations, Bailout = Extent) divided by Extent, stored as UNORM16
// (the reference's PF_G16 volume, FractalVolume.cpp:166) or as the raw float. One thread per voxel, x fastest: pure FP32 + SFU work,
// the only memory traffic is the 2 (or 4) bytes written per voxel.
struct SdfUniforms {
    int dims[3];
    float center[3];
    float extent, power;
    int g16;
    int p8;
};
__global__ void __launch_bounds__(256) mandelbulb_sdf_kernel(const SdfUniforms U, void* __restrict__ out, unsigned long long* __restrict__ iters_out) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6), z = blockIdx.z;
    unsigned int iters = 0;
    if (x < U.dims[0] && y < U.dims[1]) {
        const float u = (float) x / (float) U.dims[0], v = (float) y / (float) U.dims[1], w = (float) z / (float) U.dims[2];  // :58 (no +0.5)
        const float px = U.center[0] + ((u - 0.5f) * U.extent), py = U.center[1] + ((v - 0.5f) * U.extent), pz = U.center[2] + ((w - 0.5f) * U.extent);
        const float d = mandelbulb_sdf(px, py, pz, U.extent, U.power, 50, iters, U.p8 != 0) / U.extent;  // :26-27, :63
        const size_t i = (size_t) x + (size_t) U.dims[0] * ((size_t) y + (size_t) U.dims[1] * z);
        if (U.g16)
            ((uint16_t*) out)[i] = (uint16_t) floorf(saturatef(d) * 65535.0f + 0.5f);
        else
            ((float*) out)[i] = d;
    }
    if (iters_out) {
        unsigned int s = iters;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(iters_out, (unsigned long long) s);
    }
}

cudaError_t mandelbulb_march_normal(cudaStream_t stream, const tbrm_mandelbulb& mb, float derivation_distance, const host::CameraUniforms& cam,
                                    int row_begin, int row_end, float* d_out, unsigned long long* d_iters) {
    MbUniforms U;
    memcpy(&U.cam, &cam, sizeof(U.cam));
    U.mb = mb;
    U.row_begin = row_begin, U.row_end = row_end;
    U.p8 = mandelbulb_use_p8(mb.power) ? 1 : 0;
    const dim3 grid((cam.width + 31) / 32, (row_end - row_begin + 7) / 8);
    mandelbulb_normal_kernel<<<grid, 256, 0, stream>>>(U, derivation_distance, (float4*) d_out, d_iters);
    count_launch();
    return cudaGetLastError();
}

cudaError_t mandelbulb_sdf_bake(cudaStream_t stream, const int32_t dims[3], const float center[3], float extent, float power, int g16, void* d_out,
                                unsigned long long* d_iters) {
    SdfUniforms U;
    for (int k = 0; k < 3; ++k) U.dims[k] = dims[k], U.center[k] = center[k];
    U.extent = extent, U.power = power, U.g16 = g16;
    U.p8 = mandelbulb_use_p8(power) ? 1 : 0;
    const dim3 grid((dims[0] + 63) / 64, (dims[1] + 3) / 4, dims[2]);
    mandelbulb_sdf_kernel<<<grid, 256, 0, stream>>>(U, d_out, d_iters);
    count_launch();
    return cudaGetLastError();
}

cudaError_t mandelbulb_march(cudaStream_t stream, const tbrm_mandelbulb& mb, const host::CameraUniforms& cam, int row_begin,
                             int row_end, float* d_out, unsigned long long* d_iters) {
    MbUniforms U;
    static_assert(sizeof(MbCam) == sizeof(host::CameraUniforms), "camera uniform layouts must match");
    memcpy(&U.cam, &cam, sizeof(U.cam));
    U.mb = mb;
    U.row_begin = row_begin, U.row_end = row_end;
    U.p8 = mandelbulb_use_p8(mb.power) ? 1 : 0;
    const dim3 grid((cam.width + 31) / 32, (row_end - row_begin + 7) / 8);
    mandelbulb_kernel<<<grid, 256, 0, stream>>>(U, (float2*) d_out, d_iters);
    count_launch();
    return cudaGetLastError();
}

}  // namespace tbrm
