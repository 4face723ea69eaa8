// sweep_common.cuh — per-voxel body of the illumination sweep shared by the per-slice and the fused kernels.
// Follows AddDirLightShader.usf:69-128 / ChangeDirLightShader.usf:75-156 under the fp32
// arithmetic contract of tbrm_math.cuh.
#pragma once
#include "tbrm_math.cuh"
#include "../../include/tbrm.h"

namespace tbrm {

// Uniforms of one light in one axis pass (what SetUVOffset / SetUVWOffset / SetStepSize / the border sampler carry)
struct LightPass {
    float uv_off[2];
    float uvw_off[3];
    float step;         // StepSize * VOLUME_DENSITY
    float border;       // read-buffer sampler border colour
    float light_alpha;  // initial buffer value
};

struct SweepUniforms {
    int axis, dirn, start;  // slice j = start + dirn * k, k = 0..td[2]-1
    int td[3];              // transposed light dims: buffer is td[0] x td[1], td[2] slices
    int ldims[3];           // light volume dims
    int ddims[3];           // data volume dims
    float data_border;
    float clip_center[3], clip_dir[3];
    Windowing win;
    float sign;     // +1 add, -1 remove (Add only)
    LightPass a;    // the (added) light
    LightPass r;    // the removed light (Change only)
    int gate_saturate;  // Add: 1 (AddDirLightShader.usf:110), Change: 0
};

template <typename T>
struct Texel;
template <>
struct Texel<uint8_t> {
    static __device__ __forceinline__ float decode(uint8_t v) { return (float) v / 255.0f; }
};
template <>
struct Texel<uint16_t> {
    static __device__ __forceinline__ float decode(uint16_t v) { return (float) v / 65535.0f; }
};
template <>
struct Texel<float> {
    static __device__ __forceinline__ float decode(float v) { return v; }
};

// light volume / propagation buffer element access in the light pixel format (A.1)
__device__ __forceinline__ float light_load(const float* p, size_t i) { return p[i]; }
__device__ __forceinline__ float light_load(const uint8_t* p, size_t i) { return (float) p[i] / 255.0f; }
__device__ __forceinline__ void light_store(float* p, size_t i, float v) { p[i] = v; }
__device__ __forceinline__ void light_store(uint8_t* p, size_t i, float v) { p[i] = quant8(v); }

// pos = mul(int3(px,py,Loop), PermutationMatrix) (LightingShaderUtils.cpp:227-249)
__device__ __forceinline__ void permute(int axis, int px, int py, int loop, int& x, int& y, int& z) {
    if (axis == 0) {
        x = loop, y = px, z = py;
    } else if (axis == 1) {
        x = px, y = loop, z = py;
    } else {
        x = px, y = py, z = loop;
    }
}

// trilinear SampleLevel of the data volume with AM_Border (LightingShaders.h:88-89)
template <typename DataT>
__device__ __forceinline__ float sample_data_border(const DataT* __restrict__ data, int X, int Y, int Z, float u, float v, float w,
                                                     float border) {
    int i0, j0, k0;
    float fx, fy, fz;
    axis_taps(u, X, i0, fx);
    axis_taps(v, Y, j0, fy);
    axis_taps(w, Z, k0, fz);
    float t[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int x = i0 + (c & 1), y = j0 + ((c >> 1) & 1), z = k0 + (c >> 2);
        const bool in = (unsigned) x < (unsigned) X && (unsigned) y < (unsigned) Y && (unsigned) z < (unsigned) Z;
        t[c] = in ? Texel<DataT>::decode(__ldg(data + ((size_t) x + (size_t) X * ((size_t) y + (size_t) Y * (size_t) z)))) : border;
    }
    const float c00 = lerpf(t[0], t[1], fx), c01 = lerpf(t[2], t[3], fx);
    const float c10 = lerpf(t[4], t[5], fx), c11 = lerpf(t[6], t[7], fx);
    return lerpf(lerpf(c00, c01, fy), lerpf(c10, c11, fy), fz);
}

// AddDirLightShader.usf:84-113: opacity of the sample toward the light, weighted by the clip plane
template <typename DataT>
__device__ __forceinline__ float occlusion_sample(const SweepUniforms& U, const LightPass& L, const DataT* __restrict__ data,
                                                   const float4* __restrict__ tf, int x, int y, int z) {
    const float rx = (float) U.ldims[0], ry = (float) U.ldims[1], rz = (float) U.ldims[2];
    const float sx = ((float) x + 0.5f) / rx + L.uvw_off[0];
    const float sy = ((float) y + 0.5f) / ry + L.uvw_off[1];
    const float sz = ((float) z + 0.5f) / rz + L.uvw_off[2];
    const float dist = dot3(sx - U.clip_center[0], sy - U.clip_center[1], sz - U.clip_center[2], U.clip_dir[0], U.clip_dir[1],
                            U.clip_dir[2]);
    const float ox = sx - (sx + U.clip_dir[0] * dist), oy = sy - (sy + U.clip_dir[1] * dist), oz = sz - (sz + U.clip_dir[2] * dist);
    const float vx = ox * rx, vy = oy * ry, vz = oz * rz;
    const float vdist = sqrtf(dot3(vx, vy, vz, vx, vy, vz));
    const float sgn = dist > 0.0f ? 1.0f : (dist < 0.0f ? -1.0f : 0.0f);
    float w = 0.5f + (0.57735026919f * vdist * sgn);
    w = fminf(fmaxf(w, 0.0f), 1.0f);
    float cs = 0.0f;
    const bool inside = (sx == saturatef(sx)) && (sy == saturatef(sy)) && (sz == saturatef(sz));
    if (w > 0.0f && (!U.gate_saturate || inside)) {
        const float v = sample_data_border<DataT>(data, U.ddims[0], U.ddims[1], U.ddims[2], sx, sy, sz, U.data_border);
        float pos;
        if (tf_position(v, U.win, pos)) {
            int i0, i1;
            float f;
            tf_taps(pos, i0, i1, f);
            const float a = lerpf(__ldg(&tf[i0]).w, __ldg(&tf[i1]).w, f);
            cs = step_opacity(a, L.step) * w;
        } else {
            cs = 0.0f * w;
        }
    }
    return cs;
}

}  // namespace tbrm
