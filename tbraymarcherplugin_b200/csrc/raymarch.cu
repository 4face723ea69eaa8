// raymarch.cu — per-pixel lit ray march (PerformRaymarchCubeSetup + PerformWindowedLitRaymarch), sm_100a.
// Reference: Source/Raymarcher/Shaders/Private/RaymarchMaterialCommon.usf:23-88,
//            Source/Raymarcher/Shaders/Private/WindowedRaymarchMaterials.usf:21-96.
#include "tbrm_internal.hpp"

namespace tbrm {

struct RayCam {
    float eye[3], fwd[3], rt[3], ut[3];
    float inv_w2, inv_h2;
    float m[4][3];
    float depth;
    int width, height, frame_mod8, jitter;
};

struct MarchUniforms {
    RayCam cam;
    int ddims[3], ldims[3];
    Windowing win;
    float clip_center[3], clip_dir[3];
    float step_count;
    int row_begin, row_end;
    int data_wrap;
};

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 normalize3(V3 a) {
    const float l = sqrtf(dot3(a.x, a.y, a.z, a.x, a.y, a.z));
    return v3(a.x / l, a.y / l, a.z / l);
}
__device__ __forceinline__ V3 mul3x3(V3 v, const float m[4][3]) {
    return v3(((v.x * m[0][0]) + (v.y * m[1][0])) + (v.z * m[2][0]), ((v.x * m[0][1]) + (v.y * m[1][1])) + (v.z * m[2][1]),
              ((v.x * m[0][2]) + (v.y * m[1][2])) + (v.z * m[2][2]));
}

// MaterialParameters.CameraVector of pixel (ix,iy)
__device__ __forceinline__ V3 camera_vector(const RayCam& c, int ix, int iy) {
    const float sx = ((float) ix + 0.5f) * c.inv_w2 - 1.0f;
    const float sy = 1.0f - ((float) iy + 0.5f) * c.inv_h2;
    V3 d = v3((c.fwd[0] + c.rt[0] * sx) + c.ut[0] * sy, (c.fwd[1] + c.rt[1] * sx) + c.ut[1] * sy,
              (c.fwd[2] + c.rt[2] * sx) + c.ut[2] * sy);
    d = normalize3(d);
    return v3(-d.x, -d.y, -d.z);
}

// PerformRaymarchCubeSetup — RaymarchMaterialCommon.usf:23-69
__device__ __forceinline__ void cube_setup(const RayCam& c, V3 V, V3& entry, float& thick, V3& lcv) {
    float depth = c.depth;
    const V3 n = normalize3(V);
    V3 wd = v3(n.x * depth, n.y * depth, n.z * depth);
    wd = mul3x3(wd, c.m);
    depth = sqrtf(dot3(wd.x, wd.y, wd.z, wd.x, wd.y, wd.z));
    depth = depth / fabsf(dot3(c.fwd[0], c.fwd[1], c.fwd[2], V.x, V.y, V.z));
    V3 o = mul3x3(v3(c.eye[0], c.eye[1], c.eye[2]), c.m);
    o = v3(o.x + c.m[3][0], o.y + c.m[3][1], o.z + c.m[3][2]);
    const V3 mv = normalize3(mul3x3(V, c.m));
    lcv = v3(-mv.x, -mv.y, -mv.z);
    o = v3(o.x + 0.5f, o.y + 0.5f, o.z + 0.5f);
    // RayAABBIntersection — RaymarcherCommon.usf:66-88
    const V3 inv = v3(1.0f / lcv.x, 1.0f / lcv.y, 1.0f / lcv.z);
    const V3 tmin = v3((0.0f - o.x) * inv.x, (0.0f - o.y) * inv.y, (0.0f - o.z) * inv.z);
    const V3 tmax = v3((1.0f - o.x) * inv.x, (1.0f - o.y) * inv.y, (1.0f - o.z) * inv.z);
    float t0 = fmaxf(fminf(tmax.x, tmin.x), fmaxf(fminf(tmax.y, tmin.y), fminf(tmax.z, tmin.z)));
    float t1 = fminf(fmaxf(tmax.x, tmin.x), fminf(fmaxf(tmax.y, tmin.y), fmaxf(tmax.z, tmin.z)));
    t0 = fmaxf(0.0f, t0);
    t1 = fminf(depth, t1);
    thick = fmaxf(0.0f, t1 - t0);
    entry = v3(o.x + (t0 * lcv.x), o.y + (t0 * lcv.y), o.z + (t0 * lcv.z));
}

// Rand3DPCG16(...).x (UE Random.ush)
__device__ __forceinline__ uint32_t pcg16_x(int px, int py, int pz) {
    uint32_t x = (uint32_t) px, y = (uint32_t) py, z = (uint32_t) pz;
    x = x * 1664525u + 1013904223u;
    y = y * 1664525u + 1013904223u;
    z = z * 1664525u + 1013904223u;
    x += y * z;
    y += z * x;
    z += x * y;
    x += y * z;
    y += z * x;
    z += x * y;
    return x >> 16;
}

__device__ __forceinline__ int wrap_index(int i, int n) {
    int r = i % n;
    return r < 0 ? r + n : r;
}
__device__ __forceinline__ int clamp_index(int i, int n) { return min(max(i, 0), n - 1); }

template <typename DataT>
__device__ __forceinline__ float sample_data(const DataT* __restrict__ data, const int dims[3], V3 p, bool wrap) {
    int i0, j0, k0;
    float fx, fy, fz;
    axis_taps(p.x, dims[0], i0, fx);
    axis_taps(p.y, dims[1], j0, fy);
    axis_taps(p.z, dims[2], k0, fz);
    int xs0, xs1, ys0, ys1, zs0, zs1;
    if (wrap) {
        xs0 = wrap_index(i0, dims[0]), xs1 = wrap_index(i0 + 1, dims[0]);
        ys0 = wrap_index(j0, dims[1]), ys1 = wrap_index(j0 + 1, dims[1]);
        zs0 = wrap_index(k0, dims[2]), zs1 = wrap_index(k0 + 1, dims[2]);
    } else {
        xs0 = clamp_index(i0, dims[0]), xs1 = clamp_index(i0 + 1, dims[0]);
        ys0 = clamp_index(j0, dims[1]), ys1 = clamp_index(j0 + 1, dims[1]);
        zs0 = clamp_index(k0, dims[2]), zs1 = clamp_index(k0 + 1, dims[2]);
    }
    const size_t X = dims[0], XY = (size_t) dims[0] * dims[1];
    const size_t r00 = X * ys0 + XY * zs0, r01 = X * ys1 + XY * zs0, r10 = X * ys0 + XY * zs1, r11 = X * ys1 + XY * zs1;
    const float c00 = lerpf(Texel<DataT>::decode(__ldg(data + r00 + xs0)), Texel<DataT>::decode(__ldg(data + r00 + xs1)), fx);
    const float c01 = lerpf(Texel<DataT>::decode(__ldg(data + r01 + xs0)), Texel<DataT>::decode(__ldg(data + r01 + xs1)), fx);
    const float c10 = lerpf(Texel<DataT>::decode(__ldg(data + r10 + xs0)), Texel<DataT>::decode(__ldg(data + r10 + xs1)), fx);
    const float c11 = lerpf(Texel<DataT>::decode(__ldg(data + r11 + xs0)), Texel<DataT>::decode(__ldg(data + r11 + xs1)), fx);
    return lerpf(lerpf(c00, c01, fy), lerpf(c10, c11, fy), fz);
}

// LightVolume.SampleLevel(Material.Wrap_WorldGroupSettings, saturate(CurPos), 0).r — WindowedRaymarchMaterials.usf:30
template <typename LightT>
__device__ __forceinline__ float sample_light_wrap(const LightT* __restrict__ light, const int dims[3], V3 p) {
    int i0, j0, k0;
    float fx, fy, fz;
    axis_taps(p.x, dims[0], i0, fx);
    axis_taps(p.y, dims[1], j0, fy);
    axis_taps(p.z, dims[2], k0, fz);
    const int xs0 = wrap_index(i0, dims[0]), xs1 = wrap_index(i0 + 1, dims[0]);
    const int ys0 = wrap_index(j0, dims[1]), ys1 = wrap_index(j0 + 1, dims[1]);
    const int zs0 = wrap_index(k0, dims[2]), zs1 = wrap_index(k0 + 1, dims[2]);
    const size_t X = dims[0], XY = (size_t) dims[0] * dims[1];
    const size_t r00 = X * ys0 + XY * zs0, r01 = X * ys1 + XY * zs0, r10 = X * ys0 + XY * zs1, r11 = X * ys1 + XY * zs1;
    const float c00 = lerpf(light_load(light, r00 + xs0), light_load(light, r00 + xs1), fx);
    const float c01 = lerpf(light_load(light, r01 + xs0), light_load(light, r01 + xs1), fx);
    const float c10 = lerpf(light_load(light, r10 + xs0), light_load(light, r10 + xs1), fx);
    const float c11 = lerpf(light_load(light, r11 + xs0), light_load(light, r11 + xs1), fx);
    return lerpf(lerpf(c00, c01, fy), lerpf(c10, c11, fy), fz);
}

// AccumulateWindowedRaymarchStep + AccumulateLightEnergy
template <typename DataT, typename LightT>
__device__ __forceinline__ void accumulate_step(const MarchUniforms& U, const DataT* __restrict__ data, const LightT* __restrict__ light,
                                                 const float4* s_tf, V3 p, float step, float4& acc) {
    const float v = sample_data<DataT>(data, U.ddims, p, U.data_wrap != 0);
    float pos;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tf_position(v, U.win, pos)) {
        int i0, i1;
        float f;
        tf_taps(pos, i0, i1, f);
        const float4 a = s_tf[i0], b = s_tf[i1];
        s = make_float4(lerpf(a.x, b.x, f), lerpf(a.y, b.y, f), lerpf(a.z, b.z, f), lerpf(a.w, b.w, f));
        s.w = step_opacity(s.w, step);
    }
    const float l = sample_light_wrap<LightT>(light, U.ldims, v3(saturatef(p.x), saturatef(p.y), saturatef(p.z)));
    s.x = s.x * l, s.y = s.y * l, s.z = s.z * l;
    const float oma = 1.0f - acc.w;
    acc.x = acc.x + ((s.x * s.w) * oma);
    acc.y = acc.y + ((s.y * s.w) * oma);
    acc.z = acc.z + ((s.z * s.w) * oma);
    acc.w = acc.w + (s.w * oma);
}

// One thread per pixel; a warp covers an 8x4 pixel tile so its rays stay coherent.
template <typename DataT, typename LightT>
__global__ void __launch_bounds__(256) raymarch_lit_kernel(const MarchUniforms U, const DataT* __restrict__ data,
                                                           const LightT* __restrict__ light, const float4* __restrict__ tf,
                                                           float4* __restrict__ out, unsigned long long* __restrict__ steps_out) {
    __shared__ float4 s_tf[256];
    s_tf[threadIdx.x] = __ldg(&tf[threadIdx.x]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ix = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int iy = U.row_begin + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    unsigned int steps = 0;
    if (ix < U.cam.width && iy < U.row_end) {
        const V3 V = camera_vector(U.cam, ix, iy);
        V3 cur, lcv;
        float thick;
        cube_setup(U.cam, V, cur, thick, lcv);
        const float ss = 1 / U.step_count;
        const float fas = U.step_count * thick;
        const float fl = floorf(fas);
        const int max_steps = (int) fl;
        const float fin = fas - fl;
        const V3 sv = v3(lcv.x * ss, lcv.y * ss, lcv.z * ss);
        const float ssw = 100.0f * ss;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (U.cam.jitter) {
            const float rnd = (float) pcg16_x(ix, iy, U.cam.frame_mod8) / 65535.0f;
            cur = v3(cur.x - sv.x * rnd, cur.y - sv.y * rnd, cur.z - sv.z * rnd);
        }
        int i = 0;
        for (i = 0; i < max_steps; i++) {
            cur = v3(cur.x + sv.x, cur.y + sv.y, cur.z + sv.z);
            ++steps;
            const float cd = dot3(cur.x - U.clip_center[0], cur.y - U.clip_center[1], cur.z - U.clip_center[2], U.clip_dir[0],
                                  U.clip_dir[1], U.clip_dir[2]);
            if (!(cd <= 0.0f)) {
                accumulate_step<DataT, LightT>(U, data, light, s_tf, cur, ssw, acc);
                if (acc.w > 0.95f) {
                    acc.w = 1.0f;
                    break;
                }
            }
        }
        if (i == max_steps && fin > 0.0f) {
            cur = v3(cur.x + sv.x * fin, cur.y + sv.y * fin, cur.z + sv.z * fin);
            ++steps;
            const float cd = dot3(cur.x - U.clip_center[0], cur.y - U.clip_center[1], cur.z - U.clip_center[2], U.clip_dir[0],
                                  U.clip_dir[1], U.clip_dir[2]);
            if (!(cd <= 0.0f)) accumulate_step<DataT, LightT>(U, data, light, s_tf, cur, 100.0f * fin, acc);
        }
        out[(size_t) (iy - U.row_begin) * U.cam.width + ix] = acc;
    }
    if (steps_out) {
        unsigned int s = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && s) atomicAdd(steps_out, (unsigned long long) s);
    }
}

__global__ void cube_setup_kernel(const RayCam cam, float4* __restrict__ out) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix >= cam.width || iy >= cam.height) return;
    const V3 V = camera_vector(cam, ix, iy);
    V3 entry, lcv;
    float thick;
    cube_setup(cam, V, entry, thick, lcv);
    out[(size_t) iy * cam.width + ix] = make_float4(entry.x, entry.y, entry.z, thick);
}

static RayCam to_raycam(const host::CameraUniforms& c) {
    RayCam r;
    static_assert(sizeof(RayCam) == sizeof(host::CameraUniforms), "camera uniform layouts must match");
    memcpy(&r, &c, sizeof(r));
    return r;
}

cudaError_t raymarch_cube_setup(tbrm_resources& r, const host::CameraUniforms& cam, float* d_out) {
    const dim3 block(32, 8), grid((cam.width + 31) / 32, (cam.height + 7) / 8);
    cube_setup_kernel<<<grid, block, 0, r.stream>>>(to_raycam(cam), (float4*) d_out);
    count_launch();
    return cudaGetLastError();
}

template <typename DataT, typename LightT>
static cudaError_t launch_lit(tbrm_resources& r, const MarchUniforms& U, float* d_out, unsigned long long* d_steps) {
    const int rows = U.row_end - U.row_begin;
    const dim3 grid((U.cam.width + 31) / 32, (rows + 7) / 8);
    raymarch_lit_kernel<DataT, LightT><<<grid, 256, 0, r.stream>>>(U, (const DataT*) r.data, (const LightT*) r.light, r.tf,
                                                                   (float4*) d_out, d_steps);
    count_launch();
    return cudaGetLastError();
}

cudaError_t raymarch_lit(tbrm_resources& r, const host::CameraUniforms& cam, const float clip_center[3], const float clip_dir[3],
                         float step_count, int row_begin, int row_end, float* d_out, unsigned long long* d_steps) {
    MarchUniforms U;
    U.cam = to_raycam(cam);
    for (int k = 0; k < 3; ++k) {
        U.ddims[k] = r.ddims[k];
        U.ldims[k] = r.ldims[k];
        U.clip_center[k] = clip_center[k];
        U.clip_dir[k] = clip_dir[k];
    }
    U.win = Windowing{r.windowing.center, r.windowing.width, r.windowing.low_cutoff ? 1.0f : 0.0f, r.windowing.high_cutoff ? 1.0f : 0.0f};
    U.step_count = step_count;
    U.row_begin = row_begin, U.row_end = row_end;
    U.data_wrap = r.options.data_addr_wrap;
    const bool l8 = r.light_fmt == TBRM_FMT_G8;
    switch (r.data_fmt) {
        case TBRM_FMT_G8:
            return l8 ? launch_lit<uint8_t, uint8_t>(r, U, d_out, d_steps) : launch_lit<uint8_t, float>(r, U, d_out, d_steps);
        case TBRM_FMT_G16:
            return l8 ? launch_lit<uint16_t, uint8_t>(r, U, d_out, d_steps) : launch_lit<uint16_t, float>(r, U, d_out, d_steps);
        default:
            return l8 ? launch_lit<float, uint8_t>(r, U, d_out, d_steps) : launch_lit<float, float>(r, U, d_out, d_steps);
    }
}

}  // namespace tbrm
