// raymarch.cu — per-pixel lit ray march (PerformRaymarchCubeSetup + PerformWindowedLitRaymarch), sm_100a.
// Reference: Source/Raymarcher/Shaders/Private/RaymarchMaterialCommon.usf:23-88,
//            Source/Raymarcher/Shaders/Private/WindowedRaymarchMaterials.usf:21-96.
#include <cstdlib>

#include <type_traits>

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
    int row_block, block_stride;  // local row lr renders image row row_begin + (lr / row_block) * row_block * block_stride + lr % row_block
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
    const int lr = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);  // row of the output buffer
    const int iy = U.row_begin + (lr / U.row_block) * U.row_block * U.block_stride + lr % U.row_block;
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
        out[(size_t) lr * U.cam.width + ix] = acc;
    }
    if (steps_out) {
        unsigned int s = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && s) atomicAdd(steps_out, (unsigned long long) s);
    }
}

// ---- fast path: U8 data, R32F light volume ---------------------------------------------------------------------
// Same arithmetic contract, bit-identical results; what changes is how the numbers are produced:
//   * UNORM8 decode and the TF-position division use exact division-free sequences (see sweep_tma_kernel.cuh);
//   * the Wrap sampler of the light volume (saturate(p) keeps the tap index in [-1, N-1]) wraps with two selects
//     instead of two integer modulos per axis;
//   * samples that contribute exactly nothing are not evaluated: a sample whose 8 taps all lie in a brick whose largest
//     byte is rejected by the low cut-off returns (0,0,0,0) (WindowedSampling.usf:28) — the brick max-grid makes that
//     test one byte load —, and a sample with zero opacity adds exactly 0 to every channel, so its light-volume fetch
//     is skipped. The march position still advances by the same sequence of fp32 adds.
constexpr int kBrick = 8;  // brick edge in voxels

struct FastUniforms {
    MarchUniforms M;
    const uint8_t* bricks;  // [bz][by][bx] max over voxels [8b, 8b+8] per axis, or nullptr
    int bdims[3];
    int skip_byte;          // largest byte the low cut-off rejects (-1: no skipping)
    float rwidth;
    int same_dims;          // the light volume has the data volume's dimensions (raymarch_fast2_kernel)
    float fddims[3], fldims[3];  // the dimensions as floats (axis_taps_bounded)
    float leap_margin;      // voxels kept clear of a brick face when leaping (> drift of the accumulated position)
};

__device__ __forceinline__ float decode_u8_exact(uint32_t v) {
    const float x = (float) v;
    return __fmaf_rn(x, 0.003921568859368563f, x * -2.319175823606301e-10f);  // == RN(v / 255)
}
__device__ __forceinline__ float div_exact(float x, float w, float rw) {  // == RN(x / w), rw = RN(1/w)
    const float q = x * rw;
    return __fmaf_rn(__fmaf_rn(-q, w, x), rw, q);
}

// ---- pipe-balanced addressing of the sample (ADDR32) --------------------------------------------------------------------------------
// The profile of the lit march (profiles/r1_raymarch_fast_kernel_ncu.txt) has the ALU pipe as the busiest unit (62 %) with the FMA pipes
// at 28 %: of the ~195 ALU-pipe instructions of a full sample about 90 form addresses (64-bit IADD3 / LEA / SHF pairs per tap), 12 clamp
// the tap index in float before it is clamped again as an integer, 6 convert the dimensions to float. ADDR32 moves that work to the
// FMA pipe or drops it: tap offsets are 32-bit integer multiply-adds (IMAD), a light tap's address is ONE widening multiply-add from the
// volume's base (IMAD.WIDE index * 4 + base) and a data tap's one 64-bit add of the uniform base (ptxas splits a widening multiply-add by
// 1 into exactly that, whatever is done to hide the 1), the float clamp is dropped where the integer clamp / wrap follows (march
// positions stay within [-1, 2], so the conversion cannot overflow), and the float dimensions come from the uniforms. Same taps, same
// weights, same arithmetic on the values: bit-identical to the 64-bit form (the host uses ADDR32 for volumes below 2^31 voxels).
// axis_taps without the float clamp of the tap index: for callers that clamp or wrap the integer index and whose u is bounded
__device__ __forceinline__ void axis_taps_bounded(float u, float fN, int& i0, float& f) {
    const float x = u * fN - 0.5f;
    const float fl = floorf(x);
    f = x - fl;
    i0 = (int) fl;
}

// a light-volume texel: R32F as stored, G8 (the reference's default light format, RaymarchVolume.h:198-199) decoded as UNORM8 — v / 255,
// correctly rounded, the generic kernel's light_load
template <typename LightT, typename I>
__device__ __forceinline__ float light_tap(const LightT* p, I i) {
    if constexpr (sizeof(LightT) == 1)
        return decode_u8_exact((uint32_t) __ldg(p + i));
    else
        return __ldg(p + i);
}

// one march sample (AccumulateWindowedRaymarchStep); returns nothing, updates acc
template <bool ADDR32, typename LightT>
__device__ __forceinline__ void fast_sample(const FastUniforms& F, const uint8_t* __restrict__ data, const LightT* __restrict__ light,
                                            const float4* s_tf, V3 p, float step, float4& acc) {
    const MarchUniforms& U = F.M;
    int i0, j0, k0;
    float fx, fy, fz;
    if (ADDR32) {
        axis_taps_bounded(p.x, F.fddims[0], i0, fx);
        axis_taps_bounded(p.y, F.fddims[1], j0, fy);
        axis_taps_bounded(p.z, F.fddims[2], k0, fz);
    } else {
        axis_taps(p.x, U.ddims[0], i0, fx);
        axis_taps(p.y, U.ddims[1], j0, fy);
        axis_taps(p.z, U.ddims[2], k0, fz);
    }
    const int xs0 = clamp_index(i0, U.ddims[0]), ys0 = clamp_index(j0, U.ddims[1]), zs0 = clamp_index(k0, U.ddims[2]);
    if (F.bricks) {
        const int m = __ldg(F.bricks + (xs0 >> 3) + F.bdims[0] * ((ys0 >> 3) + F.bdims[1] * (zs0 >> 3)));
        // every tap is rejected by the low cut-off, and with weights < 1 each lerp stays between its operands, so the
        // trilinear value is rejected too: the sample is exactly (0,0,0,0)
        if (m <= F.skip_byte && fx < 1.0f && fy < 1.0f && fz < 1.0f) return;
    }
    const int xs1 = clamp_index(i0 + 1, U.ddims[0]), ys1 = clamp_index(j0 + 1, U.ddims[1]), zs1 = clamp_index(k0 + 1, U.ddims[2]);
    uint32_t b000, b100, b010, b110, b001, b101, b011, b111;
    if (ADDR32) {
        const unsigned int X = (unsigned int) U.ddims[0], XY = X * (unsigned int) U.ddims[1];
        const unsigned int zx00 = (unsigned int) zs0 * XY + (unsigned int) xs0, zx01 = (unsigned int) zs0 * XY + (unsigned int) xs1;
        const unsigned int zx10 = (unsigned int) zs1 * XY + (unsigned int) xs0, zx11 = (unsigned int) zs1 * XY + (unsigned int) xs1;
        const unsigned int ya = (unsigned int) ys0 * X, yb = (unsigned int) ys1 * X;
        b000 = __ldg(data + (ya + zx00)), b100 = __ldg(data + (ya + zx01));
        b010 = __ldg(data + (yb + zx00)), b110 = __ldg(data + (yb + zx01));
        b001 = __ldg(data + (ya + zx10)), b101 = __ldg(data + (ya + zx11));
        b011 = __ldg(data + (yb + zx10)), b111 = __ldg(data + (yb + zx11));
    } else {
        const size_t X = U.ddims[0], XY = (size_t) U.ddims[0] * U.ddims[1];
        const uint8_t* r00 = data + X * ys0 + XY * zs0;
        const uint8_t* r01 = data + X * ys1 + XY * zs0;
        const uint8_t* r10 = data + X * ys0 + XY * zs1;
        const uint8_t* r11 = data + X * ys1 + XY * zs1;
        b000 = __ldg(r00 + xs0), b100 = __ldg(r00 + xs1), b010 = __ldg(r01 + xs0), b110 = __ldg(r01 + xs1);
        b001 = __ldg(r10 + xs0), b101 = __ldg(r10 + xs1), b011 = __ldg(r11 + xs0), b111 = __ldg(r11 + xs1);
    }
    // The light taps are requested NOW, before the data taps are even decoded: their addresses depend on the position only, and the chain
    // data taps -> interpolation -> window -> TF -> pow in between (~150 cycles of dependent arithmetic) hides their latency. Round 2's
    // profile has the march waiting on loads (long scoreboard 4.5 warps per issue-active cycle, a fifth of it on these eight taps). A
    // sample the window rejects, or whose opacity is exactly 0, has fetched them for nothing (L1 hits, 97 %): same values either way.
    float q000, q100, q010, q110, q001, q101, q011, q111;
    float gx, gy, gz;
    {
        int li, lj, lk;
        if (ADDR32) {
            axis_taps_bounded(saturatef(p.x), F.fldims[0], li, gx);
            axis_taps_bounded(saturatef(p.y), F.fldims[1], lj, gy);
            axis_taps_bounded(saturatef(p.z), F.fldims[2], lk, gz);
        } else {
            axis_taps(saturatef(p.x), U.ldims[0], li, gx);
            axis_taps(saturatef(p.y), U.ldims[1], lj, gy);
            axis_taps(saturatef(p.z), U.ldims[2], lk, gz);
        }
        const int LX = U.ldims[0], LY = U.ldims[1], LZ = U.ldims[2];
        const int x0 = li < 0 ? li + LX : li, x1 = li + 1 >= LX ? li + 1 - LX : li + 1;
        const int y0 = lj < 0 ? lj + LY : lj, y1 = lj + 1 >= LY ? lj + 1 - LY : lj + 1;
        const int z0 = lk < 0 ? lk + LZ : lk, z1 = lk + 1 >= LZ ? lk + 1 - LZ : lk + 1;
        if (ADDR32) {
            const unsigned int SX = (unsigned int) LX, SXY = SX * (unsigned int) LY;
            const unsigned int zx00 = (unsigned int) z0 * SXY + (unsigned int) x0, zx01 = (unsigned int) z0 * SXY + (unsigned int) x1;
            const unsigned int zx10 = (unsigned int) z1 * SXY + (unsigned int) x0, zx11 = (unsigned int) z1 * SXY + (unsigned int) x1;
            const unsigned int ya = (unsigned int) y0 * SX, yb = (unsigned int) y1 * SX;
            q000 = light_tap(light, ya + zx00), q100 = light_tap(light, ya + zx01), q010 = light_tap(light, yb + zx00), q110 = light_tap(light, yb + zx01);
            q001 = light_tap(light, ya + zx10), q101 = light_tap(light, ya + zx11), q011 = light_tap(light, yb + zx10), q111 = light_tap(light, yb + zx11);
        } else {
            const size_t SX = LX, SXY = (size_t) LX * LY;
            const LightT* q00 = light + SX * y0 + SXY * z0;
            const LightT* q01 = light + SX * y1 + SXY * z0;
            const LightT* q10 = light + SX * y0 + SXY * z1;
            const LightT* q11 = light + SX * y1 + SXY * z1;
            q000 = light_tap(q00, x0), q100 = light_tap(q00, x1), q010 = light_tap(q01, x0), q110 = light_tap(q01, x1);
            q001 = light_tap(q10, x0), q101 = light_tap(q10, x1), q011 = light_tap(q11, x0), q111 = light_tap(q11, x1);
        }
    }
    const float c00 = lerpf(decode_u8_exact(b000), decode_u8_exact(b100), fx);
    const float c01 = lerpf(decode_u8_exact(b010), decode_u8_exact(b110), fx);
    const float c10 = lerpf(decode_u8_exact(b001), decode_u8_exact(b101), fx);
    const float c11 = lerpf(decode_u8_exact(b011), decode_u8_exact(b111), fx);
    const float v = lerpf(lerpf(c00, c01, fy), lerpf(c10, c11, fy), fz);
    const float pos = div_exact(v - U.win.center + (U.win.width / 2.0f), U.win.width, F.rwidth);
    if ((pos < 0.0f && U.win.low > 0.0f) || (pos > 1.0f && U.win.high > 0.0f)) return;
    int t0, t1;
    float tf;
    tf_taps(pos, t0, t1, tf);
    const float4 a = s_tf[t0], b = s_tf[t1];
    const float alpha = step_opacity(lerpf(a.w, b.w, tf), step);
    if (alpha == 0.0f) return;  // adds (rgb * l * 0) * (1 - A) = 0 and 0 * (1 - A) = 0: nothing changes
    float sx = lerpf(a.x, b.x, tf), sy = lerpf(a.y, b.y, tf), sz = lerpf(a.z, b.z, tf);
    // light volume: trilinear, wrap addressing, at saturate(p)
    {
        const float d00 = lerpf(q000, q100, gx), d01 = lerpf(q010, q110, gx);
        const float d10 = lerpf(q001, q101, gx), d11 = lerpf(q011, q111, gx);
        const float l = lerpf(lerpf(d00, d01, gy), lerpf(d10, d11, gy), gz);
        sx = sx * l, sy = sy * l, sz = sz * l;
    }
    const float oma = 1.0f - acc.w;
    acc.x = acc.x + ((sx * alpha) * oma);
    acc.y = acc.y + ((sy * alpha) * oma);
    acc.z = acc.z + ((sz * alpha) * oma);
    acc.w = acc.w + (alpha * oma);
}

template <bool CLIP, bool ADDR32, typename LightT>
__global__ void __launch_bounds__(256) raymarch_fast_kernel(const FastUniforms F, const uint8_t* __restrict__ data,
                                                            const LightT* __restrict__ light, const float4* __restrict__ tf,
                                                            float4* __restrict__ out, unsigned long long* __restrict__ steps_out) {
    const MarchUniforms& U = F.M;
    __shared__ float4 s_tf[256];
    s_tf[threadIdx.x] = __ldg(&tf[threadIdx.x]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ix = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int lr = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);  // row of the output buffer
    const int iy = U.row_begin + (lr / U.row_block) * U.row_block * U.block_stride + lr % U.row_block;
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
            if (CLIP) {
                const float cd = dot3(cur.x - U.clip_center[0], cur.y - U.clip_center[1], cur.z - U.clip_center[2], U.clip_dir[0],
                                      U.clip_dir[1], U.clip_dir[2]);
                if (cd <= 0.0f) continue;
            }
            fast_sample<ADDR32, LightT>(F, data, light, s_tf, cur, ssw, acc);
            if (acc.w > 0.95f) {
                acc.w = 1.0f;
                break;
            }
        }
        steps = (unsigned) (i < max_steps ? i + 1 : max_steps);
        if (i == max_steps && fin > 0.0f) {
            cur = v3(cur.x + sv.x * fin, cur.y + sv.y * fin, cur.z + sv.z * fin);
            ++steps;
            bool clipped = false;
            if (CLIP) {
                const float cd = dot3(cur.x - U.clip_center[0], cur.y - U.clip_center[1], cur.z - U.clip_center[2], U.clip_dir[0],
                                      U.clip_dir[1], U.clip_dir[2]);
                clipped = cd <= 0.0f;
            }
            if (!clipped) fast_sample<ADDR32, LightT>(F, data, light, s_tf, cur, 100.0f * fin, acc);
        }
        out[(size_t) lr * U.cam.width + ix] = acc;
    }
    if (steps_out) {
        unsigned int s = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && s) atomicAdd(steps_out, (unsigned long long) s);
    }
}

// ---- second-generation fast path (same values, fewer instructions per step) ------------------------------------------
// What the profile of raymarch_fast_kernel asks for (ALU pipe busiest, ~177 thread-instructions per executed step):
//   * INTERIOR samples (1 <= tap index <= N-2 on every axis — all but a one-voxel shell): clamp (data sampler) and wrap (light
//     sampler) addressing are identities and saturate(p) == p, so when the light volume has the data volume's dimensions both
//     samplers share ONE set of tap indices, weights and 32-bit voxel offsets;
//   * EMPTY bricks are leapt: after a sample proves its brick empty (brick max rejected by the low cut-off), the number of
//     further march positions that provably stay inside that brick is computed from the ray's per-axis voxel step (with a margin
//     larger than the drift of the accumulated fp32 position), and the position is advanced by exactly that many
//     `cur += stepVec` adds — the same fp32 adds the per-step loop performs — without touching memory. Each skipped sample
//     would have returned (0,0,0,0) exactly, so colour, alpha, early-out and the executed-step count are unchanged.
constexpr int kMaxLeap = 64;

// axis_taps without the clamp that keeps the float -> int conversion defined for wild inputs. CUDA's conversion saturates (NaN -> 0), and a
// clamp that would have acted (floor outside [-4, N + 4]) leaves the raw index outside [1, N - 2]: raymarch_fast2_kernel then sends the
// sample to the general sampler, which recomputes its taps WITH the clamp. For interior samples index and weight are the same numbers.
__device__ __forceinline__ void axis_taps_raw(float u, int N, int& i0, float& f) {
    const float x = u * (float) N - 0.5f;
    const float fl = floorf(x);
    f = x - fl;
    i0 = (int) fl;
}

template <bool CLIP, typename LightT, int NT = 256, int WXP = 4, int WW = 8>
__global__ void __launch_bounds__(NT, 1280 / NT) raymarch_fast2_kernel(const FastUniforms F, const uint8_t* __restrict__ data,
                                                             const LightT* __restrict__ light, const float4* __restrict__ tf,
                                                             float4* __restrict__ out, unsigned long long* __restrict__ steps_out) {
    const MarchUniforms& U = F.M;
    __shared__ float4 s_tf[256];
    for (int c = threadIdx.x; c < 256; c += NT) s_tf[c] = __ldg(&tf[c]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WX = NT >= 128 ? WXP : NT / 32;  // warps across a block
    constexpr int WH = 32 / WW;  // a warp covers WW x WH pixels
    const int ix = blockIdx.x * (WW * WX) + (warp % WX) * WW + (lane % WW);
    const int lr = blockIdx.y * (NT / 32 / WX * WH) + (warp / WX) * WH + (lane / WW);  // row of the output buffer
    const int iy = U.row_begin + (lr / U.row_block) * U.row_block * U.block_stride + lr % U.row_block;
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
        const int X = U.ddims[0], Y = U.ddims[1], Z = U.ddims[2];
        const int XY = X * Y;
        // leaping: per-axis step in voxels and its reciprocal (approximations: the bound below carries a margin)
        const float dqx = sv.x * (float) X, dqy = sv.y * (float) Y, dqz = sv.z * (float) Z;
        const float rqx = dqx != 0.0f ? 1.0f / fabsf(dqx) : 3.0e38f, rqy = dqy != 0.0f ? 1.0f / fabsf(dqy) : 3.0e38f,
                    rqz = dqz != 0.0f ? 1.0f / fabsf(dqz) : 3.0e38f;
        const float margin = F.leap_margin;
        int i = 0;
        for (i = 0; i < max_steps; i++) {
            cur = v3(cur.x + sv.x, cur.y + sv.y, cur.z + sv.z);
            if (CLIP) {
                const float cd = dot3(cur.x - U.clip_center[0], cur.y - U.clip_center[1], cur.z - U.clip_center[2], U.clip_dir[0],
                                      U.clip_dir[1], U.clip_dir[2]);
                if (cd <= 0.0f) continue;
            }
            int i0, j0, k0;
            float fx, fy, fz;
            axis_taps_raw(cur.x, X, i0, fx);
            axis_taps_raw(cur.y, Y, j0, fy);
            axis_taps_raw(cur.z, Z, k0, fz);
            // unsigned arithmetic: a saturated index must not overflow; a side of 1 or 2 voxels has no interior
            const bool interior = ((unsigned) i0 - 1u) < (unsigned) max(X - 2, 0) && ((unsigned) j0 - 1u) < (unsigned) max(Y - 2, 0) &&
                                  ((unsigned) k0 - 1u) < (unsigned) max(Z - 2, 0);
            if (!interior || !F.same_dims) {  // the one-voxel shell, half-resolution light volumes: the general sampler
                fast_sample<true, LightT>(F, data, light, s_tf, cur, ssw, acc);
            } else {
                if (F.bricks) {
                    const int m = __ldg(F.bricks + (i0 >> 3) + F.bdims[0] * ((j0 >> 3) + F.bdims[1] * (k0 >> 3)));
                    if (m <= F.skip_byte) {  // tap indices >= 1 make the weights exact and < 1: the sample is exactly (0,0,0,0)
                        if (!CLIP && margin >= 0.0f) {
                            // positions i+1 .. i+n stay inside this brick: q + t*dq in [8b + margin, 8b + 8 - margin) on every axis
                            const float qx = (float) i0 + fx, qy = (float) j0 + fy, qz = (float) k0 + fz;
                            const float bx = (float) (i0 & ~7), by = (float) (j0 & ~7), bz = (float) (k0 & ~7);
                            const float nx = (dqx > 0.0f ? (bx + 8.0f - margin) - qx : qx - (bx + margin)) * rqx;
                            const float ny = (dqy > 0.0f ? (by + 8.0f - margin) - qy : qy - (by + margin)) * rqy;
                            const float nz = (dqz > 0.0f ? (bz + 8.0f - margin) - qz : qz - (bz + margin)) * rqz;
                            const float nf = fminf(fminf(nx, ny), fminf(nz, (float) kMaxLeap));
                            int n = nf > 1.0f ? (int) nf - 1 : 0;  // one more step of slack on top of the margin
                            n = min(n, max_steps - 1 - i);
                            for (int t = 0; t < n; ++t) cur = v3(cur.x + sv.x, cur.y + sv.y, cur.z + sv.z);
                            i += max(n, 0);
                        }
                        continue;
                    }
                }
                const int o = i0 + X * j0 + XY * k0;  // < 2^31 voxels: host check
                const uint8_t* d0 = data + o;
                const uint32_t b000 = __ldg(d0), b100 = __ldg(d0 + 1), b010 = __ldg(d0 + X), b110 = __ldg(d0 + X + 1);
                const uint32_t b001 = __ldg(d0 + XY), b101 = __ldg(d0 + XY + 1), b011 = __ldg(d0 + XY + X), b111 = __ldg(d0 + XY + X + 1);
                // light sampler: interior => saturate(p) == p and wrap == identity, same dimensions => same taps and weights. Requested before
                // the data taps are decoded: the window / TF / pow chain in between hides their latency (this kernel waits on loads: long
                // scoreboard 6.4 warps per issue-active cycle at 62 % issue utilisation, profiles/r2_raymarch_fast2_kernel_ncu.txt)
                const LightT* l0 = light + o;
                const float q000 = light_tap(l0, 0), q100 = light_tap(l0, 1), q010 = light_tap(l0, X), q110 = light_tap(l0, X + 1);
                const float q001 = light_tap(l0, XY), q101 = light_tap(l0, XY + 1), q011 = light_tap(l0, XY + X), q111 = light_tap(l0, XY + X + 1);
                const float c00 = lerpf(decode_u8_exact(b000), decode_u8_exact(b100), fx);
                const float c01 = lerpf(decode_u8_exact(b010), decode_u8_exact(b110), fx);
                const float c10 = lerpf(decode_u8_exact(b001), decode_u8_exact(b101), fx);
                const float c11 = lerpf(decode_u8_exact(b011), decode_u8_exact(b111), fx);
                const float v = lerpf(lerpf(c00, c01, fy), lerpf(c10, c11, fy), fz);
                const float pos = div_exact(v - U.win.center + (U.win.width / 2.0f), U.win.width, F.rwidth);
                if ((pos < 0.0f && U.win.low > 0.0f) || (pos > 1.0f && U.win.high > 0.0f)) continue;
                int t0, t1;
                float tfw;
                tf_taps(pos, t0, t1, tfw);
                const float4 a = s_tf[t0], b = s_tf[t1];
                const float alpha = step_opacity(lerpf(a.w, b.w, tfw), ssw);
                if (alpha == 0.0f) continue;  // adds exactly 0 to every channel
                float sx = lerpf(a.x, b.x, tfw), sy = lerpf(a.y, b.y, tfw), sz = lerpf(a.z, b.z, tfw);
                const float d00 = lerpf(q000, q100, fx), d01 = lerpf(q010, q110, fx);
                const float d10 = lerpf(q001, q101, fx), d11 = lerpf(q011, q111, fx);
                const float l = lerpf(lerpf(d00, d01, fy), lerpf(d10, d11, fy), fz);
                sx = sx * l, sy = sy * l, sz = sz * l;
                const float oma = 1.0f - acc.w;
                acc.x = acc.x + ((sx * alpha) * oma);
                acc.y = acc.y + ((sy * alpha) * oma);
                acc.z = acc.z + ((sz * alpha) * oma);
                acc.w = acc.w + (alpha * oma);
            }
            if (acc.w > 0.95f) {
                acc.w = 1.0f;
                break;
            }
        }
        steps = (unsigned) (i < max_steps ? i + 1 : max_steps);
        if (i == max_steps && fin > 0.0f) {
            cur = v3(cur.x + sv.x * fin, cur.y + sv.y * fin, cur.z + sv.z * fin);
            ++steps;
            bool clipped = false;
            if (CLIP) {
                const float cd = dot3(cur.x - U.clip_center[0], cur.y - U.clip_center[1], cur.z - U.clip_center[2], U.clip_dir[0],
                                      U.clip_dir[1], U.clip_dir[2]);
                clipped = cd <= 0.0f;
            }
            if (!clipped) fast_sample<true, LightT>(F, data, light, s_tf, cur, 100.0f * fin, acc);
        }
        out[(size_t) lr * U.cam.width + ix] = acc;
    }
    if (steps_out) {
        unsigned int s = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && s) atomicAdd(steps_out, (unsigned long long) s);
    }
}

// brick max-grid: bricks[b] = max data byte over voxels [8b, 8b + 8] on every axis (one voxel of apron on the far side:
// a sample whose first tap lies in brick b has its second tap at most one voxel further)
__global__ void brick_max_kernel(const uint8_t* __restrict__ data, int X, int Y, int Z, int BX, int BY, int BZ, uint8_t* __restrict__ out) {
    const int b = blockIdx.x;
    const int bx = b % BX, by = (b / BX) % BY, bz = b / (BX * BY);
    unsigned int m = 0;
    for (int t = threadIdx.x; t < 9 * 9 * 9; t += blockDim.x) {
        const int x = bx * kBrick + t % 9, y = by * kBrick + (t / 9) % 9, z = bz * kBrick + t / 81;
        if (x < X && y < Y && z < Z) m = max(m, (unsigned int) __ldg(data + (size_t) x + (size_t) X * ((size_t) y + (size_t) Y * z)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ unsigned int s_m[8];
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int) (blockDim.x >> 5); ++w) m = max(m, s_m[w]);
        out[b] = (uint8_t) m;
    }
}

// The same grid for R8 volumes with X % 16 == 0 in two launches that read the volume once with 16-byte loads (the kernel above issues 729
// byte loads per brick): a 256-thread CTA owns a 128 x 8 x 8 strip = 16 bricks along x; a thread reduces its 16 voxels of one row to the
// maxima of the two bricks they belong to and folds them into eight partial maxima per brick in shared memory — the whole brick (C), its
// x = 0 / y = 0 / z = 0 faces (Fx, Fy, Fz), the three edges at the origin corner (Exy, Exz, Eyz) and the corner voxel (K). The apron of
// brick b is the near face / edge / corner of its seven neighbours on the far side, so a second, tiny launch combines
// out[b] = max(C[b], Fx[b+x], Fy[b+y], Fz[b+z], Exy[b+x+y], Exz[b+x+z], Eyz[b+y+z], K[b+x+y+z]).
enum { kPartC = 0, kPartFx, kPartFy, kPartFz, kPartExy, kPartExz, kPartEyz, kPartK, kBrickParts };

__device__ __forceinline__ unsigned int max_byte_of(unsigned int a, unsigned int b) {  // max over the 8 bytes of two words
    const unsigned int m = __vmaxu4(a, b);
    const unsigned int h = __vmaxu4(m, m >> 16);
    return max(h & 0xffu, (h >> 8) & 0xffu);
}

__global__ void __launch_bounds__(256) brick_parts_kernel(const uint8_t* __restrict__ data, int X, int Y, int Z, int BX, int BY, int BZ,
                                                          uint8_t* __restrict__ parts) {
    // per round (4 z slices), warp (4 rows of one slice) and 16-voxel segment: the packed maxima {brick 2seg, brick 2seg+1, their x = 0
    // voxels} over the warp's 4 rows (s_rows) and of its first row alone (s_row0) — no atomics
    __shared__ unsigned int s_rows[2][8][8], s_row0[2][8][8];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, seg = t & 7, ly = (t >> 3) & 7, lzq = t >> 6;
    const int x = blockIdx.x * 128 + 16 * seg, y = blockIdx.y * 8 + ly;
    uint4 w[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int z = blockIdx.z * 8 + 4 * it + lzq;
        w[it] = make_uint4(0u, 0u, 0u, 0u);  // voxels beyond the volume do not exist: 0 never raises a maximum
        if (x < X && y < Y && z < Z) w[it] = __ldg(reinterpret_cast<const uint4*>(data + (size_t) x + (size_t) X * ((size_t) y + (size_t) Y * z)));
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        unsigned int p = max_byte_of(w[it].x, w[it].y) | (max_byte_of(w[it].z, w[it].w) << 8) | ((w[it].x & 0xffu) << 16) | ((w[it].z & 0xffu) << 24);
        if (lane < 8) s_row0[it][warp][seg] = p;  // lanes 0..7 hold the warp's first row (ly = 0 or 4)
        p = __vmaxu4(p, __shfl_xor_sync(0xffffffffu, p, 8));
        p = __vmaxu4(p, __shfl_xor_sync(0xffffffffu, p, 16));
        if (lane < 8) s_rows[it][warp][seg] = p;
    }
    __syncthreads();
    if (t < kBrickParts * 16) {
        const int part = t >> 4, b = t & 15, sg = b >> 1;
        // which byte of the packed word, which warps (bit 0 of the warp: rows 0-3 / 4-7; bits 1-2: slice within the round), which rounds
        const bool first = part == kPartFx || part == kPartExy || part == kPartExz || part == kPartK;
        const bool y0 = part == kPartFy || part == kPartExy || part == kPartEyz || part == kPartK;   // only the row ly = 0
        const bool z0 = part == kPartFz || part == kPartExz || part == kPartEyz || part == kPartK;   // only the slice lz = 0
        const int shift = 8 * ((b & 1) + (first ? 2 : 0));
        unsigned int m = 0;
        for (int it = 0; it < (z0 ? 1 : 2); ++it)
            for (int wp = 0; wp < (z0 ? 2 : 8); ++wp) {
                if (y0 && (wp & 1)) continue;
                m = max(m, ((y0 ? s_row0[it][wp][sg] : s_rows[it][wp][sg]) >> shift) & 0xffu);
            }
        const int bx = blockIdx.x * 16 + b;
        if (bx < BX) parts[(size_t) part * BX * BY * BZ + (size_t) bx + (size_t) BX * (blockIdx.y + (size_t) BY * blockIdx.z)] = (uint8_t) m;
    }
}

__global__ void __launch_bounds__(256) brick_combine_kernel(const uint8_t* __restrict__ parts, int BX, int BY, int BZ, uint8_t* __restrict__ out) {
    const size_t n = (size_t) BX * BY * BZ;
    const size_t b = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const int bx = (int) (b % BX), by = (int) ((b / BX) % BY), bz = (int) (b / ((size_t) BX * BY));
    const bool hx = bx + 1 < BX, hy = by + 1 < BY, hz = bz + 1 < BZ;  // a neighbour on the far side exists
    const size_t sx = 1, sy = (size_t) BX, sz = (size_t) BX * BY;
    unsigned int m = parts[kPartC * n + b];
    if (hx) m = max(m, (unsigned int) parts[kPartFx * n + b + sx]);
    if (hy) m = max(m, (unsigned int) parts[kPartFy * n + b + sy]);
    if (hz) m = max(m, (unsigned int) parts[kPartFz * n + b + sz]);
    if (hx && hy) m = max(m, (unsigned int) parts[kPartExy * n + b + sx + sy]);
    if (hx && hz) m = max(m, (unsigned int) parts[kPartExz * n + b + sx + sz]);
    if (hy && hz) m = max(m, (unsigned int) parts[kPartEyz * n + b + sy + sz]);
    if (hx && hy && hz) m = max(m, (unsigned int) parts[kPartK * n + b + sx + sy + sz]);
    out[b] = (uint8_t) m;
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
    const int rows = raymarch_local_rows(U.row_begin, U.row_end, U.row_block, U.block_stride);
    const dim3 grid((U.cam.width + 31) / 32, (rows + 7) / 8);
    raymarch_lit_kernel<DataT, LightT><<<grid, 256, 0, r.stream>>>(U, (const DataT*) r.data, (const LightT*) r.light, r.tf,
                                                                   (float4*) d_out, d_steps);
    count_launch();
    return cudaGetLastError();
}

// largest byte the low cut-off rejects, with the shader's fp32 arithmetic (-1: none / cut-off disabled)
static int low_cut_byte(const Windowing& w) {
    if (!(w.low > 0.0f)) return -1;
    int lo = -1;
    for (int b = 0; b < 256; ++b) {
        const float v = (float) b / 255.0f;
        const float pos = (v - w.center + (w.width / 2.0f)) / w.width;
        if (pos < 0.0f) {
            if (lo != b - 1) return -1;  // not a prefix of the byte range: no skipping
            lo = b;
        }
    }
    return lo;
}

cudaError_t ensure_bricks(tbrm_resources& r) {
    if (r.bricks_valid) return cudaSuccess;
    const int BX = (r.ddims[0] + kBrick - 1) / kBrick, BY = (r.ddims[1] + kBrick - 1) / kBrick, BZ = (r.ddims[2] + kBrick - 1) / kBrick;
    cudaError_t e;
    const size_t nb = (size_t) BX * BY * BZ;
    if (!r.bricks && (e = cudaMalloc(&r.bricks, nb * (1 + kBrickParts))) != cudaSuccess) return e;  // the grid, then the 8 partial grids
    if ((r.ddims[0] & 15) == 0 && (reinterpret_cast<uintptr_t>(r.data) & 15) == 0) {
        uint8_t* parts = (uint8_t*) r.bricks + nb;
        brick_parts_kernel<<<dim3((r.ddims[0] + 127) / 128, BY, BZ), 256, 0, r.stream>>>((const uint8_t*) r.data, r.ddims[0], r.ddims[1], r.ddims[2], BX,
                                                                                       BY, BZ, parts);
        brick_combine_kernel<<<(unsigned int) ((nb + 255) / 256), 256, 0, r.stream>>>(parts, BX, BY, BZ, (uint8_t*) r.bricks);
        count_launch(2);
    } else {
        brick_max_kernel<<<BX * BY * BZ, 128, 0, r.stream>>>((const uint8_t*) r.data, r.ddims[0], r.ddims[1], r.ddims[2], BX, BY, BZ,
                                                             (uint8_t*) r.bricks);
        count_launch();
    }
    const cudaError_t le = cudaGetLastError();
    r.bricks_valid = le == cudaSuccess;  // a failed launch must not leave a grid of garbage marked usable (samples would be skipped)
    return le;
}

// conservative: the clip plane never rejects a march position (positions stay within the unit cube expanded by 1)
static bool clip_never_rejects(const float c[3], const float d[3]) {
    double dmin = 0.0;
    for (int a = 0; a < 3; ++a) {
        const double lo = (-1.0 - (double) c[a]) * d[a], hi = (2.0 - (double) c[a]) * d[a];
        dmin += lo < hi ? lo : hi;
    }
    return dmin > 1.0;
}

// rows of the output buffer when blocks of `row_block` rows, every `block_stride`-th one starting at row_begin, are rendered
int raymarch_local_rows(int row_begin, int row_end, int row_block, int block_stride) {
    int rows = 0;
    for (int b = row_begin; b < row_end; b += row_block * block_stride) rows += std::min(row_block, row_end - b);
    return rows;
}

cudaError_t raymarch_lit(tbrm_resources& r, const host::CameraUniforms& cam, const float clip_center[3], const float clip_dir[3],
                         float step_count, int row_begin, int row_end, int row_block, int block_stride, float* d_out,
                         unsigned long long* d_steps) {
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
    U.row_block = row_block, U.block_stride = block_stride;
    U.data_wrap = r.options.data_addr_wrap;
    const bool l8 = r.light_fmt == TBRM_FMT_G8;
    if (r.data_fmt == TBRM_FMT_G8 && !U.data_wrap && r.options.reserved[1] != 1) {  // both light formats (round 2: G8 took the generic kernel, 22 ms)
        FastUniforms F;
        F.M = U;
        F.rwidth = 1.0f / U.win.width;
        F.skip_byte = low_cut_byte(U.win);
        F.bricks = nullptr;
        for (int k = 0; k < 3; ++k) F.bdims[k] = (r.ddims[k] + kBrick - 1) / kBrick;
        if (F.skip_byte >= 0) {
            cudaError_t e = ensure_bricks(r);
            if (e != cudaSuccess) return e;
            F.bricks = (const uint8_t*) r.bricks;
        }
        const int rows = raymarch_local_rows(row_begin, row_end, row_block, block_stride);
        const dim3 grid((cam.width + 31) / 32, (rows + 7) / 8);
        F.same_dims = r.ldims[0] == r.ddims[0] && r.ldims[1] == r.ddims[1] && r.ldims[2] == r.ddims[2];
        const int nmax = std::max(r.ddims[0], std::max(r.ddims[1], r.ddims[2]));
        // accumulated positions drift by < kMaxLeap half-ulps of values in [-1, 2] (1.2e-7 each) = 7.7e-6 UVW = 7.7e-6 * N voxels
        F.leap_margin = std::max(0.01f, 1.0e-5f * (float) nmax);
        // reserved[1]: 0 default, 1 generic kernel, 2 fast kernel with the 64-bit tap addressing round 1 measured, 3 second-generation fast kernel
        for (int k = 0; k < 3; ++k) F.fddims[k] = (float) r.ddims[k], F.fldims[k] = (float) r.ldims[k];
        static const bool addr64_forced = [] { const char* e = getenv("TBRM_RAYMARCH_ADDR64"); return e && e[0] == '1'; }();
        const bool addr32 = r.options.reserved[1] != 2 && !addr64_forced && r.data_voxels() < (1ull << 31) && r.light_voxels() < (1ull << 31);
        // Kernel choice (every form returns the same bits). Default: the second-generation kernel WITHOUT leaping — measured on a B200 in
        // round 2 (cfg2 frame / the same frame of a 256^3 volume): first generation 4.61 / 4.00 ms, second generation with leaps 5.00 / 3.28 ms
        // (13 % fewer instructions, but the leap loop diverges: 27.5 active lanes instead of 31.4, issue slots 60 % instead of 74 %), second
        // generation without leaps 4.53 / 3.24 ms. TBRM_RAYMARCH_V2 = 0 / 1 / 2 or reserved[1] = 5 / 3 / 4 select first generation / leaps / no leaps.
        static const int v2_env = [] { const char* e = getenv("TBRM_RAYMARCH_V2"); return e ? atoi(e) : 2; }();
        const int form = r.options.reserved[1] == 3 ? 1 : (r.options.reserved[1] == 4 ? 2 : (r.options.reserved[1] == 5 ? 0 : (r.options.reserved[1] == 0 ? v2_env : 0)));
        const bool v2 = form != 0 && r.data_voxels() < (1ull << 31);
        if (form == 2) F.leap_margin = -1.0f;  // interior fast path only, no leaps
        const bool noclip = clip_never_rejects(clip_center, clip_dir);
        const uint8_t* d8 = (const uint8_t*) r.data;
        float4* o4 = (float4*) d_out;
        auto launch = [&](auto light_ptr) {  // light_ptr: const float* (R32F) or const uint8_t* (G8)
            using LightT = std::remove_cv_t<std::remove_pointer_t<decltype(light_ptr)>>;
            // blocks of 128 threads instead of 256: the SM's slots refill in finer grains as rays end (cfg2 frame 4.54 -> 4.45 ms as 32 x 4 pixels;
            // 64 threads: the same). Shape, measured on one box at cfg2: warps of 4 x 8 pixels in blocks of 2 x 2 warps (8 x 16 pixels) 4.35-4.39 ms,
            // 8 x 4 warps 4.45 ms, 16 x 2 4.94 ms, 2 x 16 4.85 ms, blocks of 4 x 1 or 1 x 4 warps 4.39-4.42 ms. TBRM_MARCH_NT=256 keeps round 1's
            // 32 x 8 blocks of 8 x 4 warps.
            static const bool nt256 = [] { const char* e = getenv("TBRM_MARCH_NT"); return e && atoi(e) == 256; }();
            const dim3 grid128((cam.width + 7) / 8, (rows + 15) / 16);
            if (v2 && noclip && !nt256)
                raymarch_fast2_kernel<false, LightT, 128, 2, 4><<<grid128, 128, 0, r.stream>>>(F, d8, light_ptr, r.tf, o4, d_steps);
            else if (v2 && !nt256)
                raymarch_fast2_kernel<true, LightT, 128, 2, 4><<<grid128, 128, 0, r.stream>>>(F, d8, light_ptr, r.tf, o4, d_steps);
            else if (v2 && noclip)
                raymarch_fast2_kernel<false, LightT><<<grid, 256, 0, r.stream>>>(F, d8, light_ptr, r.tf, o4, d_steps);
            else if (v2)
                raymarch_fast2_kernel<true, LightT><<<grid, 256, 0, r.stream>>>(F, d8, light_ptr, r.tf, o4, d_steps);
            else if (noclip && addr32)
                raymarch_fast_kernel<false, true, LightT><<<grid, 256, 0, r.stream>>>(F, d8, light_ptr, r.tf, o4, d_steps);
            else if (addr32)
                raymarch_fast_kernel<true, true, LightT><<<grid, 256, 0, r.stream>>>(F, d8, light_ptr, r.tf, o4, d_steps);
            else if (noclip)
                raymarch_fast_kernel<false, false, LightT><<<grid, 256, 0, r.stream>>>(F, d8, light_ptr, r.tf, o4, d_steps);
            else
                raymarch_fast_kernel<true, false, LightT><<<grid, 256, 0, r.stream>>>(F, d8, light_ptr, r.tf, o4, d_steps);
        };
        if (l8)
            launch((const uint8_t*) r.light);
        else
            launch((const float*) r.light);
        count_launch();
        return cudaGetLastError();
    }
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

#include "materials.cuh"
