// tbrm_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY: nothing under tbraymarcherplugin_b200/ may
// import, link or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / baseline.
//
// What it is: a plain fp32 C++ restatement of the hot path of tommybazar/TBRaymarcherPlugin (paths relative
// to the plugin root), each function citing the file:line it follows:
//   Source/Raymarcher/Shaders/Private/{RaymarcherCommon,WindowedSampling,AddDirLightShader,ChangeDirLightShader,
//       RaymarchMaterialCommon,WindowedRaymarchMaterials}.usf, Source/FractalMarcher/Shaders/Private/SDFMarcher.usf,
//   Source/Raymarcher/Private/Rendering/{LightingShaderUtils,LightingShaders}.cpp.
//
// WHAT PINS IT: the reference as a whole (Unreal Engine 5.4 / D3D11 / HLSL) cannot be built or run here and its test module pins no
// numeric result — but its own source files can be compiled for the CPU from where they lie: oracle/ref.mk builds
// LightingShaderUtils.cpp, VolumeInfo.cpp and the TextureUtilities.h templates against oracle/ue_shim, and the shaders (streamed through
// the syntactic rewrites of oracle/hlsl2cpp.py) against oracle/hlsl_shim, into oracle/_ref/libtbrm_ref.so. This oracle equals that build
// bit for bit (tests/test_ref_pin_cpu.py, test_ref_shaders_cpu.py, test_ref_materials_cpu.py; committed vectors tests/golden/ref_*.npz).
// That pins the LOGIC of the restatement. The engine semantics under the shaders (D3D11 samplers, UNORM conversions, HLSL pow,
// FLinearColor / FTransform math) are not in /root/reference: they follow the policies of SURVEY.md Appendix B (Q1..Q10, marked "Qn"
// below), stated once in the shims and here — for those, PARITY STAYS UNPINNED. Closed-form known-answer tests: tests/test_oracle_kat.py.
//
// Arithmetic contract (shared with the CUDA kernels so that parity is bit-exact):
//   * every operation is a single correctly-rounded IEEE fp32 op in the order written here; compiled with
//     -ffp-contract=off so the compiler never fuses; fmaf() appears exactly where the contract says "fma";
//   * lerp(a,b,t) = fmaf(t, b - a, a);
//   * dot(a,b) = ((a.x*b.x) + (a.y*b.y)) + (a.z*b.z);  length(v) = sqrtf(dot(v,v));  normalize(v) = v / length(v);
//   * pow(x,y) = det_exp2(y * det_log2(x)): fixed polynomials (oracle/gen_pow_coeffs.py), abs error < 1e-7 vs libm
//     (HLSL pow is exp2(y*log2(x)) with implementation-defined precision, Q10);
//   * Mandelbulb transcendentals (acos, atan2, sin, cos, pow, log) use libm here and CUDA's fp32 library in the
//     kernel; that path is compared with a tolerance, not bit-exactly.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/tbrm.h"
#include "tbrm_oracle.h"

namespace {

// ------------------------------------------------------------------------------------------------------------
// small fp32 helpers (the arithmetic contract)
// ------------------------------------------------------------------------------------------------------------
struct F3 {
    float x, y, z;
};
inline F3 f3(float x, float y, float z) { return F3{x, y, z}; }
inline float lerpf(float a, float b, float t) { return fmaf(t, b - a, a); }
inline float dot3(F3 a, F3 b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); }
inline float length3(F3 a) { return sqrtf(dot3(a, a)); }
inline F3 normalize3(F3 a) {
    float l = length3(a);
    return f3(a.x / l, a.y / l, a.z / l);
}
inline float saturatef(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }  // saturate(NaN) = 0 like HLSL
inline float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

}  // namespace
#include "tbrm_contract.h"
namespace {
using namespace tbrm_contract;

// fp32 -> fp16 -> fp32 round trip, round-to-nearest-even (FFloat16 / PF_FloatRGBA texels, Q9)
inline float round_to_half(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint32_t sign = x & 0x80000000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return f;  // inf / nan
    float af;
    std::memcpy(&af, &ax, 4);
    if (af >= 65520.0f) {  // overflows to inf in fp16
        uint32_t inf = sign | 0x7f800000u;
        float r;
        std::memcpy(&r, &inf, 4);
        return r;
    }
    float r;
    if (af < 6.103515625e-05f) {  // fp16 subnormal range: quantum 2^-24
        float q = rintf(af * 16777216.0f) / 16777216.0f;  // rintf is RNE in the default rounding mode
        r = q;
    } else {
        // keep 10 mantissa bits, RNE
        uint32_t lsb = (ax >> 13) & 1u;
        uint32_t rounded = ax + 0x0fffu + lsb;
        rounded &= ~0x1fffu;
        std::memcpy(&r, &rounded, 4);
    }
    uint32_t rb;
    std::memcpy(&rb, &r, 4);
    rb |= sign;
    std::memcpy(&r, &rb, 4);
    return r;
}

// ------------------------------------------------------------------------------------------------------------
// texture model
// ------------------------------------------------------------------------------------------------------------
struct DataTex {
    const void* p;
    int fmt;
    int X, Y, Z;
    inline float texel(int x, int y, int z) const {
        size_t i = (size_t) x + (size_t) X * ((size_t) y + (size_t) Y * (size_t) z);  // TextureUtilities.cpp:43-78
        switch (fmt) {
            case TBRM_FMT_G8: return (float) ((const uint8_t*) p)[i] / 255.0f;
            case TBRM_FMT_G16: return (float) ((const uint16_t*) p)[i] / 65535.0f;
            default: return ((const float*) p)[i];
        }
    }
};

enum AddrMode { ADDR_CLAMP = 0, ADDR_WRAP = 1, ADDR_BORDER = 2 };

// one axis of SampleLevel(linear): x = u*N - 0.5, taps floor(x), floor(x)+1
inline void axis_taps(float u, int N, int& i0, float& f) {
    float x = u * (float) N - 0.5f;
    float fl = floorf(x);
    f = x - fl;
    // keep the int conversion defined for wild inputs (inf/nan/huge): such taps are out of range anyway
    fl = fminf(fmaxf(fl, -4.0f), (float) N + 4.0f);
    i0 = (int) fl;
}
inline int addr(int i, int N, AddrMode m) {
    if (m == ADDR_CLAMP) return i < 0 ? 0 : (i >= N ? N - 1 : i);
    if (m == ADDR_WRAP) {
        int r = i % N;
        return r < 0 ? r + N : r;
    }
    return i;  // border handled by caller
}

inline float sample_data(const DataTex& t, F3 uvw, AddrMode mode, float border) {
    int i0, j0, k0;
    float fx, fy, fz;
    axis_taps(uvw.x, t.X, i0, fx);
    axis_taps(uvw.y, t.Y, j0, fy);
    axis_taps(uvw.z, t.Z, k0, fz);
    float v[2][2][2];
    for (int dz = 0; dz < 2; ++dz)
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                int x = i0 + dx, y = j0 + dy, z = k0 + dz;
                if (mode == ADDR_BORDER) {
                    bool in = x >= 0 && x < t.X && y >= 0 && y < t.Y && z >= 0 && z < t.Z;
                    v[dz][dy][dx] = in ? t.texel(x, y, z) : border;
                } else {
                    v[dz][dy][dx] = t.texel(addr(x, t.X, mode), addr(y, t.Y, mode), addr(z, t.Z, mode));
                }
            }
    float c00 = lerpf(v[0][0][0], v[0][0][1], fx);
    float c01 = lerpf(v[0][1][0], v[0][1][1], fx);
    float c10 = lerpf(v[1][0][0], v[1][0][1], fx);
    float c11 = lerpf(v[1][1][0], v[1][1][1], fx);
    float c0 = lerpf(c00, c01, fy);
    float c1 = lerpf(c10, c11, fy);
    return lerpf(c0, c1, fz);
}

// UAV store / load of the light volume and the R/W buffers (A.1: UNORM8 store = round(clamp(v,0,1)*255))
inline uint8_t quant8(float v) {
    float c = saturatef(v);  // NaN -> 0
    return (uint8_t) floorf(c * 255.0f + 0.5f);
}
struct LightTex {
    void* p;
    int fmt;  // G8 or R32F
    int X, Y, Z;
    inline size_t idx(int x, int y, int z) const { return (size_t) x + (size_t) X * ((size_t) y + (size_t) Y * (size_t) z); }
    inline float load(int x, int y, int z) const {
        return fmt == TBRM_FMT_G8 ? (float) ((const uint8_t*) p)[idx(x, y, z)] / 255.0f : ((const float*) p)[idx(x, y, z)];
    }
    inline void store(int x, int y, int z, float v) const {
        if (fmt == TBRM_FMT_G8)
            ((uint8_t*) p)[idx(x, y, z)] = quant8(v);
        else
            ((float*) p)[idx(x, y, z)] = v;
    }
};
inline float sample_light_wrap(const LightTex& t, F3 uvw) {  // Material.Wrap_WorldGroupSettings (Q5)
    int i0, j0, k0;
    float fx, fy, fz;
    axis_taps(uvw.x, t.X, i0, fx);
    axis_taps(uvw.y, t.Y, j0, fy);
    axis_taps(uvw.z, t.Z, k0, fz);
    int xs[2] = {addr(i0, t.X, ADDR_WRAP), addr(i0 + 1, t.X, ADDR_WRAP)};
    int ys[2] = {addr(j0, t.Y, ADDR_WRAP), addr(j0 + 1, t.Y, ADDR_WRAP)};
    int zs[2] = {addr(k0, t.Z, ADDR_WRAP), addr(k0 + 1, t.Z, ADDR_WRAP)};
    float c00 = lerpf(t.load(xs[0], ys[0], zs[0]), t.load(xs[1], ys[0], zs[0]), fx);
    float c01 = lerpf(t.load(xs[0], ys[1], zs[0]), t.load(xs[1], ys[1], zs[0]), fx);
    float c10 = lerpf(t.load(xs[0], ys[0], zs[1]), t.load(xs[1], ys[0], zs[1]), fx);
    float c11 = lerpf(t.load(xs[0], ys[1], zs[1]), t.load(xs[1], ys[1], zs[1]), fx);
    return lerpf(lerpf(c00, c01, fy), lerpf(c10, c11, fy), fz);
}

// 2-D propagation buffer ("Illumination Buffer", RaymarchUtils.cpp:176-196) in the light volume's pixel format
struct Buf2D {
    std::vector<float> f;
    std::vector<uint8_t> q;
    int W = 0, H = 0, fmt = TBRM_FMT_R32F;
    void init(int w, int h, int format, float clear) {  // Clear2DTexture_RenderThread, UtilityShaders.cpp:57-75
        W = w;
        H = h;
        fmt = format;
        if (fmt == TBRM_FMT_G8)
            q.assign((size_t) w * h, quant8(clear));
        else
            f.assign((size_t) w * h, clear);
    }
    inline float load(int x, int y) const {
        size_t i = (size_t) x + (size_t) W * y;
        return fmt == TBRM_FMT_G8 ? (float) q[i] / 255.0f : f[i];
    }
    inline void store(int x, int y, float v) {
        size_t i = (size_t) x + (size_t) W * y;
        if (fmt == TBRM_FMT_G8)
            q[i] = quant8(v);
        else
            f[i] = v;
    }
    // SF_Bilinear + AM_Border (LightingShaderUtils.cpp:190-195)
    inline float sample_border(float u, float v, float border) const {
        int i0, j0;
        float fx, fy;
        axis_taps(u, W, i0, fx);
        axis_taps(v, H, j0, fy);
        auto tap = [&](int x, int y) { return (x >= 0 && x < W && y >= 0 && y < H) ? load(x, y) : border; };
        float top = lerpf(tap(i0, j0), tap(i0 + 1, j0), fx);
        float bot = lerpf(tap(i0, j0 + 1), tap(i0 + 1, j0 + 1), fx);
        return lerpf(top, bot, fy);
    }
};

// ------------------------------------------------------------------------------------------------------------
// WindowedSampling.usf
// ------------------------------------------------------------------------------------------------------------
// GetTransferFuncPosition — WindowedSampling.usf:14-17
inline float tf_position(float v, float center, float width) { return (v - center + (width / 2.0f)) / width; }

// TF.SampleLevel(TFSampler, float2(TFPos, 0.5), 0) on the collapsed 256x1 table (A.1), bilinear clamp
inline void tf_lookup(const float* tf, float pos, float out[4]) {
    float x = pos * 256.0f - 0.5f;
    float fl = floorf(x);
    float f = x - fl;
    int i0 = (int) fminf(fmaxf(fl, 0.0f), 255.0f);
    int i1 = (int) fminf(fmaxf(fl + 1.0f, 0.0f), 255.0f);
    for (int c = 0; c < 4; ++c) out[c] = lerpf(tf[4 * i0 + c], tf[4 * i1 + c], f);
}

// SampleWindowedTransferFunction — WindowedSampling.usf:20-37
inline void sample_windowed_tf(float value, float step, const float* tf, const tbrm_windowing& w, float out[4]) {
    float pos = tf_position(value, w.center, w.width);
    float lowc = w.low_cutoff ? 1.0f : 0.0f, highc = w.high_cutoff ? 1.0f : 0.0f;  // VolumeInfo.h:49-52
    if ((pos < 0.0f && lowc > 0.0f) || (pos > 1.0f && highc > 0.0f)) {
        out[0] = out[1] = out[2] = out[3] = 0.0f;
        return;
    }
    tf_lookup(tf, pos, out);
    out[3] = saturatef(out[3]);
    out[3] = 1.0f - det_pow(1.0f - out[3], step);
}

// ------------------------------------------------------------------------------------------------------------
// host parameter math — LightingShaderUtils.cpp (fp64 like UE5's FVector), SURVEY.md A.3
// ------------------------------------------------------------------------------------------------------------
struct D3 {
    double x, y, z;
};
inline D3 cross(D3 a, D3 b) { return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// FQuat::UnrotateVector (Q3): v' = v + w*t + q' x t with q' = -q.xyz, t = 2 (q' x v)
inline D3 unrotate(const double q[4], D3 v) {
    D3 qv{-q[0], -q[1], -q[2]};
    D3 t = cross(qv, v);
    t = D3{2.0 * t.x, 2.0 * t.y, 2.0 * t.z};
    D3 c = cross(qv, t);
    return D3{v.x + q[3] * t.x + c.x, v.y + q[3] * t.y + c.y, v.z + q[3] * t.z + c.z};
}
inline double safe_recip(double s) { return std::fabs(s) <= 1e-8 ? 0.0 : 1.0 / s; }  // GetSafeScaleReciprocal
inline void normalize_d(D3& v) {  // FVector::Normalize(1e-8)
    double sq = v.x * v.x + v.y * v.y + v.z * v.z;
    if (sq > 1e-8) {
        double s = 1.0 / std::sqrt(sq);
        v = D3{v.x * s, v.y * s, v.z * s};
    }
}
inline D3 inv_transform_vector(const tbrm_world& w, D3 v) {  // FTransform::InverseTransformVector
    D3 u = unrotate(w.rotation, v);
    return D3{u.x * safe_recip(w.scale[0]), u.y * safe_recip(w.scale[1]), u.z * safe_recip(w.scale[2])};
}
inline D3 inv_transform_position(const tbrm_world& w, D3 p) {  // FTransform::InverseTransformPosition
    D3 u = unrotate(w.rotation, D3{p.x - w.translation[0], p.y - w.translation[1], p.z - w.translation[2]});
    return D3{u.x * safe_recip(w.scale[0]), u.y * safe_recip(w.scale[1]), u.z * safe_recip(w.scale[2])};
}

inline double srgb_to_linear(double c) { return c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4); }
inline double linear_to_srgb(double c) { return c <= 0.0031308 ? c * 12.92 : 1.055 * std::pow(c, 1.0 / 2.4) - 0.055; }
inline double clamp01(double x) { return x < 0 ? 0 : (x > 1 ? 1 : x); }
// Q2: FLinearColor(I*w).ToFColor(true) -> packed -> RHI border colour
inline float light_border(float alpha, int exact) {
    if (exact) return alpha;
    double q = std::floor(linear_to_srgb(clamp01((double) alpha)) * 255.0 + 0.5) / 255.0;
    return (float) srgb_to_linear(q);
}
// Q1: FLinearColor(C - W/2).ToFColor(false) -> packed -> RHI border colour (LightingShaders.h:82-89)
inline float data_border(const tbrm_windowing& w, int exact) {
    float zero_tf = w.center - 0.5f * w.width;
    if (exact) return zero_tf;
    double q = std::floor(clamp01((double) zero_tf) * 255.0 + 0.5) / 255.0;
    return (float) srgb_to_linear(q);
}

const double kFaceNormal[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};  // LightingShaderUtils.h:21-44

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// exported: host plan
// ------------------------------------------------------------------------------------------------------------
extern "C" void tbo_prepare_tf(const float* rgba, int width, int height, float* out_256x4) {
    // PF_FloatRGBA texels (RaymarchUtils.cpp:120-136,151-170), sampled at v = 0.5 with a bilinear clamp sampler
    float y = 0.5f * (float) height - 0.5f;
    float fl = floorf(y), fy = y - fl;
    int j0 = (int) fminf(fmaxf(fl, 0.0f), (float) (height - 1));
    int j1 = (int) fminf(fmaxf(fl + 1.0f, 0.0f), (float) (height - 1));
    for (int i = 0; i < width && i < 256; ++i)
        for (int c = 0; c < 4; ++c) {
            float a = round_to_half(rgba[((size_t) j0 * width + i) * 4 + c]);
            float b = round_to_half(rgba[((size_t) j1 * width + i) * 4 + c]);
            out_256x4[4 * i + c] = lerpf(a, b, fy);
        }
}

extern "C" void tbo_default_tf(float* out_256x4) {  // MakeDefaultTFTexture, RaymarchUtils.cpp:113-141
    float rgba[256 * 4];
    for (unsigned i = 0; i < 256; ++i) {
        float w = (float) i / (float) (256 - 1);
        rgba[4 * i] = rgba[4 * i + 1] = rgba[4 * i + 2] = w;
        rgba[4 * i + 3] = 1.0f;
    }
    tbo_prepare_tf(rgba, 256, 1, out_256x4);
}

extern "C" int tbo_plan_dir_light(const int32_t ldims[3], const tbrm_windowing* win, int border_exact,
                                  const tbrm_dir_light* light, const tbrm_world* world, tbo_light_plan* out) {
    std::memset(out, 0, sizeof(*out));
    // LightingShaders.cpp:41-46 — zero direction: nothing happens
    if (light->direction[0] == 0.0 && light->direction[1] == 0.0 && light->direction[2] == 0.0) {
        out->zero_direction = 1;
        return 0;
    }
    // GetLocalLightParamsAndAxes — LightingShaderUtils.cpp:160-188
    D3 d = inv_transform_vector(*world, D3{light->direction[0], light->direction[1], light->direction[2]});
    normalize_d(d);
    D3 p{-d.x, -d.y, -d.z};  // light "position"
    // FMajorAxes::GetMajorAxes — LightingShaderUtils.cpp:29-46
    struct FW {
        int face;
        float w;
    } fw[6];
    for (int i = 0; i < 6; ++i) {
        float weight = (float) (kFaceNormal[i][0] * p.x + kFaceNormal[i][1] * p.y + kFaceNormal[i][2] * p.z);
        weight = (weight > 0 ? weight * weight : 0);
        fw[i] = FW{i, weight};
    }
    // std::sort with SortDescendingWeights is unstable on ties; policy: stable, lower face index first (A.3 step 2)
    std::stable_sort(fw, fw + 6, [](const FW& a, const FW& b) { return a.w > b.w; });
    if (fw[0].w > 0.99f) fw[0].w = 1.0f;  // :181-184
    fw[1].w = 1 - fw[0].w;                // :187

    // GetLocalClippingParameters — LightingShaderUtils.cpp:205-220
    D3 cc = inv_transform_position(*world, D3{world->clip.center[0], world->clip.center[1], world->clip.center[2]});
    cc = D3{cc.x + 0.5, cc.y + 0.5, cc.z + 0.5};
    D3 cd = unrotate(world->rotation, D3{world->clip.direction[0], world->clip.direction[1], world->clip.direction[2]});
    cd = D3{cd.x * world->scale[0], cd.y * world->scale[1], cd.z * world->scale[2]};
    normalize_d(cd);
    out->clip_center[0] = (float) cc.x, out->clip_center[1] = (float) cc.y, out->clip_center[2] = (float) cc.z;
    out->clip_dir[0] = (float) cd.x, out->clip_dir[1] = (float) cd.y, out->clip_dir[2] = (float) cd.z;
    out->data_border = data_border(*win, border_exact);
    out->local_dir[0] = d.x, out->local_dir[1] = d.y, out->local_dir[2] = d.z;

    for (int i = 0; i < 2; ++i) {
        tbo_pass& P = out->pass[i];
        P.face = fw[i].face;
        P.axis = P.face / 2;
        P.weight = fw[i].w;
        // GetTransposedDimensions — :48-64
        const int X = ldims[0], Y = ldims[1], Z = ldims[2];
        if (P.axis == 0)
            P.td[0] = Y, P.td[1] = Z, P.td[2] = X;
        else if (P.axis == 1)
            P.td[0] = X, P.td[1] = Z, P.td[2] = Y;
        else
            P.td[0] = X, P.td[1] = Y, P.td[2] = Z;
        // GetAxisDirection / GetLoopStartStopIndexes — :66-70, 251-265
        P.dirn = (P.face % 2) ? 1 : -1;
        if (P.dirn == -1)
            P.start = P.td[2] - 1, P.stop = -1;
        else
            P.start = 0, P.stop = P.td[2];
        // GetLightAlpha — :222-225 ; border colour — :197-203
        P.light_alpha = light->intensity * fw[i].w;
        P.border = light_border(P.light_alpha, border_exact);
        // GetUVOffset — :82-129 (FVector /= scalar multiplies by the reciprocal)
        double pa = (P.axis == 0 ? p.x : (P.axis == 1 ? p.y : p.z));
        if (pa == 0.0) {
            // Exactly axis-aligned light: the reference divides by zero here and feeds inf/NaN offsets to a pass whose
            // light alpha is 0 (weight 0), i.e. a pass that contributes nothing on hardware that maps NaN sampler
            // coordinates to 0. Policy: zero offsets and zero step size, so the pass propagates exact zeros.
            P.uv_offset[0] = P.uv_offset[1] = 0.0f;
            P.uvw_offset[0] = P.uvw_offset[1] = P.uvw_offset[2] = 0.0f;
            P.step_size = 0.0f;
            continue;
        }
        double div = (P.face % 2 == 0) ? pa : -pa;
        double r = 1.0 / div;
        D3 q{p.x * r, p.y * r, p.z * r};
        double u = (P.axis == 0 ? q.y : q.x), v = (P.axis == 2 ? q.y : q.z);
        double rz = 1.0 / (double) P.td[2];  // TVector2 /= scalar
        P.uv_offset[0] = (float) (u * rz);
        P.uv_offset[1] = (float) (v * rz);
        // GetStepSizeAndUVWOffset — :132-158
        double r2 = 1.0 / (std::fabs(pa) * (double) P.td[2]);
        D3 off{p.x * r2, p.y * r2, p.z * r2};
        P.step_size = (float) std::sqrt(off.x * off.x + off.y * off.y + off.z * off.z);
        // renormalise to the longest voxel side — LightingShaders.cpp:119-124
        int lowest = std::min(P.td[0], std::min(P.td[1], P.td[2]));
        float longest = 1.0f / lowest;
        normalize_d(off);
        off = D3{off.x * (double) longest, off.y * (double) longest, off.z * (double) longest};
        P.uvw_offset[0] = (float) off.x, P.uvw_offset[1] = (float) off.y, P.uvw_offset[2] = (float) off.z;
    }
    // Add: "break if the axis weight == 0" — LightingShaders.cpp:65-68, 94-97
    out->add_passes = (out->pass[0].weight == 0) ? 0 : ((out->pass[1].weight == 0) ? 1 : 2);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// sweep — per-voxel bodies
// ------------------------------------------------------------------------------------------------------------
namespace {

struct SweepCtx {
    DataTex data;
    LightTex light;
    const float* tf;
    tbrm_windowing win;
    float data_border;
    F3 clip_center, clip_dir;
};

// pos = mul(int3(px,py,Loop), PermutationMatrix) — GetPermutationMatrix, LightingShaderUtils.cpp:227-249 (Q7)
inline void permute(int axis, int px, int py, int loop, int& x, int& y, int& z) {
    if (axis == 0)
        x = loop, y = px, z = py;
    else if (axis == 1)
        x = px, y = loop, z = py;
    else
        x = px, y = py, z = loop;
}

// AddDirLightShader.usf:84-113 — sample position, clip weight, opacity toward the light. `gate_saturate` is the
// all(SampleUVW == saturate(SampleUVW)) test, present in Add (:110) and absent in Change (ChangeDirLightShader.usf:130,136)
inline float occlusion_sample(const SweepCtx& c, int x, int y, int z, const float uvw_off[3], float step_size,
                              bool gate_saturate) {
    const float rx = (float) c.light.X, ry = (float) c.light.Y, rz = (float) c.light.Z;
    // GetUVW(pos, dims) + UVWOffset — RaymarcherCommon.usf:40-43
    F3 S = f3(((float) x + 0.5f) / rx + uvw_off[0], ((float) y + 0.5f) / ry + uvw_off[1], ((float) z + 0.5f) / rz + uvw_off[2]);
    float dist = dot3(f3(S.x - c.clip_center.x, S.y - c.clip_center.y, S.z - c.clip_center.z), c.clip_dir);
    F3 ip = f3(S.x + c.clip_dir.x * dist, S.y + c.clip_dir.y * dist, S.z + c.clip_dir.z * dist);
    F3 off = f3(S.x - ip.x, S.y - ip.y, S.z - ip.z);
    F3 voff = f3(off.x * rx, off.y * ry, off.z * rz);
    float vdist = length3(voff);
    float w = 0.5f + (0.57735026919f * vdist * signf(dist));
    w = fminf(fmaxf(w, 0.0f), 1.0f);
    float cs = 0.0f;
    bool inside = (S.x == saturatef(S.x)) && (S.y == saturatef(S.y)) && (S.z == saturatef(S.z));
    if (w > 0.0f && (!gate_saturate || inside)) {
        float v = sample_data(c.data, S, ADDR_BORDER, c.data_border);
        float rgba[4];
        sample_windowed_tf(v, step_size * 100.0f, c.tf, c.win, rgba);  // StepSize * VOLUME_DENSITY
        cs = rgba[3] * w;
    }
    return cs;
}

}  // namespace

extern "C" int tbo_clear_light_volume(void* light, const int32_t ldims[3], int light_fmt, float value) {
    // ClearVolumeTextureShader.usf:14-20 (the extra i == ZSize slice is an out-of-bounds UAV write, dropped by D3D)
    size_t n = (size_t) ldims[0] * ldims[1] * ldims[2];
    if (light_fmt == TBRM_FMT_G8)
        std::memset(light, quant8(value), n);
    else
        std::fill((float*) light, (float*) light + n, value);
    return 0;
}

static void make_ctx(const tbo_volume* vol, const tbo_light_plan& plan, SweepCtx& c) {
    c.data = DataTex{vol->data, vol->data_fmt, vol->ddims[0], vol->ddims[1], vol->ddims[2]};
    c.light = LightTex{vol->light, vol->light_fmt, vol->ldims[0], vol->ldims[1], vol->ldims[2]};
    c.tf = vol->tf;
    c.win = vol->win;
    c.data_border = plan.data_border;
    c.clip_center = f3(plan.clip_center[0], plan.clip_center[1], plan.clip_center[2]);
    c.clip_dir = f3(plan.clip_dir[0], plan.clip_dir[1], plan.clip_dir[2]);
}

// AddDirLightToSingleLightVolume_RenderThread — LightingShaders.cpp:35-166, kernel AddDirLightShader.usf:69-128
extern "C" int tbo_add_dir_light(const tbo_volume* vol, const tbrm_dir_light* light, int added, const tbrm_world* world,
                                 uint8_t* near_gate) {
    tbo_light_plan plan;
    tbo_plan_dir_light(vol->ldims, &vol->win, vol->border_exact, light, world, &plan);
    if (plan.zero_direction) return 0;
    SweepCtx c;
    make_ctx(vol, plan, c);
    const float sign = added ? 1.0f : -1.0f;  // SetLightAdded, LightingShaders.h:70-74
    int passes = 0;
    for (int i = 0; i < plan.add_passes; ++i) {
        const tbo_pass& P = plan.pass[i];
        const int tx = P.td[0], ty = P.td[1];
        Buf2D buf[2];
        buf[0].init(tx, ty, vol->light_fmt, P.light_alpha);  // :74-79
        buf[1].init(tx, ty, vol->light_fmt, P.light_alpha);
        for (int j = P.start; j != P.stop; j += P.dirn) {  // :132
            const Buf2D& rd = (j % 2 == 0) ? buf[0] : buf[1];  // :149-156
            Buf2D& wr = (j % 2 == 0) ? buf[1] : buf[0];
#pragma omp parallel for schedule(static)
            for (int py = 0; py < ty; ++py)
                for (int px = 0; px < tx; ++px) {
                    int x, y, z;
                    permute(P.axis, px, py, j, x, y, z);
                    // :81-82
                    float u = ((float) px + 0.5f) / (float) tx + P.uv_offset[0];
                    float v = ((float) py + 0.5f) / (float) ty + P.uv_offset[1];
                    float prev = rd.sample_border(u, v, P.border);
                    float cs = occlusion_sample(c, x, y, z, P.uvw_offset, P.step_size, true);
                    float cur = prev * (1.0f - cs);  // :117
                    wr.store(px, py, cur);           // :120
                    if (near_gate && fabsf(fabsf(cur) - 1e-3f) < 2e-6f) near_gate[c.light.idx(x, y, z)] = 1;
                    if (fabsf(cur) > 1e-3f) c.light.store(x, y, z, c.light.load(x, y, z) + (cur * sign));  // :123-127
                }
        }
        ++passes;
    }
    return passes;
}

// NOT in the reference (its Readme.md:165-166, 186-187 names it as the missing optimisation of the paper): the passes of several lights
// that propagate from the SAME cube face joined into one sweep — SURVEY.md §8(f) row 1. The CPU twin of tbrm_add_dir_lights_joined:
//   * every light is planned like AddDirLight (zero directions skipped, passes with weight 0 dropped);
//   * passes are grouped by face in order of first appearance (light order, then pass order), at most 8 per group;
//   * a group is swept once: per slice and voxel the members are evaluated in group order, each with its own propagation buffers and exactly
//     the arithmetic of AddDirLightShader.usf, and each adds to the light volume in turn (through the volume's pixel format).
// A group of one is bit-identical to tbo_add_dir_light's pass. With several members the result differs from consecutive AddDirLight calls
// only in the ORDER in which a voxel's contributions are summed (tests bound it).
extern "C" int tbo_add_dir_lights_joined(const tbo_volume* vol, const tbrm_dir_light* lights, int n_lights, int added, const tbrm_world* world) {
    struct Group {
        int face;
        std::vector<tbo_pass> members;
    };
    std::vector<Group> groups;
    tbo_light_plan first;
    bool have_plan = false;
    for (int i = 0; i < n_lights; ++i) {
        tbo_light_plan plan;
        tbo_plan_dir_light(vol->ldims, &vol->win, vol->border_exact, &lights[i], world, &plan);
        if (plan.zero_direction) continue;
        if (!have_plan) first = plan, have_plan = true;
        for (int p = 0; p < plan.add_passes; ++p) {
            size_t g = 0;
            while (g < groups.size() && !(groups[g].face == plan.pass[p].face && groups[g].members.size() < 8)) ++g;
            if (g == groups.size()) groups.push_back(Group{plan.pass[p].face, {}});
            groups[g].members.push_back(plan.pass[p]);
        }
    }
    if (!have_plan) return 0;
    SweepCtx c;
    make_ctx(vol, first, c);  // clip plane and data border depend on the world and the window only
    const float sign = added ? 1.0f : -1.0f;
    for (const Group& G : groups) {
        const tbo_pass& P0 = G.members[0];
        const int tx = P0.td[0], ty = P0.td[1], K = (int) G.members.size();
        std::vector<Buf2D> buf(2 * K);
        for (int m = 0; m < K; ++m) {
            buf[2 * m].init(tx, ty, vol->light_fmt, G.members[m].light_alpha);
            buf[2 * m + 1].init(tx, ty, vol->light_fmt, G.members[m].light_alpha);
        }
        for (int j = P0.start; j != P0.stop; j += P0.dirn) {
            const int par = (j % 2 == 0) ? 0 : 1;
#pragma omp parallel for schedule(static)
            for (int py = 0; py < ty; ++py)
                for (int px = 0; px < tx; ++px) {
                    int x, y, z;
                    permute(P0.axis, px, py, j, x, y, z);
                    for (int m = 0; m < K; ++m) {
                        const tbo_pass& P = G.members[m];
                        float u = ((float) px + 0.5f) / (float) tx + P.uv_offset[0];
                        float v = ((float) py + 0.5f) / (float) ty + P.uv_offset[1];
                        float prev = buf[2 * m + par].sample_border(u, v, P.border);
                        float cs = occlusion_sample(c, x, y, z, P.uvw_offset, P.step_size, true);
                        float cur = prev * (1.0f - cs);
                        buf[2 * m + (par ^ 1)].store(px, py, cur);
                        if (fabsf(cur) > 1e-3f) c.light.store(x, y, z, c.light.load(x, y, z) + (cur * sign));
                    }
                }
        }
    }
    return (int) groups.size();
}

// ChangeDirLightInSingleLightVolume_RenderThread — LightingShaders.cpp:168-326, kernel ChangeDirLightShader.usf:75-156
extern "C" int tbo_change_dir_light(const tbo_volume* vol, const tbrm_dir_light* old_light, const tbrm_dir_light* new_light,
                                    const tbrm_world* world, uint8_t* near_gate) {
    tbo_light_plan rem, add;
    tbo_plan_dir_light(vol->ldims, &vol->win, vol->border_exact, old_light, world, &rem);
    tbo_plan_dir_light(vol->ldims, &vol->win, vol->border_exact, new_light, world, &add);
    if (rem.zero_direction || add.zero_direction) return 0;  // :173-179
    if (rem.pass[0].face != add.pass[0].face || rem.pass[1].face != add.pass[1].face) {  // :192-198
        int n = tbo_add_dir_light(vol, old_light, 0, world, near_gate);
        n += tbo_add_dir_light(vol, new_light, 1, world, near_gate);
        return 100 + n;
    }
    SweepCtx c;
    make_ctx(vol, rem, c);
    for (int i = 0; i < 2; ++i) {  // always both axes (:203, :238)
        const tbo_pass& R = rem.pass[i];
        const tbo_pass& A = add.pass[i];
        const int tx = R.td[0], ty = R.td[1];
        Buf2D buf[4];
        buf[0].init(tx, ty, vol->light_fmt, R.light_alpha);  // :213-222
        buf[1].init(tx, ty, vol->light_fmt, R.light_alpha);
        buf[2].init(tx, ty, vol->light_fmt, A.light_alpha);
        buf[3].init(tx, ty, vol->light_fmt, A.light_alpha);
        for (int j = R.start; j != R.stop; j += R.dirn) {  // :289
            const bool even = (j % 2 == 0);
            const Buf2D& rrd = even ? buf[0] : buf[1];  // :303-316
            Buf2D& rwr = even ? buf[1] : buf[0];
            const Buf2D& ard = even ? buf[2] : buf[3];
            Buf2D& awr = even ? buf[3] : buf[2];
#pragma omp parallel for schedule(static)
            for (int py = 0; py < ty; ++py)
                for (int px = 0; px < tx; ++px) {
                    int x, y, z;
                    permute(R.axis, px, py, j, x, y, z);
                    float ub = ((float) px + 0.5f) / (float) tx, vb = ((float) py + 0.5f) / (float) ty;
                    float rprev = rrd.sample_border(ub + R.uv_offset[0], vb + R.uv_offset[1], R.border);
                    float aprev = ard.sample_border(ub + A.uv_offset[0], vb + A.uv_offset[1], A.border);
                    float rcs = occlusion_sample(c, x, y, z, R.uvw_offset, R.step_size, false);
                    float acs = occlusion_sample(c, x, y, z, A.uvw_offset, A.step_size, false);
                    float rcur = rprev * (1.0f - rcs);
                    float acur = aprev * (1.0f - acs);
                    rwr.store(px, py, rcur);
                    awr.store(px, py, acur);
                    float diff = acur - rcur;
                    if (near_gate && fabsf(fabsf(diff) - 1e-3f) < 2e-6f) near_gate[c.light.idx(x, y, z)] = 1;
                    if (fabsf(diff) > 1e-3f) c.light.store(x, y, z, c.light.load(x, y, z) + acur - rcur);  // :152-155
                }
        }
    }
    return 2;
}

// ------------------------------------------------------------------------------------------------------------
// raymarch
// ------------------------------------------------------------------------------------------------------------
namespace {

struct CamF {
    F3 eye, fwd, rt, ut;   // rt = right * tan(hfov/2), ut = up * tan(hfov/2) * H/W
    float inv_w2, inv_h2;  // 2/W, 2/H
    float m[4][3];         // WorldToLocal, row-vector convention: local = [w,1] * M
    float depth;
};

// camera basis + WorldToLocal (Q3, Q6): fp64 on the host, rounded once to fp32
void make_cam(const tbrm_camera* cam, const tbrm_world* world, CamF& c) {
    D3 e{cam->eye[0], cam->eye[1], cam->eye[2]};
    D3 f{cam->look_at[0] - e.x, cam->look_at[1] - e.y, cam->look_at[2] - e.z};
    double fl = std::sqrt(f.x * f.x + f.y * f.y + f.z * f.z);
    f = D3{f.x / fl, f.y / fl, f.z / fl};
    D3 r = cross(f, D3{cam->up[0], cam->up[1], cam->up[2]});
    double rl = std::sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
    r = D3{r.x / rl, r.y / rl, r.z / rl};
    D3 u = cross(r, f);
    double tx = std::tan(cam->hfov_deg * 3.14159265358979323846 / 360.0);
    double ty = tx * (double) cam->height / (double) cam->width;
    c.eye = f3((float) e.x, (float) e.y, (float) e.z);
    c.fwd = f3((float) f.x, (float) f.y, (float) f.z);
    c.rt = f3((float) (r.x * tx), (float) (r.y * tx), (float) (r.z * tx));
    c.ut = f3((float) (u.x * ty), (float) (u.y * ty), (float) (u.z * ty));
    c.inv_w2 = 2.0f / (float) cam->width;
    c.inv_h2 = 2.0f / (float) cam->height;
    const double basis[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i) {
        D3 row = inv_transform_vector(*world, D3{basis[i][0], basis[i][1], basis[i][2]});
        c.m[i][0] = (float) row.x, c.m[i][1] = (float) row.y, c.m[i][2] = (float) row.z;
    }
    D3 t = inv_transform_position(*world, D3{0, 0, 0});
    c.m[3][0] = (float) t.x, c.m[3][1] = (float) t.y, c.m[3][2] = (float) t.z;
    c.depth = cam->scene_depth > 0.0f ? cam->scene_depth : 1e8f;
}

inline F3 mul3x3(F3 v, const float m[4][3]) {  // mul(float3, (float3x3) WorldToLocal)
    return f3(((v.x * m[0][0]) + (v.y * m[1][0])) + (v.z * m[2][0]), ((v.x * m[0][1]) + (v.y * m[1][1])) + (v.z * m[2][1]),
              ((v.x * m[0][2]) + (v.y * m[1][2])) + (v.z * m[2][2]));
}
inline F3 mulpos(F3 v, const float m[4][3]) {  // mul(float4(v,1), WorldToLocal).xyz
    F3 a = mul3x3(v, m);
    return f3(a.x + m[3][0], a.y + m[3][1], a.z + m[3][2]);
}

// MaterialParameters.CameraVector for pixel (ix,iy): unit vector from the pixel towards the camera
inline F3 camera_vector(const CamF& c, int ix, int iy) {
    float sx = ((float) ix + 0.5f) * c.inv_w2 - 1.0f;
    float sy = 1.0f - ((float) iy + 0.5f) * c.inv_h2;
    F3 d = f3((c.fwd.x + c.rt.x * sx) + c.ut.x * sy, (c.fwd.y + c.rt.y * sx) + c.ut.y * sy, (c.fwd.z + c.rt.z * sx) + c.ut.z * sy);
    d = normalize3(d);
    return f3(-d.x, -d.y, -d.z);
}

// RayAABBIntersection — RaymarcherCommon.usf:66-88 with BoxMin = 0, BoxMax = 1
inline void ray_aabb(F3 o, F3 r, float& t0, float& t1) {
    F3 inv = f3(1.0f / r.x, 1.0f / r.y, 1.0f / r.z);
    F3 tmin = f3((0.0f - o.x) * inv.x, (0.0f - o.y) * inv.y, (0.0f - o.z) * inv.z);
    F3 tmax = f3((1.0f - o.x) * inv.x, (1.0f - o.y) * inv.y, (1.0f - o.z) * inv.z);
    F3 cl = f3(fminf(tmax.x, tmin.x), fminf(tmax.y, tmin.y), fminf(tmax.z, tmin.z));
    F3 fa = f3(fmaxf(tmax.x, tmin.x), fmaxf(tmax.y, tmin.y), fmaxf(tmax.z, tmin.z));
    t0 = fmaxf(cl.x, fmaxf(cl.y, cl.z));
    t1 = fminf(fa.x, fminf(fa.y, fa.z));
}

// PerformRaymarchCubeSetup — RaymarchMaterialCommon.usf:23-69. Returns entry (UVW) and thickness; also the
// local camera vector (unit) the march recomputes at WindowedRaymarchMaterials.usf:56.
inline void cube_setup(const CamF& c, F3 V, F3& entry, float& thick, F3& lcv) {
    float depth = c.depth;                      // :26
    F3 n = normalize3(V);                       // :32
    F3 wd = f3(n.x * depth, n.y * depth, n.z * depth);
    wd = mul3x3(wd, c.m);                       // :35
    depth = length3(wd);                        // :38
    depth = depth / fabsf(dot3(c.fwd, V));      // :44
    F3 o = mulpos(c.eye, c.m);                  // :47
    F3 mv = normalize3(mul3x3(V, c.m));         // :48
    lcv = f3(-mv.x, -mv.y, -mv.z);
    o = f3(o.x + 0.5f, o.y + 0.5f, o.z + 0.5f);  // :51
    float t0, t1;
    ray_aabb(o, lcv, t0, t1);                   // :54
    t0 = fmaxf(0.0f, t0);                       // :57
    t1 = fminf(depth, t1);                      // :60
    thick = fmaxf(0.0f, t1 - t0);               // :63
    entry = f3(o.x + (t0 * lcv.x), o.y + (t0 * lcv.y), o.z + (t0 * lcv.z));  // :66
}

// Rand3DPCG16 (UE Random.ush, Q4)
inline uint32_t pcg16_x(int px, int py, int pz) {
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

struct MarchCtx {
    DataTex data;
    LightTex light;
    const float* tf;
    tbrm_windowing win;
    AddrMode data_mode;
    F3 clip_center, clip_dir;
};

// AccumulateWindowedRaymarchStep + AccumulateLightEnergy — WindowedRaymarchMaterials.usf:21-33, RaymarchMaterialCommon.usf:82-88
inline void accumulate_step(const MarchCtx& c, F3 p, float step, float acc[4]) {
    float v = sample_data(c.data, p, c.data_mode, 0.0f);
    float s[4];
    sample_windowed_tf(v, step, c.tf, c.win, s);
    float l = sample_light_wrap(c.light, f3(saturatef(p.x), saturatef(p.y), saturatef(p.z)));
    s[0] = s[0] * l, s[1] = s[1] * l, s[2] = s[2] * l;
    float oma = 1.0f - acc[3];
    acc[0] = acc[0] + ((s[0] * s[3]) * oma);
    acc[1] = acc[1] + ((s[1] * s[3]) * oma);
    acc[2] = acc[2] + ((s[2] * s[3]) * oma);
    acc[3] = acc[3] + (s[3] * oma);
}

}  // namespace

extern "C" void tbo_make_camera_uniforms(const tbrm_camera* cam, const tbrm_world* world, tbo_camera_uniforms* out) {
    CamF c;
    make_cam(cam, world, c);
    const F3 v[4] = {c.eye, c.fwd, c.rt, c.ut};
    float* dst[4] = {out->eye, out->fwd, out->rt, out->ut};
    for (int i = 0; i < 4; ++i) dst[i][0] = v[i].x, dst[i][1] = v[i].y, dst[i][2] = v[i].z;
    out->inv_w2 = c.inv_w2, out->inv_h2 = c.inv_h2;
    std::memcpy(out->m, c.m, sizeof(c.m));
    out->depth = c.depth;
    out->width = cam->width, out->height = cam->height, out->frame_mod8 = cam->frame_index % 8, out->jitter = cam->jitter;
}

extern "C" int tbo_raymarch_cube_setup(const tbrm_camera* cam, const tbrm_world* world, float* out) {
    CamF c;
    make_cam(cam, world, c);
#pragma omp parallel for schedule(dynamic, 4)
    for (int iy = 0; iy < cam->height; ++iy)
        for (int ix = 0; ix < cam->width; ++ix) {
            F3 V = camera_vector(c, ix, iy), entry, lcv;
            float thick;
            cube_setup(c, V, entry, thick, lcv);
            float* o = out + 4 * ((size_t) iy * cam->width + ix);
            o[0] = entry.x, o[1] = entry.y, o[2] = entry.z, o[3] = thick;
        }
    return 0;
}

// PerformWindowedLitRaymarch — WindowedRaymarchMaterials.usf:36-96
extern "C" int tbo_raymarch_lit(const tbo_volume* vol, const tbrm_camera* cam, const tbrm_world* world, float step_count,
                                int row_begin, int row_end, float* out_rgba, uint64_t* out_steps, uint8_t* near_gate) {
    CamF c;
    make_cam(cam, world, c);
    tbo_light_plan plan;
    tbrm_dir_light dummy{{0, 0, 1}, 0.0f};
    tbo_plan_dir_light(vol->ldims, &vol->win, vol->border_exact, &dummy, world, &plan);  // local clip params (RaymarchVolume.cpp:705-728)
    MarchCtx m;
    m.data = DataTex{vol->data, vol->data_fmt, vol->ddims[0], vol->ddims[1], vol->ddims[2]};
    m.light = LightTex{vol->light, vol->light_fmt, vol->ldims[0], vol->ldims[1], vol->ldims[2]};
    m.tf = vol->tf;
    m.win = vol->win;
    m.data_mode = vol->data_addr_wrap ? ADDR_WRAP : ADDR_CLAMP;
    m.clip_center = f3(plan.clip_center[0], plan.clip_center[1], plan.clip_center[2]);
    m.clip_dir = f3(plan.clip_dir[0], plan.clip_dir[1], plan.clip_dir[2]);
    const int W = cam->width;
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : total)
    for (int iy = row_begin; iy < row_end; ++iy)
        for (int ix = 0; ix < W; ++ix) {
            F3 V = camera_vector(c, ix, iy), cur, lcv;
            float thick;
            cube_setup(c, V, cur, thick, lcv);
            float ss = 1 / step_count;                    // :47
            float fas = step_count * thick;               // :49
            float fl = floorf(fas);
            int max_steps = (int) fl;                     // :51
            float fin = fas - fl;                         // :53 frac()
            F3 sv = f3(lcv.x * ss, lcv.y * ss, lcv.z * ss);  // :56
            float ssw = 100.0f * ss;                      // :58
            float acc[4] = {0, 0, 0, 0};                  // :60
            if (cam->jitter) {                            // :62, RaymarchMaterialCommon.usf:73-78
                float rnd = (float) pcg16_x(ix, iy, cam->frame_index % 8) / 65535.0f;
                cur = f3(cur.x - sv.x * rnd, cur.y - sv.y * rnd, cur.z - sv.z * rnd);
            }
            uint64_t steps = 0;
            bool flagged = false;
            int i = 0;
            for (i = 0; i < max_steps; i++) {             // :65
                cur = f3(cur.x + sv.x, cur.y + sv.y, cur.z + sv.z);  // :67
                ++steps;
                // IsCurPosClipped — RaymarcherCommon.usf:22-25
                float cd = dot3(f3(cur.x - m.clip_center.x, cur.y - m.clip_center.y, cur.z - m.clip_center.z), m.clip_dir);
                if (!(cd <= 0.0f)) {
                    accumulate_step(m, cur, ssw, acc);    // :71
                    if (fabsf(acc[3] - 0.95f) < 2e-6f) flagged = true;
                    if (acc[3] > 0.95f) {                 // :75-79
                        acc[3] = 1.0f;
                        break;
                    }
                }
            }
            if (i == max_steps && fin > 0.0f) {           // :84
                cur = f3(cur.x + sv.x * fin, cur.y + sv.y * fin, cur.z + sv.z * fin);  // :86
                ++steps;
                float cd = dot3(f3(cur.x - m.clip_center.x, cur.y - m.clip_center.y, cur.z - m.clip_center.z), m.clip_dir);
                if (!(cd <= 0.0f)) accumulate_step(m, cur, 100.0f * fin, acc);  // :90-91 (VOLUME_DENSITY * FinalStep)
            }
            size_t o = 4 * ((size_t) (iy - row_begin) * W + ix);
            out_rgba[o] = acc[0], out_rgba[o + 1] = acc[1], out_rgba[o + 2] = acc[2], out_rgba[o + 3] = acc[3];
            if (near_gate) near_gate[(size_t) (iy - row_begin) * W + ix] = flagged ? 1 : 0;
            total += steps;
        }
    if (out_steps) *out_steps = total;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// the other two materials and the octree they read — SURVEY.md §8(f) row 2
// ------------------------------------------------------------------------------------------------------------
namespace {
// per-pixel preamble shared by the three marches (WindowedRaymarchMaterials.usf:46-62, 113-130, 196-208)
struct MarchSetup {
    F3 cur, sv;
    int max_steps;
    float fin, ss;
};
inline MarchSetup march_setup(const CamF& c, const tbrm_camera* cam, float step_count, int ix, int iy) {
    F3 V = camera_vector(c, ix, iy), cur, lcv;
    float thick;
    cube_setup(c, V, cur, thick, lcv);
    MarchSetup m;
    m.ss = 1 / step_count;
    float fas = step_count * thick;
    float fl = floorf(fas);
    m.max_steps = (int) fl;
    m.fin = fas - fl;
    m.sv = f3(lcv.x * m.ss, lcv.y * m.ss, lcv.z * m.ss);
    if (cam->jitter) {
        float rnd = (float) pcg16_x(ix, iy, cam->frame_index % 8) / 65535.0f;
        cur = f3(cur.x - m.sv.x * rnd, cur.y - m.sv.y * rnd, cur.z - m.sv.z * rnd);
    }
    m.cur = cur;
    return m;
}
inline bool clipped(F3 p, F3 cc, F3 cd) {  // IsCurPosClipped — RaymarcherCommon.usf:22-25
    return dot3(f3(p.x - cc.x, p.y - cc.y, p.z - cc.z), cd) <= 0.0f;
}
inline void local_clip(const tbo_volume* vol, const tbrm_world* world, F3& cc, F3& cd) {
    tbo_light_plan plan;
    tbrm_dir_light dummy{{0, 0, 1}, 0.0f};
    const int32_t one[3] = {1, 1, 1};
    const tbrm_windowing w0{0.5f, 1.0f, 1, 1};
    tbo_plan_dir_light(vol ? vol->ldims : one, vol ? &vol->win : &w0, 1, &dummy, world, &plan);
    cc = f3(plan.clip_center[0], plan.clip_center[1], plan.clip_center[2]);
    cd = f3(plan.clip_dir[0], plan.clip_dir[1], plan.clip_dir[2]);
}
// UNORM16 UAV store (A.1)
inline uint16_t quant16(float v) { return (uint16_t) floorf(saturatef(v) * 65535.0f + 0.5f); }
struct OctreeTex {
    const uint16_t* mip[4];
    int X[4], Y[4], Z[4];
    inline float load(int m, int x, int y, int z) const {  // Texture3D.Load: out-of-bounds -> 0
        if (x < 0 || y < 0 || z < 0 || x >= X[m] || y >= Y[m] || z >= Z[m]) return 0.0f;
        return (float) mip[m][(size_t) x + (size_t) X[m] * ((size_t) y + (size_t) Y[m] * (size_t) z)] / 65535.0f;
    }
};
inline void octree_dims(const int32_t odims[3], OctreeTex& t) {
    for (int m = 0; m < 4; ++m) t.X[m] = std::max(1, odims[0] >> m), t.Y[m] = std::max(1, odims[1] >> m), t.Z[m] = std::max(1, odims[2] >> m);
}
}  // namespace

// PerformWindowedIntensityRaymarch — WindowedRaymarchMaterials.usf:187-242: the first unclipped sample, windowed to grey, alpha 1
extern "C" int tbo_raymarch_intensity(const tbo_volume* vol, const tbrm_camera* cam, const tbrm_world* world, float step_count,
                                      int row_begin, int row_end, float* out_rgba, uint64_t* out_steps) {
    CamF c;
    make_cam(cam, world, c);
    F3 cc, cd;
    local_clip(vol, world, cc, cd);
    DataTex data{vol->data, vol->data_fmt, vol->ddims[0], vol->ddims[1], vol->ddims[2]};
    const int W = cam->width;
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : total)
    for (int iy = row_begin; iy < row_end; ++iy)
        for (int ix = 0; ix < W; ++ix) {
            MarchSetup m = march_setup(c, cam, step_count, ix, iy);
            F3 cur = m.cur;
            float o[4] = {0, 0, 0, 0};  // :241 "didn't hit anything"
            uint64_t steps = 0;
            bool hit = false;
            for (int i = 0; i < m.max_steps; i++) {  // :211
                cur = f3(cur.x + m.sv.x, cur.y + m.sv.y, cur.z + m.sv.z);
                ++steps;
                F3 sp = f3(saturatef(cur.x), saturatef(cur.y), saturatef(cur.z));
                if (!clipped(sp, cc, cd)) {           // :215 tests the SATURATED position
                    float v = sample_data(data, sp, ADDR_CLAMP, 0.0f);  // :217
                    float pos = saturatef(tf_position(v, vol->win.center, vol->win.width));  // :220 clamp(..., 0, 1)
                    o[0] = o[1] = o[2] = pos, o[3] = 1.0f;
                    hit = true;
                    break;                            // :222 return
                }
            }
            if (!hit && m.fin > 0.0f) {               // :227
                cur = f3(cur.x + m.sv.x * m.fin, cur.y + m.sv.y * m.fin, cur.z + m.sv.z * m.fin);
                ++steps;
                if (!clipped(cur, cc, cd)) {          // :231 tests the raw position
                    float v = sample_data(data, cur, ADDR_CLAMP, 0.0f);
                    float pos = saturatef(tf_position(v, vol->win.center, vol->win.width));
                    o[0] = o[1] = o[2] = pos, o[3] = 1.0f;
                }
            }
            size_t p = 4 * ((size_t) (iy - row_begin) * W + ix);
            out_rgba[p] = o[0], out_rgba[p + 1] = o[1], out_rgba[p + 2] = o[2], out_rgba[p + 3] = o[3];
            total += steps;
        }
    if (out_steps) *out_steps = total;
    return 0;
}

// GenerateOctreeShader.usf:28-107 dispatched by GenerateOctreeForVolume_RenderThread (OctreeShaders.cpp:28-54): one thread per 8^3 leaf
// of a render target whose sides are the data volume's rounded up to powers of two (RaymarchVolume.cpp:873-877), 4 mips, PF_G16.
// mips[m]: (odims >> m, at least 1) UNORM16 texels. Mip 0 = data value * MinMaxValues.y (= 1), 0 outside the data volume; mip m =
// max over 2x2x2 texels of mip m-1 (read back through the UNORM16 UAV, i.e. after quantisation).
extern "C" int tbo_generate_octree(const void* data, const int32_t ddims[3], int data_fmt, const int32_t odims[3], void* const* mips) {
    DataTex vol{data, data_fmt, ddims[0], ddims[1], ddims[2]};
    OctreeTex t;
    octree_dims(odims, t);
    uint16_t* w[4];
    for (int m = 0; m < 4; ++m) w[m] = (uint16_t*) mips[m], t.mip[m] = w[m];
    auto store = [&](int m, int x, int y, int z, float v) {  // out-of-bounds UAV stores are dropped
        if (x < t.X[m] && y < t.Y[m] && z < t.Z[m]) w[m][(size_t) x + (size_t) t.X[m] * ((size_t) y + (size_t) t.Y[m] * (size_t) z)] = quant16(v);
    };
    const int gx = (odims[0] + 7) / 8, gy = (odims[1] + 7) / 8, gz = (odims[2] + 7) / 8;
#pragma omp parallel for schedule(static) collapse(2)
    for (int lz = 0; lz < gz; ++lz)
        for (int ly = 0; ly < gy; ++ly)
            for (int lx = 0; lx < gx; ++lx) {
                const int ox = lx * 8, oy = ly * 8, oz = lz * 8;  // ThreadOffset :33
                for (int x = 0; x < 8; x++)                       // :36-49
                    for (int y = 0; y < 8; y++)
                        for (int z = 0; z < 8; z++) {
                            const int ax = ox + x, ay = oy + y, az = oz + z;
                            const bool in = ax < vol.X && ay < vol.Y && az < vol.Z;
                            store(0, ax, ay, az, (in ? vol.texel(ax, ay, az) : 0.0f) * 1.0f);
                        }
                for (int mip = 1; mip < 4; mip++) {               // :60-105
                    const int div = 1 << mip;
                    const int lox = (2 * ox) / div, loy = (2 * oy) / div, loz = (2 * oz) / div;  // LowerMipOffset :72
                    for (int x = 0; x < 8 / div; x++)
                        for (int y = 0; y < 8 / div; y++)
                            for (int z = 0; z < 8 / div; z++) {
                                float mx = 0;
                                for (int a = 0; a < 2; a++)
                                    for (int b = 0; b < 2; b++)
                                        for (int cc = 0; cc < 2; cc++) {
                                            float nv = t.load(mip - 1, lox + x * 2 + a, loy + y * 2 + b, loz + z * 2 + cc);
                                            if (mx < nv) mx = nv;
                                        }
                                store(mip, ox / div + x, oy / div + y, oz / div + z, mx);
                            }
                }
            }
    return 0;
}

// PerformWindowedRaymarchOctree — WindowedRaymarchMaterials.usf:99-183: the lit march's loop with a point Load from one octree mip
// instead of the trilinear data sample, and no light volume.
extern "C" int tbo_raymarch_octree(const tbo_volume* vol, const tbrm_camera* cam, const tbrm_world* world, float step_count, int row_begin,
                                   int row_end, const void* const* mips, const int32_t odims[3], int octree_mip, float* out_rgba,
                                   uint64_t* out_steps) {
    CamF c;
    make_cam(cam, world, c);
    F3 cc, cd;
    local_clip(vol, world, cc, cd);
    OctreeTex t;
    octree_dims(odims, t);
    for (int m = 0; m < 4; ++m) t.mip[m] = (const uint16_t*) mips[m];
    const int mip = octree_mip;
    const bool mip_ok = mip >= 0 && mip < 4;
    const int mq = mip_ok ? mip : 3;  // GetDimensions of a missing mip: nothing to march (Load returns 0 everywhere)
    const float ow = (float) t.X[mq], oh = (float) t.Y[mq], od = (float) t.Z[mq];  // :135
    const float od0 = (float) t.Z[0];                                              // OctreeDepthConst :132
    const float dd = (float) vol->ddims[2];                                        // DataVolumeDepth :126
    const int W = cam->width;
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : total)
    for (int iy = row_begin; iy < row_end; ++iy)
        for (int ix = 0; ix < W; ++ix) {
            MarchSetup m = march_setup(c, cam, step_count, ix, iy);
            F3 cur = m.cur;
            const float ssw = 100.0f * m.ss;  // :119
            float acc[4] = {0, 0, 0, 0};
            uint64_t steps = 0;
            auto sample = [&](F3 p) {
                // :148 int3 VoxelPos = float3(...): truncation towards zero
                int vx = (int) (p.x * ow), vy = (int) (p.y * oh), vz = (int) (((p.z * dd) / od0) * od);
                float v = mip_ok ? t.load(mip, vx, vy, vz) : 0.0f;  // SampleWindowedVolumeOctreeStep, WindowedSampling.usf:47-52
                float s[4];
                sample_windowed_tf(v, ssw, vol->tf, vol->win, s);
                float oma = 1.0f - acc[3];  // AccumulateLightEnergy
                acc[0] = acc[0] + ((s[0] * s[3]) * oma);
                acc[1] = acc[1] + ((s[1] * s[3]) * oma);
                acc[2] = acc[2] + ((s[2] * s[3]) * oma);
                acc[3] = acc[3] + (s[3] * oma);
            };
            int i = 0;
            for (i = 0; i < m.max_steps; i++) {  // :138
                cur = f3(cur.x + m.sv.x, cur.y + m.sv.y, cur.z + m.sv.z);
                ++steps;
                if (!clipped(cur, cc, cd)) {
                    sample(cur);
                    if (acc[3] > 0.95f) {  // :156-160
                        acc[3] = 1.0f;
                        break;
                    }
                }
            }
            if (i == m.max_steps && m.fin > 0.0f) {  // :165 — note: the opacity step stays StepSizeWorld here (:173), unlike the lit march
                cur = f3(cur.x + m.sv.x * m.fin, cur.y + m.sv.y * m.fin, cur.z + m.sv.z * m.fin);
                ++steps;
                if (!clipped(cur, cc, cd)) sample(cur);
            }
            size_t p = 4 * ((size_t) (iy - row_begin) * W + ix);
            out_rgba[p] = acc[0], out_rgba[p + 1] = acc[1], out_rgba[p + 2] = acc[2], out_rgba[p + 3] = acc[3];
            total += steps;
        }
    if (out_steps) *out_steps = total;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Mandelbulb — SDFMarcher.usf
// ------------------------------------------------------------------------------------------------------------
namespace {
int g_mandelbulb_variant = 0;  // 0: the reference's formulation; 1: the transcendental-free power-8 twin of csrc/mandelbulb.cu

// Mandelbulb_SDF — SDFMarcher.usf:24-52
inline float mandelbulb_sdf_reference(F3 pos, float bailout, float power, int iterations, uint64_t& iters) {
    F3 z = pos;
    float dr = 1.0f, r = 0.0f;
    for (int i = 0; i < iterations; i++) {
        r = length3(z);
        if (r > bailout) break;
        ++iters;
        float theta = acosf(z.z / r);
        float phi = atan2f(z.y, z.x);
        dr = powf(r, power - 1.0f) * power * dr + 1.0f;
        float zr = powf(r, power);
        theta = theta * power;
        phi = phi * power;
        z = f3(zr * (sinf(theta) * cosf(phi)), zr * (sinf(phi) * sinf(theta)), zr * cosf(theta));
        z = f3(z.x + pos.x, z.y + pos.y, z.z + pos.z);
    }
    return 0.5f * logf(r) * r / dr;
}

// NOT the reference's code: the CPU twin of the kernels' Power == 8 fast path (mandelbulb_sdf_p8 in csrc/mandelbulb.cu). The same
// iteration with cos(theta) = z * (1/r), sin(theta) = rho * (1/r), cos(phi) = x * (1/rho), sin(phi) = y * (1/rho), three angle doublings and r^8 by squaring:
// only +, -, *, /, sqrt — the same bits on the CPU and on the GPU up to the final log. Selected by tbo_set_mandelbulb_variant(1); the
// tests use it to check the kernels tightly, and variant 0 to bound how far the fast path strays from the reference's formulation.
inline float mandelbulb_sdf_p8(F3 pos, float bailout, int iterations, uint64_t& iters) {
    float zx = pos.x, zy = pos.y, zz = pos.z;
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
        for (int d = 0; d < 3; ++d) {
            const float c2 = (ct * ct) - (st * st), s2 = 2.0f * (ct * st);
            ct = c2, st = s2;
            const float c3 = (cp * cp) - (sp * sp), s3 = 2.0f * (cp * sp);
            cp = c3, sp = s3;
        }
        zx = r8 * (st * cp) + pos.x;
        zy = r8 * (sp * st) + pos.y;
        zz = r8 * ct + pos.z;
    }
    return 0.5f * logf(r) * r / dr;
}
inline float mandelbulb_sdf(F3 pos, float bailout, float power, int iterations, uint64_t& iters) {
    if (g_mandelbulb_variant == 1 && power == 8.0f) return mandelbulb_sdf_p8(pos, bailout, iterations, iters);
    return mandelbulb_sdf_reference(pos, bailout, power, iterations, iters);
}
}  // namespace

// PerformMandelbulbRaymarchReturnDistance — SDFMarcher.usf:61-112 (entry from PerformRaymarchCubeSetup)
extern "C" int tbo_mandelbulb_march(const tbrm_mandelbulb* mb, const tbrm_camera* cam, const tbrm_world* world, int row_begin,
                                    int row_end, float* out_xy, uint64_t* out_iterations) {
    CamF c;
    make_cam(cam, world, c);
    const int W = cam->width;
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : total)
    for (int iy = row_begin; iy < row_end; ++iy)
        for (int ix = 0; ix < W; ++ix) {
            F3 V = camera_vector(c, ix, iy), cur, lcv;
            float thick;
            cube_setup(c, V, cur, thick, lcv);
            float ox = 0.0f, oy = 0.0f;
            uint64_t iters = 0;
            if (thick > 0.0f) {  // the cube mesh only rasterises pixels the ray actually crosses
                F3 step = f3(lcv.x / mb->extent, lcv.y / mb->extent, lcv.z / mb->extent);  // :76
                float dist = 0.0f;
                bool done = false;
                for (int s = 0; (float) s < mb->max_steps; s++) {  // :80
                    F3 ap = f3(mb->center[0] + ((cur.x - 0.5f) * mb->extent), mb->center[1] + ((cur.y - 0.5f) * mb->extent),
                               mb->center[2] + ((cur.z - 0.5f) * mb->extent));  // GetActualPosition :54-58
                    dist = mandelbulb_sdf(ap, mb->bailout, mb->power, (int) mb->max_iterations, iters);
                    if (dist < mb->high_precision_eps) {  // :85-90
                        float ratio = (float) s / (float) mb->max_steps;
                        ratio = ratio * 10.0f;
                        ox = 1.0f - ratio, oy = 1.0f;
                        done = true;
                        break;
                    }
                    cur = f3(cur.x + (dist * step.x), cur.y + (dist * step.y), cur.z + (dist * step.z));  // :93
                    if (saturatef(cur.x) != cur.x || saturatef(cur.y) != cur.y || saturatef(cur.z) != cur.z) {  // :96-99
                        ox = 0.0f, oy = 0.0f;
                        done = true;
                        break;
                    }
                }
                if (!done && dist < mb->low_precision_eps) ox = 0.0f, oy = 1.0f;  // :104-108
            }
            size_t o = 2 * ((size_t) (iy - row_begin) * W + ix);
            out_xy[o] = ox, out_xy[o + 1] = oy;
            total += iters;
        }
    if (out_iterations) *out_iterations = total;
    return 0;
}

// PerformMandelbulbRaymarchReturnNormal — SDFMarcher.usf:117-188: the same sphere tracing; a hit returns the normalised vector of three
// SDF evaluations at positions offset BACKWARDS by DerivationDistance / Extent along each axis (:156-165 — not a true gradient), alpha 1.
extern "C" int tbo_mandelbulb_march_normal(const tbrm_mandelbulb* mb, float derivation_distance, const tbrm_camera* cam, const tbrm_world* world,
                                           int row_begin, int row_end, float* out_rgba, uint64_t* out_iterations) {
    CamF c;
    make_cam(cam, world, c);
    const int W = cam->width;
    uint64_t total = 0;
    const float dd = derivation_distance / mb->extent;  // :138
    auto actual = [&](F3 p) {                           // GetActualPosition :54-58
        return f3(mb->center[0] + ((p.x - 0.5f) * mb->extent), mb->center[1] + ((p.y - 0.5f) * mb->extent), mb->center[2] + ((p.z - 0.5f) * mb->extent));
    };
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : total)
    for (int iy = row_begin; iy < row_end; ++iy)
        for (int ix = 0; ix < W; ++ix) {
            F3 V = camera_vector(c, ix, iy), cur, lcv;
            float thick;
            cube_setup(c, V, cur, thick, lcv);
            float o[4] = {0, 0, 0, 0};
            uint64_t iters = 0;
            if (thick > 0.0f) {
                F3 step = f3(lcv.x / mb->extent, lcv.y / mb->extent, lcv.z / mb->extent);  // :134
                float dist = 0.0f;
                bool done = false;
                for (int s = 0; (float) s < mb->max_steps; s++) {  // :142
                    dist = mandelbulb_sdf(actual(cur), mb->bailout, mb->power, (int) mb->max_iterations, iters);
                    if (dist < mb->high_precision_eps) {  // :147
                        F3 n;
                        n.x = mandelbulb_sdf(actual(f3(cur.x - dd, cur.y - 0.0f, cur.z - 0.0f)), mb->bailout, mb->power, (int) mb->max_iterations, iters);
                        n.y = mandelbulb_sdf(actual(f3(cur.x - 0.0f, cur.y - dd, cur.z - 0.0f)), mb->bailout, mb->power, (int) mb->max_iterations, iters);
                        n.z = mandelbulb_sdf(actual(f3(cur.x - 0.0f, cur.y - 0.0f, cur.z - dd)), mb->bailout, mb->power, (int) mb->max_iterations, iters);
                        n = normalize3(n);  // :166
                        o[0] = n.x, o[1] = n.y, o[2] = n.z, o[3] = 1.0f;
                        done = true;
                        break;
                    }
                    cur = f3(cur.x + (dist * step.x), cur.y + (dist * step.y), cur.z + (dist * step.z));  // :171
                    if (saturatef(cur.x) != cur.x || saturatef(cur.y) != cur.y || saturatef(cur.z) != cur.z) {  // :174-177
                        done = true;
                        break;
                    }
                }
                if (!done && dist < mb->low_precision_eps) o[3] = 1.0f;  // :182-186 "return black normal"
            }
            size_t p = 4 * ((size_t) (iy - row_begin) * W + ix);
            out_rgba[p] = o[0], out_rgba[p + 1] = o[1], out_rgba[p + 2] = o[2], out_rgba[p + 3] = o[3];
            total += iters;
        }
    if (out_iterations) *out_iterations = total;
    return 0;
}

// CalculateMandelbulbSDF.usf:24-65 dispatched by CalculateMandelbulbSDF_RenderThread (FractalShaders.cpp:41-70): per voxel of the volume
// texture (PF_G16 in the reference, FractalVolume.cpp:166) the distance estimate with 50 iterations and Bailout = Extent, divided by Extent.
// out_fmt: TBRM_FMT_G16 (UNORM16 store, as the reference) or TBRM_FMT_R32F (the unquantised value).
extern "C" int tbo_mandelbulb_sdf(const int32_t dims[3], const float center[3], float extent, float power, int out_fmt, void* out,
                                  uint64_t* out_iterations) {
    if (!(extent > 0.0f)) return 0;  // EnqueueRenderCommand_CalculateMandelbulbSDF :28-31
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int z = 0; z < dims[2]; ++z)
        for (int y = 0; y < dims[1]; ++y)
            for (int x = 0; x < dims[0]; ++x) {
                F3 uvw = f3((float) x / (float) dims[0], (float) y / (float) dims[1], (float) z / (float) dims[2]);  // :58 (no +0.5)
                F3 n = f3(uvw.x - 0.5f, uvw.y - 0.5f, uvw.z - 0.5f);
                F3 co = f3(center[0] + (n.x * extent), center[1] + (n.y * extent), center[2] + (n.z * extent));
                uint64_t iters = 0;
                float v = mandelbulb_sdf(co, extent, power, 50, iters) / extent;  // :26-27, :63
                size_t i = (size_t) x + (size_t) dims[0] * ((size_t) y + (size_t) dims[1] * (size_t) z);
                if (out_fmt == TBRM_FMT_G16)
                    ((uint16_t*) out)[i] = quant16(v);
                else
                    ((float*) out)[i] = v;
                total += iters;
            }
    if (out_iterations) *out_iterations = total;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// volume ingest — SURVEY.md §8(f) row 3
// ------------------------------------------------------------------------------------------------------------
namespace {
// UVolumeTextureToolkit::ConvertArrayToNormalizedArray<InType, OutType> — TextureUtilities.h:103-149. Notes that follow the reference:
// InMax starts at numeric_limits<InType>::min(), which for float is the smallest POSITIVE normal (:111); the float -> OutType
// conversion truncates (:142); a constant volume divides 0 by 0 — the NaN converts to 0 on x86 (policy; undefined in C++).
template <typename In, typename Out>
void normalize_array(const In* in, size_t n, Out* out, float& omin, float& omax) {
    In mn = std::numeric_limits<In>::max(), mx = std::numeric_limits<In>::min();
    for (size_t i = 0; i < n; i++) {
        if (in[i] < mn) mn = in[i];
        if (in[i] > mx) mx = in[i];
    }
    const float fmn = (float) mn, fmx = (float) mx;
    const float range = fmx - fmn;
    const float omaxf = (float) std::numeric_limits<Out>::max();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        float nrm = ((float) in[i] - fmn) / range;
        float v = 0.0f + (nrm * omaxf);
        out[i] = (v >= 0.0f && v < omaxf + 1.0f) ? (Out) v : (Out) 0;
    }
    omin = fmn, omax = fmx;
}
template <typename In>
void to_float_array(const In* in, size_t n, float* out) {  // ConvertArrayToFloatTemplated — TextureUtilities.h:153-178
    for (size_t i = 0; i < n; i++) out[i] = static_cast<float>(in[i]);
}
}  // namespace

// UVolumeTextureToolkit::NormalizeArrayByFormat — TextureUtilities.cpp:304-327. fmt = EVolumeVoxelFormat (VolumeInfo.h:12-27):
// 0 u8, 1 i8, 2 u16, 3 i16, 4 u32, 5 i32, 6 f32. 1-byte inputs normalise to u8, everything else to u16. Returns bytes per output voxel.
extern "C" int tbo_normalize_array(int fmt, const void* in, uint64_t count, void* out, float* out_min, float* out_max) {
    switch (fmt) {
        case 0: normalize_array((const uint8_t*) in, count, (uint8_t*) out, *out_min, *out_max); return 1;
        case 1: normalize_array((const int8_t*) in, count, (uint8_t*) out, *out_min, *out_max); return 1;
        case 2: normalize_array((const uint16_t*) in, count, (uint16_t*) out, *out_min, *out_max); return 2;
        case 3: normalize_array((const int16_t*) in, count, (uint16_t*) out, *out_min, *out_max); return 2;
        case 4: normalize_array((const uint32_t*) in, count, (uint16_t*) out, *out_min, *out_max); return 2;
        case 5: normalize_array((const int32_t*) in, count, (uint16_t*) out, *out_min, *out_max); return 2;
        case 6: normalize_array((const float*) in, count, (uint16_t*) out, *out_min, *out_max); return 2;
        default: return 0;
    }
}
// UVolumeTextureToolkit::ConvertArrayToFloat — TextureUtilities.cpp:329-350 (a float input is not converted: returns 1)
extern "C" int tbo_convert_to_float(int fmt, const void* in, uint64_t count, float* out) {
    switch (fmt) {
        case 0: to_float_array((const uint8_t*) in, count, out); return 0;
        case 1: to_float_array((const int8_t*) in, count, out); return 0;
        case 2: to_float_array((const uint16_t*) in, count, out); return 0;
        case 3: to_float_array((const int16_t*) in, count, out); return 0;
        case 4: to_float_array((const uint32_t*) in, count, out); return 0;
        case 5: to_float_array((const int32_t*) in, count, out); return 0;
        default: return 1;
    }
}

// ------------------------------------------------------------------------------------------------------------
// synthetic inputs of SURVEY.md §8(d) (fp64; bit-identical to tbraymarcherplugin_b200/synth.py and csrc/synth.cu)
// ------------------------------------------------------------------------------------------------------------
namespace {
inline uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
inline double lattice(int x, int y, int z, uint32_t seed) {
    return (double) lowbias32((uint32_t) x + 374761393u * (uint32_t) y + 668265263u * (uint32_t) z + seed) / 4294967296.0;
}
}  // namespace

extern "C" int tbo_synth_volume_u8(int kind, const int32_t dims[3], uint32_t seed, uint8_t* out) {
    const int X = dims[0], Y = dims[1], Z = dims[2];
#pragma omp parallel for schedule(static)
    for (int z = 0; z < Z; ++z)
        for (int y = 0; y < Y; ++y)
            for (int x = 0; x < X; ++x) {
                const double u = ((double) x + 0.5) / (double) X, v = ((double) y + 0.5) / (double) Y, w = ((double) z + 0.5) / (double) Z;
                double val;
                if (kind == 0) {
                    const double du = u - 0.5, dv = v - 0.5, dw = w - 0.5;
                    val = std::fmax(0.0, 1.0 - std::sqrt(du * du + dv * dv + dw * dw) / 0.4);
                } else {
                    double sum = 0.0, amp = 1.0, norm = 0.0;
                    for (int o = 0; o < 4; ++o) {
                        const double cells = (double) (4 << o);
                        const double px = u * cells, py = v * cells, pz = w * cells;
                        const double fx0 = std::floor(px), fy0 = std::floor(py), fz0 = std::floor(pz);
                        const int ix = (int) fx0, iy = (int) fy0, iz = (int) fz0;
                        double fx = px - fx0, fy = py - fy0, fz = pz - fz0;
                        fx = fx * fx * (3.0 - 2.0 * fx);
                        fy = fy * fy * (3.0 - 2.0 * fy);
                        fz = fz * fz * (3.0 - 2.0 * fz);
                        const uint32_t s = seed + (uint32_t) o * 0x9E3779B9u;
                        const double c000 = lattice(ix, iy, iz, s), c100 = lattice(ix + 1, iy, iz, s);
                        const double c010 = lattice(ix, iy + 1, iz, s), c110 = lattice(ix + 1, iy + 1, iz, s);
                        const double c001 = lattice(ix, iy, iz + 1, s), c101 = lattice(ix + 1, iy, iz + 1, s);
                        const double c011 = lattice(ix, iy + 1, iz + 1, s), c111 = lattice(ix + 1, iy + 1, iz + 1, s);
                        const double x00 = c000 + fx * (c100 - c000), x10 = c010 + fx * (c110 - c010);
                        const double x01 = c001 + fx * (c101 - c001), x11 = c011 + fx * (c111 - c011);
                        const double y0 = x00 + fy * (x10 - x00), y1 = x01 + fy * (x11 - x01);
                        sum = sum + amp * (y0 + fz * (y1 - y0));
                        norm = norm + amp;
                        amp = amp * 0.5;
                    }
                    const double noise = sum / norm;
                    const double eu = (u - 0.5) / 0.45, ev = (v - 0.5) / 0.40, ew = (w - 0.5) / 0.48;
                    const double e = std::sqrt(eu * eu + ev * ev + ew * ew);
                    val = noise * std::fmin(1.0, std::fmax(0.0, (1.0 - e) / 0.02));
                }
                out[(size_t) x + (size_t) X * ((size_t) y + (size_t) Y * z)] = (uint8_t) std::floor(255.0 * val + 0.5);
            }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// exported helpers for the known-answer tests
// ------------------------------------------------------------------------------------------------------------
extern "C" void tbo_set_mandelbulb_variant(int v) { g_mandelbulb_variant = v; }
// Mandelbulb_SDF at one position: variant 0 = the reference's formulation, 1 = the power-8 twin (whatever the global switch says)
extern "C" float tbo_mandelbulb_sdf_at(const float pos[3], float bailout, float power, int iterations, int variant, uint64_t* out_iterations) {
    uint64_t it = 0;
    const F3 p = f3(pos[0], pos[1], pos[2]);
    const float d = (variant == 1 && power == 8.0f) ? mandelbulb_sdf_p8(p, bailout, iterations, it) : mandelbulb_sdf_reference(p, bailout, power, iterations, it);
    if (out_iterations) *out_iterations = it;
    return d;
}
extern "C" float tbo_det_pow(float x, float y) { return det_pow(x, y); }
extern "C" float tbo_round_to_half(float x) { return round_to_half(x); }
extern "C" void tbo_sample_windowed_tf(float value, float step, const float* tf, const tbrm_windowing* w, float out[4]) {
    sample_windowed_tf(value, step, tf, *w, out);
}
extern "C" float tbo_sample_data(const void* data, const int32_t dims[3], int fmt, float u, float v, float w, int mode, float border) {
    DataTex t{data, fmt, dims[0], dims[1], dims[2]};
    return sample_data(t, f3(u, v, w), (AddrMode) mode, border);
}
extern "C" uint32_t tbo_pcg16_x(int x, int y, int z) { return pcg16_x(x, y, z); }
extern "C" int tbo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
extern "C" void tbo_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void) n;
#endif
}
