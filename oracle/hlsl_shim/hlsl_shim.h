// hlsl_shim.h — SHIM (test infrastructure), ours: the slice of HLSL and of Unreal's shader environment that the reference's
// shaders use, as C++ types, so that oracle/ref.mk can run the reference's OWN shader code on the CPU. oracle/hlsl2cpp.py streams
// the .usf files from /root/reference into oracle/_ref/gen/ (a build intermediate, never committed) with purely syntactic
// rewrites — see its header — and oracle/ref_shaders.cpp includes the result inside namespace hlsl.
//
// What this pins and what it does not. Control flow, expression structure, gating conditions, thresholds and the order of
// operations inside the shaders are the reference's. Everything below is OUR statement of what the GPU / engine does with them,
// and deliberately the SAME statement as the oracle's arithmetic contract:
//   * every float operation is one correctly rounded fp32 op (the .so is built with -ffp-contract=off);
//   * dot(a,b) = ((ax*bx)+(ay*by))+(az*bz); length = sqrt(dot); normalize = v / length; mul(v, M) sums in the same order;
//   * lerp(a,b,t) = fma(t, b-a, a) and the samplers are lerp trees in x, then y, then z, taps at floor(u*N - 0.5);
//   * texture formats: UNORM8 / UNORM16 loads are v/255, v/65535; UNORM stores round(saturate(v)*max); out-of-bounds UAV
//     writes are dropped, out-of-bounds loads return 0; border / clamp / wrap addressing per sampler;
//   * pow / transcendentals are chosen per shader by ref_shaders.cpp (det_pow for the opacity correction, libm for the Mandelbulb).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace hlsl {

typedef unsigned int uint;

struct float2;
struct float3;
struct float4;
struct int3;
struct uint3;

struct uint2 {
    uint x, y;
    uint2() : x(0), y(0) {}
    uint2(uint x_, uint y_) : x(x_), y(y_) {}
};
struct float2 {
    float x, y;
    float2() : x(0), y(0) {}
    float2(float s) : x(s), y(s) {}
    float2(float x_, float y_) : x(x_), y(y_) {}
    float2(const uint2& u) : x((float) u.x), y((float) u.y) {}
};
inline float2 operator+(float2 a, float2 b) { return float2(a.x + b.x, a.y + b.y); }
inline float2 operator/(float2 a, float2 b) { return float2(a.x / b.x, a.y / b.y); }

struct bool3 {
    bool x, y, z;
};
inline bool all(bool3 b) { return b.x && b.y && b.z; }
inline bool any(bool3 b) { return b.x || b.y || b.z; }

struct float3 {
    float x, y, z;
    float3() : x(0), y(0), z(0) {}
    float3(float s) : x(s), y(s), z(s) {}
    float3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float3(const int3& i);
    float3(const uint3& u);
};
inline float3 operator+(float3 a, float3 b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(float3 a, float3 b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator*(float3 a, float3 b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator/(float3 a, float3 b) { return float3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline float3 operator*(float3 a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float s, float3 a) { return float3(s * a.x, s * a.y, s * a.z); }
inline float3 operator/(float3 a, float s) { return float3(a.x / s, a.y / s, a.z / s); }
inline float3 operator/(float s, float3 a) { return float3(s / a.x, s / a.y, s / a.z); }
inline float3 operator+(float3 a, float s) { return float3(a.x + s, a.y + s, a.z + s); }
inline float3 operator-(float3 a, float s) { return float3(a.x - s, a.y - s, a.z - s); }
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
inline float3& operator+=(float3& a, float3 b) { return a = a + b; }
inline float3& operator-=(float3& a, float3 b) { return a = a - b; }
inline float3& operator+=(float3& a, float s) { return a = a + s; }
inline bool3 operator==(float3 a, float3 b) { return bool3{a.x == b.x, a.y == b.y, a.z == b.z}; }
inline bool3 operator!=(float3 a, float3 b) { return bool3{a.x != b.x, a.y != b.y, a.z != b.z}; }

struct int3 {
    int x, y, z;
    int3() : x(0), y(0), z(0) {}
    int3(int x_, int y_, int z_) : x(x_), y(y_), z(z_) {}
    int3(const float3& f) : x((int) f.x), y((int) f.y), z((int) f.z) {}  // ftoi: truncation towards zero
    int3(const uint3& u);
    int3(const float2& f, int z_) : x((int) f.x), y((int) f.y), z(z_) {}
};
inline int3 operator+(int3 a, int3 b) { return int3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline int3 operator*(int3 a, int s) { return int3(a.x * s, a.y * s, a.z * s); }
inline int3 operator*(int s, int3 a) { return int3(s * a.x, s * a.y, s * a.z); }
inline int3 operator/(int3 a, int s) { return int3(a.x / s, a.y / s, a.z / s); }
struct uint3 {
    uint x, y, z;
    uint3() : x(0), y(0), z(0) {}
    uint3(uint x_, uint y_, uint z_) : x(x_), y(y_), z(z_) {}
    uint3(const int3& i) : x((uint) i.x), y((uint) i.y), z((uint) i.z) {}
};
inline int3::int3(const uint3& u) : x((int) u.x), y((int) u.y), z((int) u.z) {}
inline float3::float3(const int3& i) : x((float) i.x), y((float) i.y), z((float) i.z) {}
inline float3::float3(const uint3& u) : x((float) u.x), y((float) u.y), z((float) u.z) {}
struct int4 {
    int x, y, z, w;
    int4(int x_, int y_, int z_, int w_) : x(x_), y(y_), z(z_), w(w_) {}
    int4(const int3& p, int w_) : x(p.x), y(p.y), z(p.z), w(w_) {}
};

struct float4 {
    union { float x, r; };
    union { float y, g; };
    union { float z, b; };
    union { float w, a; };
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float s) : x(s), y(s), z(s), w(s) {}
    float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    float4(const float3& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    float3 xyz() const { return float3(x, y, z); }
    float3 rgb() const { return float3(x, y, z); }
    float2 xy() const { return float2(x, y); }
    void set_rgb(const float3& v) { x = v.x, y = v.y, z = v.z; }
};

struct float3x3 {
    float m[3][3];
};
struct float4x4 {
    float m[4][4];
};

// ---- intrinsics (the arithmetic contract) ----------------------------------------------------------------------
inline float lerp(float a, float b, float t) { return fmaf(t, b - a, a); }
inline float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
inline float3 saturate(float3 v) { return float3(saturate(v.x), saturate(v.y), saturate(v.z)); }
inline float clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline float min(int a, float b) { return fminf((float) a, b); }
inline float max(int a, float b) { return fmaxf((float) a, b); }
inline float min(float a, int b) { return fminf(a, (float) b); }
inline float max(float a, int b) { return fmaxf(a, (float) b); }
inline float3 min(float3 a, float3 b) { return float3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
inline float3 max(float3 a, float3 b) { return float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
inline float abs(float x) { return fabsf(x); }
inline float floor(float x) { return floorf(x); }
inline float round(float x) { return nearbyintf(x); }
inline float frac(float x) { return x - floorf(x); }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float sqrt(float x) { return sqrtf(x); }
inline float dot(float3 a, float3 b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); }
inline float length(float3 v) { return sqrtf(dot(v, v)); }
inline float distance(float3 a, float3 b) { return length(a - b); }
inline float3 normalize(float3 v) {
    const float l = length(v);
    return float3(v.x / l, v.y / l, v.z / l);
}
// row vector times matrix
inline float3 mul(float3 v, const float3x3& M) {
    return float3(((v.x * M.m[0][0]) + (v.y * M.m[1][0])) + (v.z * M.m[2][0]), ((v.x * M.m[0][1]) + (v.y * M.m[1][1])) + (v.z * M.m[2][1]),
                  ((v.x * M.m[0][2]) + (v.y * M.m[1][2])) + (v.z * M.m[2][2]));
}
inline float3 mul(int3 v, const float3x3& M) { return mul(float3(v), M); }
// float3 x float4x4: HLSL truncates the matrix to its upper 3x3
inline float3 mul(float3 v, const float4x4& M) {
    return float3(((v.x * M.m[0][0]) + (v.y * M.m[1][0])) + (v.z * M.m[2][0]), ((v.x * M.m[0][1]) + (v.y * M.m[1][1])) + (v.z * M.m[2][1]),
                  ((v.x * M.m[0][2]) + (v.y * M.m[1][2])) + (v.z * M.m[2][2]));
}
inline float4 mul(float4 v, const float4x4& M) {
    float o[4];
    for (int j = 0; j < 4; ++j) o[j] = (((v.x * M.m[0][j]) + (v.y * M.m[1][j])) + (v.z * M.m[2][j])) + (v.w * M.m[3][j]);
    return float4(o[0], o[1], o[2], o[3]);
}

// ---- resources ---------------------------------------------------------------------------------------------------
enum Fmt { FMT_UNORM8 = 0, FMT_UNORM16 = 1, FMT_R32F = 2, FMT_RGBA32F = 3 };
enum Addr { ADDR_CLAMP = 0, ADDR_WRAP = 1, ADDR_BORDER = 2 };

struct SamplerState {
    int addr = ADDR_CLAMP;
    float border = 0.0f;
};

struct Storage {
    void* p = nullptr;
    int fmt = FMT_R32F;
    int X = 1, Y = 1, Z = 1;
    inline bool inside(int x, int y, int z) const { return x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z; }
    inline size_t idx(int x, int y, int z) const { return (size_t) x + (size_t) X * ((size_t) y + (size_t) Y * (size_t) z); }
    inline float load(int x, int y, int z, int c = 0) const {
        const size_t i = idx(x, y, z);
        switch (fmt) {
            case FMT_UNORM8: return (float) ((const uint8_t*) p)[i] / 255.0f;
            case FMT_UNORM16: return (float) ((const uint16_t*) p)[i] / 65535.0f;
            case FMT_RGBA32F: return ((const float*) p)[4 * i + c];
            default: return ((const float*) p)[i];
        }
    }
    inline void store(int x, int y, int z, float v) const {
        if (!inside(x, y, z)) return;  // out-of-bounds UAV writes are dropped
        const size_t i = idx(x, y, z);
        switch (fmt) {
            case FMT_UNORM8: ((uint8_t*) p)[i] = (uint8_t) floorf(saturate(v) * 255.0f + 0.5f); break;
            case FMT_UNORM16: ((uint16_t*) p)[i] = (uint16_t) floorf(saturate(v) * 65535.0f + 0.5f); break;
            default: ((float*) p)[i] = v;
        }
    }
};

inline void axis_taps(float u, int N, int& i0, float& f) {
    float x = u * (float) N - 0.5f;
    float fl = floorf(x);
    f = x - fl;
    fl = fminf(fmaxf(fl, -4.0f), (float) N + 4.0f);
    i0 = (int) fl;
}
inline int address(int i, int N, int mode) {
    if (mode == ADDR_CLAMP) return i < 0 ? 0 : (i >= N ? N - 1 : i);
    if (mode == ADDR_WRAP) {
        int r = i % N;
        return r < 0 ? r + N : r;
    }
    return i;
}
inline float fetch(const Storage& s, const SamplerState& smp, int x, int y, int z, int c) {
    if (smp.addr == ADDR_BORDER) return s.inside(x, y, z) ? s.load(x, y, z, c) : (c == 0 ? smp.border : 0.0f);
    return s.load(address(x, s.X, smp.addr), address(y, s.Y, smp.addr), address(z, s.Z, smp.addr), c);
}

struct Texture3D {
    Storage s;            // mip 0
    const Storage* mips = nullptr;  // optional mip chain (mips[0] == s)
    int nmips = 1;
    void GetDimensions(float& w, float& h, float& d) const { w = (float) s.X, h = (float) s.Y, d = (float) s.Z; }
    void GetDimensions(int& w, int& h, int& d) const { w = s.X, h = s.Y, d = s.Z; }
    void GetDimensions(uint mip, float& w, float& h, float& d, float& n) const {
        const Storage& m = mips ? mips[mip < (uint) nmips ? mip : nmips - 1] : s;
        w = (float) m.X, h = (float) m.Y, d = (float) m.Z, n = (float) nmips;
    }
    float4 SampleLevel(const SamplerState& smp, float3 uvw, float) const {
        int i0, j0, k0;
        float fx, fy, fz;
        axis_taps(uvw.x, s.X, i0, fx), axis_taps(uvw.y, s.Y, j0, fy), axis_taps(uvw.z, s.Z, k0, fz);
        const float c00 = lerp(fetch(s, smp, i0, j0, k0, 0), fetch(s, smp, i0 + 1, j0, k0, 0), fx);
        const float c01 = lerp(fetch(s, smp, i0, j0 + 1, k0, 0), fetch(s, smp, i0 + 1, j0 + 1, k0, 0), fx);
        const float c10 = lerp(fetch(s, smp, i0, j0, k0 + 1, 0), fetch(s, smp, i0 + 1, j0, k0 + 1, 0), fx);
        const float c11 = lerp(fetch(s, smp, i0, j0 + 1, k0 + 1, 0), fetch(s, smp, i0 + 1, j0 + 1, k0 + 1, 0), fx);
        return float4(lerp(lerp(c00, c01, fy), lerp(c10, c11, fy), fz), 0.0f, 0.0f, 1.0f);
    }
    float4 Load(int4 p, int = 0) const {  // p.w = mip level; out-of-bounds loads return 0
        const Storage& m = mips ? mips[p.w >= 0 && p.w < nmips ? p.w : 0] : s;
        if ((mips && (p.w < 0 || p.w >= nmips)) || !m.inside(p.x, p.y, p.z)) return float4(0.0f);
        return float4(m.load(p.x, p.y, p.z), 0.0f, 0.0f, 1.0f);
    }
};

struct Texture2D {
    Storage s;
    float4 SampleLevel(const SamplerState& smp, float2 uv, float) const {
        int i0, j0;
        float fx, fy;
        axis_taps(uv.x, s.X, i0, fx), axis_taps(uv.y, s.Y, j0, fy);
        const int nc = s.fmt == FMT_RGBA32F ? 4 : 1;
        float o[4] = {0.0f, 0.0f, 0.0f, 1.0f};
        for (int c = 0; c < nc; ++c) {
            const float r0 = lerp(fetch(s, smp, i0, j0, 0, c), fetch(s, smp, i0 + 1, j0, 0, c), fx);
            const float r1 = lerp(fetch(s, smp, i0, j0 + 1, 0, c), fetch(s, smp, i0 + 1, j0 + 1, 0, c), fx);
            o[c] = lerp(r0, r1, fy);
        }
        return float4(o[0], o[1], o[2], o[3]);
    }
};

struct Texel {  // the l-value / r-value of UAV[pos]
    const Storage* s;
    int x, y, z;
    operator float() const { return s->inside(x, y, z) ? s->load(x, y, z) : 0.0f; }
    Texel& operator=(float v) {
        s->store(x, y, z, v);
        return *this;
    }
    Texel& operator=(const Texel& o) { return *this = (float) o; }
};
template <typename T>
struct RWTexture3D {
    Storage s;
    void GetDimensions(uint& w, uint& h, uint& d) const { w = (uint) s.X, h = (uint) s.Y, d = (uint) s.Z; }
    Texel operator[](const int3& p) const { return Texel{&s, p.x, p.y, p.z}; }
};
template <typename T>
struct RWTexture2D {
    Storage s;
    void GetDimensions(float& w, float& h) const { w = (float) s.X, h = (float) s.Y; }
    Texel operator[](const uint2& p) const { return Texel{&s, (int) p.x, (int) p.y, 0}; }
};

// ---- Unreal's material / view environment (stand-ins fed by ref_shaders.cpp from the same camera model the oracle uses) --------
struct FPrimitiveData {
    float4x4 WorldToLocal;
};
struct FMaterialPixelParameters {
    float3 CameraVector;  // unit vector from the pixel towards the camera
    int PrimitiveId = 0;
    float4 SvPosition;    // pixel centre
    float SceneDepth = 1e8f;
};
struct FViewState {
    float3 WorldCameraOrigin;
    float4x4 ViewToTranslatedWorld;  // row 2 = camera forward
    uint StateFrameIndexMod8 = 0;
};
struct FMaterialSamplers {
    SamplerState Clamp_WorldGroupSettings{ADDR_CLAMP, 0.0f};
    SamplerState Wrap_WorldGroupSettings{ADDR_WRAP, 0.0f};
};
extern FPrimitiveData g_primitive;  // constant over a frame: set by ref_shaders.cpp before the pixel loop
extern FViewState ResolvedView;
#define View ResolvedView
extern FMaterialSamplers Material;
inline const FPrimitiveData& GetPrimitiveData(int) { return g_primitive; }
inline const float4x4& LWCHackToFloat(const float4x4& m) { return m; }
inline float3 LWCHackToFloat(const float3& v) { return v; }
inline const FMaterialPixelParameters& GetScreenPosition(const FMaterialPixelParameters& p) { return p; }
inline const FMaterialPixelParameters& ScreenAlignedPosition(const FMaterialPixelParameters& p) { return p; }
inline float CalcSceneDepth(const FMaterialPixelParameters& p) { return p.SceneDepth; }
// Rand3DPCG16 (UE Random.ush)
inline uint3 Rand3DPCG16(int3 p) {
    uint x = (uint) p.x, y = (uint) p.y, z = (uint) p.z;
    x = x * 1664525u + 1013904223u;
    y = y * 1664525u + 1013904223u;
    z = z * 1664525u + 1013904223u;
    x += y * z;
    y += z * x;
    z += x * y;
    x += y * z;
    y += z * x;
    z += x * y;
    return uint3(x >> 16, y >> 16, z >> 16);
}

}  // namespace hlsl
