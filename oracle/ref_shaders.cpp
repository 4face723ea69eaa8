// ref_shaders.cpp — runs the REFERENCE'S OWN SHADER CODE on the CPU. TEST INFRASTRUCTURE ONLY, part of oracle/_ref/libtbrm_ref.so.
//
// The .inc files included below are the reference's .usf sources (AddDirLightShader, ChangeDirLightShader,
// WindowedRaymarchMaterials [+ RaymarchMaterialCommon, WindowedSampling, RaymarcherCommon], GenerateOctreeShader, SDFMarcher,
// CalculateMandelbulbSDF), streamed at build time from /root/reference by oracle/hlsl2cpp.py with syntactic rewrites only (see
// its header) into oracle/_ref/gen/ — a build intermediate that is never committed. They compile against oracle/hlsl_shim (OUR
// statement of HLSL types, intrinsics, samplers and UE's material environment, sharing the oracle's arithmetic contract).
//
// What is ours in this file: the dispatch drivers, which bind the uniforms the reference's host functions compute
// (tbref_plan_dir_light in ref_wrap.cpp = LightingShaderUtils.cpp compiled from the reference) and loop over slices / pixels the way
// the render-thread drivers that cannot be compiled here do (Source/Raymarcher/Private/Rendering/LightingShaders.cpp:35-326, RHI code;
// OctreeShaders.cpp:28-54; Source/FractalMarcher/Private/Rendering/FractalShaders.cpp:41-70), and the camera stand-in for UE's view.
//
// What a match between this and tbrm_oracle.cpp proves: the oracle restates the shaders' logic faithfully (gating, thresholds,
// operation order, addressing). What it cannot prove: that the shim's engine semantics equal D3D11's — those stay the policies of
// SURVEY.md Appendix B.
#include <algorithm>
#include <cmath>
#include <vector>

#include "hlsl_shim/hlsl_shim.h"
#include "tbrm_contract.h"

#include "../include/tbrm.h"
#include "tbrm_oracle.h"

extern "C" int tbref_plan_dir_light(const int32_t ldims[3], const tbrm_dir_light* light, const tbrm_world* world, tbo_light_plan* out);
extern "C" void tbref_permutation_rows(int face, double rows[9]);
extern "C" void tbref_local_clipping(const tbrm_world* world, float center[3], float dir[3]);
extern "C" float tbref_data_border(const tbrm_windowing* win, int exact);

namespace hlsl {

FPrimitiveData g_primitive;
FViewState ResolvedView;
FMaterialSamplers Material;

#define TBREF_DET_POW \
    inline float pow(float a, float b) { return tbrm_contract::det_pow(a, b); }
#define TBREF_LIBM                                              \
    inline float pow(float a, float b) { return powf(a, b); }   \
    inline float acos(float x) { return acosf(x); }             \
    inline float atan2(float y, float x) { return atan2f(y, x); } \
    inline float sin(float x) { return sinf(x); }               \
    inline float cos(float x) { return cosf(x); }               \
    inline float log(float x) { return logf(x); }

namespace add_dir_light {
TBREF_DET_POW
#include "_ref/gen/AddDirLightShader.inc"
}  // namespace add_dir_light
namespace change_dir_light {
TBREF_DET_POW
#include "_ref/gen/ChangeDirLightShader.inc"
}  // namespace change_dir_light
namespace materials {
TBREF_DET_POW
#include "_ref/gen/WindowedRaymarchMaterials.inc"
}  // namespace materials
namespace generate_octree {
#include "_ref/gen/GenerateOctreeShader.inc"
}  // namespace generate_octree
namespace sdf_marcher {
TBREF_LIBM
#include "_ref/gen/SDFMarcher.inc"
}  // namespace sdf_marcher
namespace mandelbulb_sdf {
TBREF_LIBM
#include "_ref/gen/CalculateMandelbulbSDF.inc"
}  // namespace mandelbulb_sdf

}  // namespace hlsl

namespace {
using namespace hlsl;

int to_fmt(int tbrm_fmt) { return tbrm_fmt == TBRM_FMT_G8 ? FMT_UNORM8 : (tbrm_fmt == TBRM_FMT_G16 ? FMT_UNORM16 : FMT_R32F); }
Storage data_storage(const tbo_volume* v) { return Storage{const_cast<void*>(v->data), to_fmt(v->data_fmt), v->ddims[0], v->ddims[1], v->ddims[2]}; }
Storage light_storage(const tbo_volume* v) { return Storage{v->light, to_fmt(v->light_fmt), v->ldims[0], v->ldims[1], v->ldims[2]}; }
Storage tf_storage(const tbo_volume* v) { return Storage{const_cast<float*>(v->tf), FMT_RGBA32F, 256, 1, 1}; }
float4 windowing(const tbo_volume* v) {  // FWindowingParameters::ToLinearColor, VolumeInfo.h:49-52
    return float4(v->win.center, v->win.width, v->win.low_cutoff ? 1.0f : 0.0f, v->win.high_cutoff ? 1.0f : 0.0f);
}
float3x3 permutation(int face) {
    double rows[9];
    tbref_permutation_rows(face, rows);  // GetPermutationMatrix of the reference
    float3x3 m;
    for (int i = 0; i < 9; ++i) m.m[i / 3][i % 3] = (float) rows[i];
    return m;
}
inline float finite_or_zero(float v) { return std::isfinite(v) ? v : 0.0f; }  // policy Q11

// a propagation buffer in the light volume's pixel format ("Illumination Buffer", RaymarchUtils.cpp:176-196)
struct Buffer2D {
    std::vector<uint8_t> bytes;
    Storage s;
    Buffer2D(int w, int h, int fmt, float clear) : bytes((size_t) w * h * (fmt == FMT_UNORM8 ? 1 : 4)) {
        s = Storage{bytes.data(), fmt, w, h, 1};
        for (int y = 0; y < h; ++y)  // Clear2DTexture_RenderThread: a UAV store of the value per texel
            for (int x = 0; x < w; ++x) s.store(x, y, 0, clear);
    }
};

template <typename F>
void dispatch_2d(int w, int h, F&& thread) {  // groups of 16 x 16 threads, DivideAndRoundUp: the padding threads run too
    const int gw = (w + 15) / 16 * 16, gh = (h + 15) / 16 * 16;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < gh; ++y)
        for (int x = 0; x < gw; ++x) thread(uint2((uint) x, (uint) y));
}

int run_add_dir_light(const tbo_volume* vol, const tbrm_dir_light* light, int added, const tbrm_world* world) {
    namespace S = hlsl::add_dir_light;
    tbo_light_plan plan;
    tbref_plan_dir_light(vol->ldims, light, world, &plan);
    if (plan.zero_direction) return 0;  // LightingShaders.cpp:41-46
    S::Volume.s = data_storage(vol);
    S::VolumeSampler = SamplerState{ADDR_BORDER, tbref_data_border(&vol->win, vol->border_exact)};  // LightingShaders.h:76-94
    S::TransferFunc.s = tf_storage(vol);
    S::TransferFuncSampler = SamplerState{ADDR_CLAMP, 0.0f};
    S::ALightVolume.s = light_storage(vol);
    S::LocalClippingCenter = float3(plan.clip_center[0], plan.clip_center[1], plan.clip_center[2]);
    S::LocalClippingDirection = float3(plan.clip_dir[0], plan.clip_dir[1], plan.clip_dir[2]);
    S::WindowingParameters = windowing(vol);
    S::bAdded = added ? 1 : -1;
    const int lfmt = to_fmt(vol->light_fmt);
    for (unsigned i = 0; i < 2; i++) {
        const tbo_pass& P = plan.pass[i];
        if (P.weight == 0) break;  // :94-97
        Buffer2D b0(P.td[0], P.td[1], lfmt, P.light_alpha), b1(P.td[0], P.td[1], lfmt, P.light_alpha);  // :78-81
        S::ReadBufferSampler = SamplerState{ADDR_BORDER, vol->border_exact ? P.light_alpha : P.border};  // :102-103
        S::PrevPixelOffset = float2(P.uv_offset[0], P.uv_offset[1]);
        S::UVWOffset = float3(P.uvw_offset[0], P.uvw_offset[1], P.uvw_offset[2]);
        S::PermutationMatrix = permutation(P.face);
        S::StepSize = P.step_size;
        for (int j = P.start; j != P.stop; j += P.dirn) {  // :132-158
            S::Loop = j;
            S::ReadBuffer.s = (j % 2 == 0) ? b0.s : b1.s;
            S::WriteBuffer.s = (j % 2 == 0) ? b1.s : b0.s;
            dispatch_2d(P.td[0], P.td[1], [](uint2 p) { S::MainComputeShader(p); });
        }
    }
    return 0;
}

int run_change_dir_light(const tbo_volume* vol, const tbrm_dir_light* old_light, const tbrm_dir_light* new_light, const tbrm_world* world) {
    namespace S = hlsl::change_dir_light;
    tbo_light_plan rem, add;
    tbref_plan_dir_light(vol->ldims, old_light, world, &rem);
    tbref_plan_dir_light(vol->ldims, new_light, world, &add);
    if (rem.zero_direction || add.zero_direction) return 0;  // LightingShaders.cpp:173-179
    if (rem.pass[0].face != add.pass[0].face || rem.pass[1].face != add.pass[1].face) {  // :192-198
        run_add_dir_light(vol, old_light, 0, world);
        run_add_dir_light(vol, new_light, 1, world);
        return 1;
    }
    S::Volume.s = data_storage(vol);
    S::VolumeSampler = SamplerState{ADDR_BORDER, tbref_data_border(&vol->win, vol->border_exact)};
    S::TransferFunc.s = tf_storage(vol);
    S::TransferFuncSampler = SamplerState{ADDR_CLAMP, 0.0f};
    S::ALightVolume.s = light_storage(vol);
    S::LocalClippingCenter = float3(rem.clip_center[0], rem.clip_center[1], rem.clip_center[2]);
    S::LocalClippingDirection = float3(rem.clip_dir[0], rem.clip_dir[1], rem.clip_dir[2]);
    S::WindowingParameters = windowing(vol);
    const int lfmt = to_fmt(vol->light_fmt);
    for (unsigned a = 0; a < 2; a++) {  // both axes, whatever the weights (:242)
        const tbo_pass &R = rem.pass[a], &A = add.pass[a];
        Buffer2D r0(R.td[0], R.td[1], lfmt, R.light_alpha), r1(R.td[0], R.td[1], lfmt, R.light_alpha);  // UAVs[0], [1]
        Buffer2D a0(R.td[0], R.td[1], lfmt, A.light_alpha), a1(R.td[0], R.td[1], lfmt, A.light_alpha);  // UAVs[2], [3]
        S::RemovedReadBufferSampler = SamplerState{ADDR_BORDER, vol->border_exact ? R.light_alpha : R.border};
        S::ReadBufferSampler = SamplerState{ADDR_BORDER, vol->border_exact ? A.light_alpha : A.border};
        S::PrevPixelOffset = float2(finite_or_zero(A.uv_offset[0]), finite_or_zero(A.uv_offset[1]));
        S::RemovedPrevPixelOffset = float2(finite_or_zero(R.uv_offset[0]), finite_or_zero(R.uv_offset[1]));
        S::UVWOffset = float3(finite_or_zero(A.uvw_offset[0]), finite_or_zero(A.uvw_offset[1]), finite_or_zero(A.uvw_offset[2]));
        S::RemovedUVWOffset = float3(finite_or_zero(R.uvw_offset[0]), finite_or_zero(R.uvw_offset[1]), finite_or_zero(R.uvw_offset[2]));
        S::StepSize = finite_or_zero(A.step_size);
        S::RemovedStepSize = finite_or_zero(R.step_size);
        S::PermutationMatrix = permutation(R.face);
        for (int j = R.start; j != R.stop; j += R.dirn) {  // :289-318
            S::Loop = j;
            const bool even = j % 2 == 0;
            S::RemovedReadBuffer.s = even ? r0.s : r1.s, S::RemovedWriteBuffer.s = even ? r1.s : r0.s;
            S::ReadBuffer.s = even ? a0.s : a1.s, S::WriteBuffer.s = even ? a1.s : a0.s;
            dispatch_2d(R.td[0], R.td[1], [](uint2 p) { S::MainComputeShader(p); });
        }
    }
    return 0;
}

// UE's view for pixel (ix, iy): the stand-in pinhole camera of include/tbrm.h (uniforms computed by the caller)
void bind_view(const tbo_camera_uniforms* c) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) g_primitive.WorldToLocal.m[i][j] = j < 3 ? c->m[i][j] : (i == 3 ? 1.0f : 0.0f);
    ResolvedView.WorldCameraOrigin = float3(c->eye[0], c->eye[1], c->eye[2]);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) ResolvedView.ViewToTranslatedWorld.m[i][j] = 0.0f;
    for (int j = 0; j < 3; ++j) ResolvedView.ViewToTranslatedWorld.m[2][j] = c->fwd[j];
    ResolvedView.StateFrameIndexMod8 = (uint) c->frame_mod8;
}
FMaterialPixelParameters pixel_parameters(const tbo_camera_uniforms* c, int ix, int iy) {
    const float sx = ((float) ix + 0.5f) * c->inv_w2 - 1.0f;
    const float sy = 1.0f - ((float) iy + 0.5f) * c->inv_h2;
    float3 d((c->fwd[0] + c->rt[0] * sx) + c->ut[0] * sy, (c->fwd[1] + c->rt[1] * sx) + c->ut[1] * sy, (c->fwd[2] + c->rt[2] * sx) + c->ut[2] * sy);
    d = normalize(d);
    FMaterialPixelParameters mp;
    mp.CameraVector = -d;
    mp.SvPosition = float4((float) ix + 0.5f, (float) iy + 0.5f, 0.0f, 1.0f);
    mp.SceneDepth = c->depth;
    return mp;
}
}  // namespace

extern "C" int tbref_add_dir_light(const tbo_volume* vol, const tbrm_dir_light* light, int added, const tbrm_world* world) {
    return run_add_dir_light(vol, light, added, world);
}
extern "C" int tbref_change_dir_light(const tbo_volume* vol, const tbrm_dir_light* old_light, const tbrm_dir_light* new_light,
                                      const tbrm_world* world) {
    return run_change_dir_light(vol, old_light, new_light, world);
}

// material: -1 PerformRaymarchCubeSetup only, 0 PerformWindowedLitRaymarch, 1 PerformWindowedIntensityRaymarch,
// 2 PerformWindowedRaymarchOctree (octree_mips: 4 UNORM16 mip pointers, odims: mip-0 dimensions). out: float4 per pixel of the rows.
extern "C" int tbref_raymarch(int material, const tbo_volume* vol, const tbo_camera_uniforms* cam, const tbrm_world* world, float step_count,
                              int row_begin, int row_end, const void* const* octree_mips, const int32_t* odims, int octree_mip, float* out) {
    namespace M = hlsl::materials;
    bind_view(cam);
    float cc[3], cd[3];
    tbref_local_clipping(world, cc, cd);  // GetLocalClippingParameters of the reference (RaymarchVolume.cpp:705-728 feeds the material)
    const float3 clip_c(cc[0], cc[1], cc[2]), clip_d(cd[0], cd[1], cd[2]);
    Texture3D data, light, octree;
    Texture2D tf;
    Storage mips[4];
    if (vol) {
        data.s = data_storage(vol);
        light.s = light_storage(vol);
        tf.s = tf_storage(vol);
    }
    if (material == 2) {
        for (int m = 0; m < 4; ++m)
            mips[m] = Storage{const_cast<void*>(octree_mips[m]), FMT_UNORM16, std::max(1, odims[0] >> m), std::max(1, odims[1] >> m), std::max(1, odims[2] >> m)};
        octree.s = mips[0], octree.mips = mips, octree.nmips = 4;
    }
    const SamplerState data_sampler{vol && vol->data_addr_wrap ? ADDR_WRAP : ADDR_CLAMP, 0.0f};
    const float4 win = vol ? windowing(vol) : float4(0.0f);
    const int W = cam->width;
#pragma omp parallel for schedule(dynamic, 2)
    for (int iy = row_begin; iy < row_end; ++iy)
        for (int ix = 0; ix < W; ++ix) {
            const FMaterialPixelParameters mp = pixel_parameters(cam, ix, iy);
            const float4 setup = M::PerformRaymarchCubeSetup(mp);
            float4 r = setup;
            if (material == 0)
                r = M::PerformWindowedLitRaymarch(data, data_sampler, tf, light, setup.xyz(), setup.w, step_count, clip_c, clip_d, win, mp);
            else if (material == 1)
                r = M::PerformWindowedIntensityRaymarch(data, setup.xyz(), setup.w, step_count, clip_c, clip_d, win, mp);
            else if (material == 2)
                r = M::PerformWindowedRaymarchOctree(data, data_sampler, tf, setup.xyz(), setup.w, step_count, clip_c, clip_d, win, octree,
                                                     Material.Clamp_WorldGroupSettings, (uint) octree_mip, mp);
            float* o = out + 4 * ((size_t) (iy - row_begin) * W + ix);
            o[0] = r.x, o[1] = r.y, o[2] = r.z, o[3] = r.w;
        }
    return 0;
}

// GenerateOctreeForVolume_RenderThread (OctreeShaders.cpp:28-54): one thread per 8^3 leaf of the pow-2 sized UNORM16 render target
extern "C" int tbref_generate_octree(const void* data, const int32_t ddims[3], int data_fmt, const int32_t odims[3], void* const* mips) {
    namespace S = hlsl::generate_octree;
    S::Volume.s = Storage{const_cast<void*>(data), to_fmt(data_fmt), ddims[0], ddims[1], ddims[2]};
    RWTexture3D<float>* uav[4] = {&S::OctreeVolumeMip0, &S::OctreeVolumeMip1, &S::OctreeVolumeMip2, &S::OctreeVolumeMip3};
    for (int m = 0; m < 4; ++m)
        uav[m]->s = Storage{mips[m], FMT_UNORM16, std::max(1, odims[0] >> m), std::max(1, odims[1] >> m), std::max(1, odims[2] >> m)};
    S::MinMaxValues = float2(0.0f, 1.0f);  // OctreeShaders.h: SetShaderValue(..., MinMaxValues, FVector2f(0.0, 1.0))
    S::LeafNodeSize = 8;                   // LEAF_NODE_SIZE
    S::NumberOfMips = 4;                   // GetNumMips()
    const int gx = (odims[0] + 7) / 8, gy = (odims[1] + 7) / 8, gz = (odims[2] + 7) / 8;
#pragma omp parallel for schedule(static) collapse(2)
    for (int z = 0; z < gz; ++z)
        for (int y = 0; y < gy; ++y)
            for (int x = 0; x < gx; ++x) S::MainComputeShader(uint3((uint) x, (uint) y, (uint) z));
    return 0;
}

// variant 0: PerformMandelbulbRaymarchReturnDistance -> out[2*pixel]; variant 1: ...ReturnNormal -> out[4*pixel]
extern "C" int tbref_mandelbulb_march(int variant, const tbrm_mandelbulb* mb, float derivation_distance, const tbo_camera_uniforms* cam, int row_begin,
                                      int row_end, float* out) {
    namespace M = hlsl::materials;
    namespace F = hlsl::sdf_marcher;
    bind_view(cam);
    const int W = cam->width;
    const float3 center(mb->center[0], mb->center[1], mb->center[2]);
#pragma omp parallel for schedule(dynamic, 2)
    for (int iy = row_begin; iy < row_end; ++iy)
        for (int ix = 0; ix < W; ++ix) {
            const FMaterialPixelParameters mp = pixel_parameters(cam, ix, iy);
            const float4 setup = M::PerformRaymarchCubeSetup(mp);
            const size_t p = (size_t) (iy - row_begin) * W + ix;
            if (variant == 0) {
                float2 r(0.0f, 0.0f);
                if (setup.w > 0.0f)  // the cube mesh only rasterises pixels whose ray crosses it
                    r = F::PerformMandelbulbRaymarchReturnDistance(center, mb->extent, mb->power, mb->max_steps, mb->max_iterations, setup.xyz(),
                                                                   setup.w, mb->bailout, mb->high_precision_eps, mb->low_precision_eps, mp);
                out[2 * p] = r.x, out[2 * p + 1] = r.y;
            } else {
                float4 r(0.0f);
                if (setup.w > 0.0f)
                    r = F::PerformMandelbulbRaymarchReturnNormal(center, mb->extent, mb->power, mb->max_steps, mb->max_iterations, derivation_distance,
                                                                 setup.xyz(), setup.w, mb->bailout, mb->high_precision_eps, mb->low_precision_eps, mp);
                out[4 * p] = r.x, out[4 * p + 1] = r.y, out[4 * p + 2] = r.z, out[4 * p + 3] = r.w;
            }
        }
    return 0;
}

// CalculateMandelbulbSDF_RenderThread (FractalShaders.cpp:41-70): groups of 16 x 16 x 4 threads over the volume; out_fmt = tbrm_format
extern "C" int tbref_mandelbulb_sdf(const int32_t dims[3], const float center[3], float extent, float power, int out_fmt, void* out) {
    namespace S = hlsl::mandelbulb_sdf;
    if (!(extent > 0.0f)) return 0;  // EnqueueRenderCommand_CalculateMandelbulbSDF: Extent <= 0 -> nothing happens
    S::MandelbulbVolumeUAV.s = Storage{out, to_fmt(out_fmt), dims[0], dims[1], dims[2]};
    S::MandelbulbVolumeDimensions = float3((float) dims[0], (float) dims[1], (float) dims[2]);
    S::Center = float3(center[0], center[1], center[2]);
    S::Extent = extent;
    S::Power = power;
    const int gx = (dims[0] + 15) / 16 * 16, gy = (dims[1] + 15) / 16 * 16, gz = (dims[2] + 3) / 4 * 4;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int z = 0; z < gz; ++z)
        for (int y = 0; y < gy; ++y)
            for (int x = 0; x < gx; ++x) S::MainComputeShader(uint3((uint) x, (uint) y, (uint) z));
    return 0;
}
