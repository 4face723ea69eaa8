// ref_wrap.cpp — C entry points around the REFERENCE'S OWN host code, compiled by oracle/ref.mk from where it lies under
// /root/reference (nothing of it is copied into this repository). TEST INFRASTRUCTURE ONLY, like everything in oracle/.
//
// What is the reference's and what is ours:
//   * reference, unmodified:  Source/Raymarcher/Private/Rendering/LightingShaderUtils.cpp (FMajorAxes::GetMajorAxes, GetTransposedDimensions,
//     GetAxisDirection, GetUVOffset, GetStepSizeAndUVWOffset, GetLocalLightParamsAndAxes, GetBorderColorIntSingle,
//     GetLocalClippingParameters, GetLightAlpha, GetPermutationMatrix, GetLoopStartStopIndexes),
//     Source/VolumeTextureToolkit/Private/VolumeAsset/VolumeInfo.cpp (FVolumeInfo::Normalize* / Denormalize* / VoxelFormatByteSize),
//     the header templates UVolumeTextureToolkit::ConvertArrayToNormalizedArray / ConvertArrayToFloatTemplated
//     (Source/VolumeTextureToolkit/Public/TextureUtilities.h:103-178);
//   * ours: the engine-type shim in oracle/ue_shim/ (FVector, FTransform, FLinearColor ...: engine semantics restated), and this file,
//     which calls the reference functions in the order of the render-thread drivers it cannot compile
//     (AddDirLightToSingleLightVolume_RenderThread, Source/Raymarcher/Private/Rendering/LightingShaders.cpp:35-130, RHI code) and the
//     per-format dispatch of NormalizeArrayByFormat / ConvertArrayToFloat (Source/VolumeTextureToolkit/Private/TextureUtilities.cpp:304-350).
#include "Rendering/LightingShaderUtils.h"
#include "TextureUtilities.h"

#include "../include/tbrm.h"
#include "tbrm_oracle.h"

namespace {
FTransform to_transform(const tbrm_world& w) {
    FTransform t;
    t.Translation = FVector(w.translation[0], w.translation[1], w.translation[2]);
    t.Rotation.X = w.rotation[0], t.Rotation.Y = w.rotation[1], t.Rotation.Z = w.rotation[2], t.Rotation.W = w.rotation[3];
    t.Scale3D = FVector(w.scale[0], w.scale[1], w.scale[2]);
    return t;
}
// what the RHI makes of a packed sRGB border colour (policy Q2 of SURVEY.md Appendix B): byte / 255, sRGB -> linear
float border_from_packed(uint32 argb) {
    const double c = (double) ((argb >> 16) & 0xff) / 255.0;
    return (float) (c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
}
}  // namespace

// The host part of AddDirLightToSingleLightVolume_RenderThread (LightingShaders.cpp:41-130), every value computed by the
// reference's functions. data_border (a sampler set up inside the shader class, LightingShaders.h:76-94) is not produced: NaN.
extern "C" int tbref_plan_dir_light(const int32_t ldims[3], const tbrm_dir_light* light, const tbrm_world* world, tbo_light_plan* out) {
    std::memset(out, 0, sizeof(*out));
    out->data_border = std::numeric_limits<float>::quiet_NaN();
    const FDirLightParameters LightParameters(FVector(light->direction[0], light->direction[1], light->direction[2]), light->intensity);
    FRaymarchWorldParameters WorldParameters;
    WorldParameters.VolumeTransform = to_transform(*world);
    WorldParameters.ClippingPlaneParameters = FClippingPlaneParameters(FVector(world->clip.center[0], world->clip.center[1], world->clip.center[2]),
                                                                      FVector(world->clip.direction[0], world->clip.direction[1], world->clip.direction[2]));
    if (LightParameters.LightDirection == FVector(0.0, 0.0, 0.0)) {  // :41-46
        out->zero_direction = 1;
        return 0;
    }
    FDirLightParameters LocalLightParams;
    FMajorAxes LocalMajorAxes;
    GetLocalLightParamsAndAxes(LightParameters, WorldParameters.VolumeTransform, LocalLightParams, LocalMajorAxes);  // :51
    const FClippingPlaneParameters LocalClippingParameters = GetLocalClippingParameters(WorldParameters);         // :54
    out->clip_center[0] = (float) LocalClippingParameters.Center.X, out->clip_center[1] = (float) LocalClippingParameters.Center.Y,
    out->clip_center[2] = (float) LocalClippingParameters.Center.Z;
    out->clip_dir[0] = (float) LocalClippingParameters.Direction.X, out->clip_dir[1] = (float) LocalClippingParameters.Direction.Y,
    out->clip_dir[2] = (float) LocalClippingParameters.Direction.Z;
    out->local_dir[0] = LocalLightParams.LightDirection.X, out->local_dir[1] = LocalLightParams.LightDirection.Y,
    out->local_dir[2] = LocalLightParams.LightDirection.Z;
    FRHITexture3D LightVolume;
    LightVolume.SizeX = ldims[0], LightVolume.SizeY = ldims[1], LightVolume.SizeZ = ldims[2];
    out->add_passes = 0;
    bool counting = true;
    for (unsigned i = 0; i < 2; i++) {
        if (LocalMajorAxes.FaceWeight[i].second == 0) counting = false;  // Add breaks here (:65-68, 94-97); Change runs both (:242-)
        if (counting) out->add_passes = (int) i + 1;
        tbo_pass& P = out->pass[i];
        P.face = (int) LocalMajorAxes.FaceWeight[i].first;
        P.axis = P.face / 2;
        P.weight = LocalMajorAxes.FaceWeight[i].second;
        const FIntVector TransposedDimensions = GetTransposedDimensions(LocalMajorAxes, &LightVolume, i);  // :107
        P.td[0] = TransposedDimensions.X, P.td[1] = TransposedDimensions.Y, P.td[2] = TransposedDimensions.Z;
        P.light_alpha = GetLightAlpha(LocalLightParams, LocalMajorAxes, i);                                 // :76
        P.border = border_from_packed(GetBorderColorIntSingle(LocalLightParams, LocalMajorAxes, i));        // :102
        const FVector2D UVOffset = GetUVOffset(LocalMajorAxes.FaceWeight[i].first, -LocalLightParams.LightDirection, TransposedDimensions);  // :110
        P.uv_offset[0] = (float) UVOffset.X, P.uv_offset[1] = (float) UVOffset.Y;
        FVector UVWOffset;
        float StepSize;
        GetStepSizeAndUVWOffset(LocalMajorAxes.FaceWeight[i].first, -LocalLightParams.LightDirection, TransposedDimensions, WorldParameters,
                                StepSize, UVWOffset);  // :116-118
        // :121-124 — "Normalize UVW offset to length of largest voxel size"
        const int LowestVoxelCount = std::min(TransposedDimensions.X, std::min(TransposedDimensions.Y, TransposedDimensions.Z));
        const float LongestVoxelSide = 1.0f / LowestVoxelCount;
        UVWOffset.Normalize();
        UVWOffset *= LongestVoxelSide;
        P.uvw_offset[0] = (float) UVWOffset.X, P.uvw_offset[1] = (float) UVWOffset.Y, P.uvw_offset[2] = (float) UVWOffset.Z;
        P.step_size = StepSize;
        int Start, Stop, AxisDirection;
        GetLoopStartStopIndexes(Start, Stop, AxisDirection, LocalMajorAxes, i, TransposedDimensions.Z);  // :130
        P.start = Start, P.stop = Stop, P.dirn = AxisDirection;
    }
    return 0;
}

// GetLocalClippingParameters (LightingShaderUtils.cpp:205-220), as ARaymarchVolume feeds it to the materials (RaymarchVolume.cpp:705-728)
extern "C" void tbref_local_clipping(const tbrm_world* world, float center[3], float dir[3]) {
    FRaymarchWorldParameters WorldParameters;
    WorldParameters.VolumeTransform = to_transform(*world);
    WorldParameters.ClippingPlaneParameters = FClippingPlaneParameters(FVector(world->clip.center[0], world->clip.center[1], world->clip.center[2]),
                                                                      FVector(world->clip.direction[0], world->clip.direction[1], world->clip.direction[2]));
    const FClippingPlaneParameters L = GetLocalClippingParameters(WorldParameters);
    center[0] = (float) L.Center.X, center[1] = (float) L.Center.Y, center[2] = (float) L.Center.Z;
    dir[0] = (float) L.Direction.X, dir[1] = (float) L.Direction.Y, dir[2] = (float) L.Direction.Z;
}

// Border colour of the data-volume sampler: the statements of FAddDirLightShader::SetRaymarchResources (LightingShaders.h:76-94; the class
// itself is RHI code) on the shim's FLinearColor, then what the RHI makes of the packed colour (policy Q1: byte / 255, sRGB -> linear).
extern "C" float tbref_data_border(const tbrm_windowing* win, int exact) {
    FWindowingParameters WindowingParams;
    WindowingParams.Center = win->center, WindowingParams.Width = win->width;
    float ZeroTFValue = WindowingParams.Center - 0.5 * WindowingParams.Width;
    if (exact) return ZeroTFValue;
    FLinearColor VolumeClearColor = FLinearColor(ZeroTFValue, 0.0, 0.0, 0.0);
    const uint32 BorderColorInt = VolumeClearColor.ToFColor(false).ToPackedARGB();
    return border_from_packed(BorderColorInt);
}

// rows of GetPermutationMatrix(LocalMajorAxes, index) for a face: pos = px*row0 + py*row1 + Loop*row2 (AddDirLightShader.usf)
extern "C" void tbref_permutation_rows(int face, double rows[9]) {
    FMajorAxes axes;
    axes.FaceWeight.push_back(std::make_pair(FCubeFace(face), 1.0f));
    const FMatrix m = GetPermutationMatrix(axes, 0);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) rows[3 * r + c] = m.M[r][c];
}

// NormalizeArrayByFormat (TextureUtilities.cpp:304-327): fmt = EVolumeVoxelFormat. Writes bytes/voxel-out * count bytes into `out`
// (uint8 for 1-byte inputs, uint16 otherwise); returns bytes per output voxel, 0 on a bad format.
extern "C" int tbref_normalize_array(int fmt, const void* in, int64_t byte_size, void* out, float* out_min, float* out_max) {
    uint8* src = (uint8*) const_cast<void*>(in);
    uint8* res = nullptr;
    int out_bytes = 2;
    switch ((EVolumeVoxelFormat) fmt) {
        case EVolumeVoxelFormat::UnsignedChar: res = UVolumeTextureToolkit::ConvertArrayToNormalizedArray<uint8, uint8>(src, byte_size, *out_min, *out_max), out_bytes = 1; break;
        case EVolumeVoxelFormat::SignedChar: res = UVolumeTextureToolkit::ConvertArrayToNormalizedArray<int8, uint8>(src, byte_size, *out_min, *out_max), out_bytes = 1; break;
        case EVolumeVoxelFormat::UnsignedShort: res = UVolumeTextureToolkit::ConvertArrayToNormalizedArray<uint16, uint16>(src, byte_size, *out_min, *out_max); break;
        case EVolumeVoxelFormat::SignedShort: res = UVolumeTextureToolkit::ConvertArrayToNormalizedArray<int16, uint16>(src, byte_size, *out_min, *out_max); break;
        case EVolumeVoxelFormat::UnsignedInt: res = UVolumeTextureToolkit::ConvertArrayToNormalizedArray<uint32, uint16>(src, byte_size, *out_min, *out_max); break;
        case EVolumeVoxelFormat::SignedInt: res = UVolumeTextureToolkit::ConvertArrayToNormalizedArray<int32, uint16>(src, byte_size, *out_min, *out_max); break;
        case EVolumeVoxelFormat::Float: res = UVolumeTextureToolkit::ConvertArrayToNormalizedArray<float, uint16>(src, byte_size, *out_min, *out_max); break;
        default: return 0;
    }
    const int in_bytes = FVolumeInfo::VoxelFormatByteSize((EVolumeVoxelFormat) fmt);
    std::memcpy(out, res, (size_t) (byte_size / in_bytes) * out_bytes);
    delete[] res;
    return out_bytes;
}

// ConvertArrayToFloat (TextureUtilities.cpp:329-350)
extern "C" int tbref_convert_to_float(int fmt, const void* in, int32_t voxels, float* out) {
    uint8* src = (uint8*) const_cast<void*>(in);
    float* res = nullptr;
    switch ((EVolumeVoxelFormat) fmt) {
        case EVolumeVoxelFormat::UnsignedChar: res = UVolumeTextureToolkit::ConvertArrayToFloatTemplated<uint8>(src, voxels); break;
        case EVolumeVoxelFormat::SignedChar: res = UVolumeTextureToolkit::ConvertArrayToFloatTemplated<int8>(src, voxels); break;
        case EVolumeVoxelFormat::UnsignedShort: res = UVolumeTextureToolkit::ConvertArrayToFloatTemplated<uint16>(src, voxels); break;
        case EVolumeVoxelFormat::SignedShort: res = UVolumeTextureToolkit::ConvertArrayToFloatTemplated<int16>(src, voxels); break;
        case EVolumeVoxelFormat::UnsignedInt: res = UVolumeTextureToolkit::ConvertArrayToFloatTemplated<uint32>(src, voxels); break;
        case EVolumeVoxelFormat::SignedInt: res = UVolumeTextureToolkit::ConvertArrayToFloatTemplated<int32>(src, voxels); break;
        default: return 1;
    }
    std::memcpy(out, res, (size_t) voxels * sizeof(float));
    delete[] res;
    return 0;
}

// FVolumeInfo::{Normalize,Denormalize}{Value,Range} (VolumeInfo.cpp:18-55): what = 0 NormalizeValue, 1 DenormalizeValue,
// 2 NormalizeRange, 3 DenormalizeRange
extern "C" float tbref_volume_info_map(int what, int is_normalized, float min_value, float max_value, float v) {
    FVolumeInfo info;
    info.bIsNormalized = is_normalized != 0;
    info.MinValue = min_value, info.MaxValue = max_value;
    switch (what) {
        case 0: return info.NormalizeValue(v);
        case 1: return info.DenormalizeValue(v);
        case 2: return info.NormalizeRange(v);
        default: return info.DenormalizeRange(v);
    }
}
extern "C" int tbref_voxel_format_bytes(int fmt) { return FVolumeInfo::VoxelFormatByteSize((EVolumeVoxelFormat) fmt); }
