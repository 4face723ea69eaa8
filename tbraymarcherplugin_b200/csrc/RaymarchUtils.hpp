// RaymarchUtils.hpp — C++ host mirror of the reference's operator surface, header-only over the C ABI (include/tbrm.h).
//
// Same names, argument order and error behaviour as the plugin (paths relative to the plugin root):
//   URaymarchUtils                         Source/Raymarcher/Public/Util/RaymarchUtils.h:20-99
//   FBasicRaymarchRenderingResources       Source/Raymarcher/Public/Rendering/RaymarchTypes.h:87-129
//   FDirLightParameters / FClippingPlaneParameters / FRaymarchWorldParameters   RaymarchTypes.h:20-71, 136-153
//   FWindowingParameters                   Source/VolumeTextureToolkit/Public/VolumeAsset/VolumeInfo.h:32-53
// UE types are replaced by minimal stand-ins (FVector = 3 doubles, FQuat = x,y,z,w, FTransform = {Translation, Rotation,
// Scale3D}); UObject texture pointers become the opaque tbrm_resources handle that owns the device memory.
#pragma once
#include <array>
#include <vector>
#include <cstdint>

#include "../../include/tbrm.h"

namespace tbrm_ue {

struct FVector {
    double X = 0, Y = 0, Z = 0;
    FVector() = default;
    FVector(double x, double y, double z) : X(x), Y(y), Z(z) {}
    bool operator==(const FVector& o) const { return X == o.X && Y == o.Y && Z == o.Z; }
};
struct FQuat {
    double X = 0, Y = 0, Z = 0, W = 1;
};
struct FTransform {
    FVector Translation;
    FQuat Rotation;
    FVector Scale3D{1, 1, 1};
};

// RaymarchTypes.h:20-41
struct FDirLightParameters {
    FVector LightDirection;
    float LightIntensity = 0;
    FDirLightParameters() = default;
    FDirLightParameters(FVector LightDir, float LightInt) : LightDirection(LightDir), LightIntensity(LightInt) {}
    bool operator==(const FDirLightParameters& rhs) const { return LightDirection == rhs.LightDirection && LightIntensity == rhs.LightIntensity; }
    bool operator!=(const FDirLightParameters& rhs) const { return !(*this == rhs); }
};

// RaymarchTypes.h:45-71 (defaults: "ridiculously far and facing away", RaymarchVolume.cpp:640-641)
struct FClippingPlaneParameters {
    FVector Center{0, 0, 100000};
    FVector Direction{0, 0, -1};
};

// RaymarchTypes.h:136-153
struct FRaymarchWorldParameters {
    FTransform VolumeTransform;
    FClippingPlaneParameters ClippingPlaneParameters;
};

// VolumeInfo.h:32-53
struct FWindowingParameters {
    float Center = 0.5f;
    float Width = 1.0f;
    bool LowCutoff = true;
    bool HighCutoff = true;
};

// RaymarchTypes.h:87-129. The UObject* texture members collapse into the handle; bIsInitialized keeps its meaning.
struct FBasicRaymarchRenderingResources {
    bool bIsInitialized = false;
    tbrm_resources* Handle = nullptr;  // DataVolumeTextureRef + TFTextureRef + LightVolumeRenderTarget + XYZReadWriteBuffers
    bool LightVolumeHalfResolution = false;
    FWindowingParameters WindowingParameters;
};

inline tbrm_dir_light ToC(const FDirLightParameters& l) {
    return tbrm_dir_light{{l.LightDirection.X, l.LightDirection.Y, l.LightDirection.Z}, l.LightIntensity};
}
inline tbrm_world ToC(const FRaymarchWorldParameters& w) {
    tbrm_world c{};
    const FTransform& t = w.VolumeTransform;
    c.translation[0] = t.Translation.X, c.translation[1] = t.Translation.Y, c.translation[2] = t.Translation.Z;
    c.rotation[0] = t.Rotation.X, c.rotation[1] = t.Rotation.Y, c.rotation[2] = t.Rotation.Z, c.rotation[3] = t.Rotation.W;
    c.scale[0] = t.Scale3D.X, c.scale[1] = t.Scale3D.Y, c.scale[2] = t.Scale3D.Z;
    const FClippingPlaneParameters& p = w.ClippingPlaneParameters;
    c.clip.center[0] = p.Center.X, c.clip.center[1] = p.Center.Y, c.clip.center[2] = p.Center.Z;
    c.clip.direction[0] = p.Direction.X, c.clip.direction[1] = p.Direction.Y, c.clip.direction[2] = p.Direction.Z;
    return c;
}

class URaymarchUtils {
public:
    /** Adds a light to light volume. Also works for removing a light by setting bLightAdded to false. (RaymarchUtils.h:31-35)
        bGPUSync selects the fused single-launch sweep; results are identical. */
    static void AddDirLightToSingleVolume(const FBasicRaymarchRenderingResources& Resources, const FDirLightParameters& LightParameters,
                                          const bool Added, const FRaymarchWorldParameters WorldParameters, bool& LightAdded,
                                          bool bGPUSync = false) {
        const tbrm_dir_light l = ToC(LightParameters);
        const tbrm_world w = ToC(WorldParameters);
        int added = 0;
        tbrm_add_dir_light(Resources.Handle, &l, Added ? 1 : 0, &w, &added, bGPUSync ? 1 : 0);
        LightAdded = added != 0;  // false iff a resource is missing (RaymarchUtils.cpp:39-49)
    }

    /** Changes a light in the light volume. (RaymarchUtils.h:37-41) */
    static void ChangeDirLightInSingleVolume(FBasicRaymarchRenderingResources& Resources, const FDirLightParameters OldLightParameters,
                                             const FDirLightParameters NewLightParameters, const FRaymarchWorldParameters WorldParameters,
                                             bool& LightAdded, bool bGPUSync = false) {
        const tbrm_dir_light o = ToC(OldLightParameters), n = ToC(NewLightParameters);
        const tbrm_world w = ToC(WorldParameters);
        int added = 0;
        tbrm_change_dir_light(Resources.Handle, &o, &n, &w, &added, bGPUSync ? 1 : 0);
        LightAdded = added != 0;
    }

    /** Not in the reference (its Readme.md:165-166, 186-187 names it as the missing optimisation): adds all lights at once, the passes that
        propagate from the same cube face joined into one per-slice sweep. */
    static void AddDirLightsToSingleVolumeJoined(const FBasicRaymarchRenderingResources& Resources, const std::vector<FDirLightParameters>& Lights,
                                                 const bool Added, const FRaymarchWorldParameters WorldParameters, bool& LightAdded) {
        std::vector<tbrm_dir_light> l;
        for (const FDirLightParameters& L : Lights) l.push_back(ToC(L));
        const tbrm_world w = ToC(WorldParameters);
        int added = 0;
        const tbrm_status s = tbrm_add_dir_lights_joined(Resources.Handle, l.data(), (int) l.size(), Added ? 1 : 0, &w, &added, nullptr);
        LightAdded = s == TBRM_OK;
    }

    /** Clears a light volume in provided raymarch resources. (RaymarchUtils.h:47-49) */
    static void ClearResourceLightVolumes(FBasicRaymarchRenderingResources Resources, float ClearValue) {
        if (!Resources.Handle) return;  // RaymarchUtils.cpp:106-109
        tbrm_clear_light_volume(Resources.Handle, ClearValue);
    }

    /** Default transfer function: full opacity, black at 0 to white at 1. (RaymarchUtils.h:57-60) */
    static void MakeDefaultTFTexture(FBasicRaymarchRenderingResources& Resources) { tbrm_make_default_tf(Resources.Handle); }

    /** TF texture from 256 samples of a colour curve at i/255 (RaymarchUtils.h:62-64; the curve object itself is UE's). */
    static void ColorCurveToTexture(const std::array<float, 256 * 4>& CurveSamples, FBasicRaymarchRenderingResources& Resources) {
        static thread_local float tex[16 * 256 * 4];
        for (int row = 0; row < 16; ++row)
            for (int i = 0; i < 256 * 4; ++i) tex[row * 256 * 4 + i] = CurveSamples[i];
        tbrm_set_transfer_function(Resources.Handle, tex, 256, 16);
    }

    /** ARaymarchVolume::InitializeRaymarchResources (RaymarchVolume.cpp:821-920): light volume + the 3x4 R/W buffers. */
    static bool InitializeRaymarchResources(FBasicRaymarchRenderingResources& Resources, const int32_t DataDims[3], tbrm_format DataFormat,
                                            const void* HostVolume, bool bLightVolume32Bit, int Device = 0) {
        if (Resources.bIsInitialized) FreeRaymarchResources(Resources);
        tbrm_resources* h = nullptr;
        if (tbrm_create(Device, DataDims, DataFormat, bLightVolume32Bit ? TBRM_FMT_R32F : TBRM_FMT_G8, Resources.LightVolumeHalfResolution, &h) != TBRM_OK)
            return false;
        Resources.Handle = h;
        if (HostVolume && tbrm_upload_volume(h, HostVolume, 0) != TBRM_OK) return false;
        const tbrm_windowing w{Resources.WindowingParameters.Center, Resources.WindowingParameters.Width,
                               Resources.WindowingParameters.LowCutoff ? 1 : 0, Resources.WindowingParameters.HighCutoff ? 1 : 0};
        tbrm_set_windowing(h, &w);
        tbrm_flush(h);  // FlushRenderingCommands(), RaymarchVolume.cpp:880,915
        Resources.bIsInitialized = HostVolume != nullptr;
        return true;
    }

    static void FreeRaymarchResources(FBasicRaymarchRenderingResources& Resources) {  // RaymarchVolume.cpp:922-949
        tbrm_destroy(Resources.Handle);
        Resources.Handle = nullptr;
        Resources.bIsInitialized = false;
    }

    // ---- material entry points (Custom nodes of M_Raymarch) -------------------------------------------------------
    /** PerformRaymarchCubeSetup + PerformWindowedLitRaymarch (WindowedRaymarchMaterials.usf:36-96) for the whole frame;
        OutRGBA receives premultiplied RGBA float32, W*H*4 floats, host memory. Returns false if the resources are not ready. */
    static bool PerformWindowedLitRaymarch(const FBasicRaymarchRenderingResources& Resources, const tbrm_camera& Camera,
                                           const FRaymarchWorldParameters& WorldParameters, float StepCount, float* OutRGBA,
                                           uint64_t* OutSteps = nullptr) {
        const tbrm_world w = ToC(WorldParameters);
        return tbrm_raymarch_lit(Resources.Handle, &Camera, &w, StepCount, 0, Camera.height, OutRGBA, 0, OutSteps) == TBRM_OK;
    }

    // ---- the other materials, the octree, volume ingest ---------------------------------------------
    /** Generates the octree acceleration volume of the resources. (RaymarchUtils.h:43-45) */
    static void GenerateOctree(FBasicRaymarchRenderingResources& Resources) { tbrm_generate_octree(Resources.Handle); }

    /** PerformWindowedIntensityRaymarch (WindowedRaymarchMaterials.usf:187-242), whole frame into host memory. */
    static bool PerformWindowedIntensityRaymarch(const FBasicRaymarchRenderingResources& Resources, const tbrm_camera& Camera,
                                                 const FRaymarchWorldParameters& WorldParameters, float StepCount, float* OutRGBA,
                                                 uint64_t* OutSteps = nullptr) {
        const tbrm_world w = ToC(WorldParameters);
        return tbrm_raymarch_intensity(Resources.Handle, &Camera, &w, StepCount, 0, Camera.height, OutRGBA, 0, OutSteps) == TBRM_OK;
    }

    /** PerformWindowedRaymarchOctree (WindowedRaymarchMaterials.usf:99-183) on mip OctreeMip (ARaymarchVolume::OctreeVolumeMip). */
    static bool PerformWindowedRaymarchOctree(const FBasicRaymarchRenderingResources& Resources, const tbrm_camera& Camera,
                                              const FRaymarchWorldParameters& WorldParameters, float StepCount, uint32_t OctreeMip, float* OutRGBA,
                                              uint64_t* OutSteps = nullptr) {
        const tbrm_world w = ToC(WorldParameters);
        return tbrm_raymarch_octree(Resources.Handle, &Camera, &w, StepCount, (int) OctreeMip, 0, Camera.height, OutRGBA, 0, OutSteps) == TBRM_OK;
    }

    /** UMHDLoader::CreateVolumeFromFile (MHDLoader.cpp:183-227) + InitializeRaymarchResources: header, raw / zlib data file, normalisation
        or float conversion on the GPU, resources whose data volume it is. */
    static bool CreateVolumeFromFile(FBasicRaymarchRenderingResources& Resources, const char* FileName, tbrm_volume_info& OutInfo,
                                     bool bNormalize = true, bool bConvertToFloat = true, bool bLightVolume32Bit = false, int Device = 0) {
        if (Resources.bIsInitialized) FreeRaymarchResources(Resources);
        tbrm_resources* h = nullptr;
        if (tbrm_load_mhd_volume(Device, FileName, bNormalize ? 1 : 0, bConvertToFloat ? 1 : 0, bLightVolume32Bit ? TBRM_FMT_R32F : TBRM_FMT_G8,
                                 Resources.LightVolumeHalfResolution, &OutInfo, &h) != TBRM_OK)
            return false;
        Resources.Handle = h;
        Resources.bIsInitialized = true;
        return true;
    }

    // ---- streaming (time-varying volumes): the engine streams texture updates while the render thread keeps drawing ------
    static bool SetDataVolumeAsync(FBasicRaymarchRenderingResources& Resources, const void* PinnedHostVolume) {
        return tbrm_upload_volume_async(Resources.Handle, PinnedHostVolume) == TBRM_OK;
    }
    static bool PresentDataVolume(FBasicRaymarchRenderingResources& Resources) {
        const bool ok = tbrm_present_volume(Resources.Handle) == TBRM_OK;
        Resources.bIsInitialized = Resources.bIsInitialized || ok;
        return ok;
    }
    static bool PerformWindowedLitRaymarchAsync(const FBasicRaymarchRenderingResources& Resources, const tbrm_camera& Camera,
                                                const FRaymarchWorldParameters& WorldParameters, float StepCount, float* PinnedOutRGBA) {
        const tbrm_world w = ToC(WorldParameters);
        return tbrm_raymarch_lit_to_host_async(Resources.Handle, &Camera, &w, StepCount, PinnedOutRGBA) == TBRM_OK;
    }
    static void WaitForDownloads(const FBasicRaymarchRenderingResources& Resources) { tbrm_download_wait(Resources.Handle); }

    // ---- one volume sharded over the GPUs of a box as Z-slabs (the caller runs one process / context per GPU) ------------
    /** Configure this GPU's slab (rank of nranks) following the library's partition rule. Neighbours are then connected with
        tbrm_slab_ipc_handle / tbrm_slab_open_peer (handles travel through the host's own IPC). */
    static bool ConfigureSlab(FBasicRaymarchRenderingResources& Resources, int Rank, int NumRanks) {
        int32_t dims[3];
        if (tbrm_light_volume_dims(Resources.Handle, dims) != TBRM_OK) return false;
        tbrm_slab s{Rank, NumRanks, 0, 0};
        tbrm_slab_partition(dims[2], NumRanks, Rank, &s.z_begin, &s.z_end);
        return tbrm_slab_configure(Resources.Handle, &s) == TBRM_OK;
    }
};

}  // namespace tbrm_ue
