// RaymarchVolume.hpp — C++ host mirror of the caller of the hot path: ARaymarchVolume's light bookkeeping and per-tick update
// policy, header-only over RaymarchUtils.hpp. Twin of tbraymarcherplugin_b200/raymarch_volume.py.
//
// Reference (paths relative to the plugin root):
//   ARaymarchVolume::OnConstruction     Source/Raymarcher/Private/Actor/RaymarchVolume.cpp:161-173
//   ARaymarchVolume::Tick               RaymarchVolume.cpp:326-416   (world change => full reset; reset-vs-incremental rule)
//   ARaymarchVolume::ResetAllLights     RaymarchVolume.cpp:418-451
//   ARaymarchVolume::UpdateSingleLight  RaymarchVolume.cpp:453-465
//   ARaymarchVolume::GetWorldParameters RaymarchVolume.cpp:632-648
//   ARaymarchLight::GetCurrentParameters Source/Raymarcher/Private/Actor/RaymarchLight.cpp:28-31
// The engine plumbing of the actors (components, materials, editor hooks) is out of scope; what is mirrored is the decision
// WHICH operator of the boundary runs each frame. `Ops` is the operator surface (URaymarchUtils by default) so that the policy
// can be exercised without a device.
#pragma once
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "RaymarchUtils.hpp"

namespace tbrm_ue {

enum class ERaymarchMaterial { Lit, Intensity, Octree };

struct ARaymarchLight {
    FVector ForwardVector{0, 0, -1};
    float LightIntensity = 1.0f;
    std::string Name = "RaymarchLight";
    FDirLightParameters GetCurrentParameters() const { return FDirLightParameters(ForwardVector, LightIntensity); }  // RaymarchLight.cpp:28-31
};

struct ARaymarchClipPlane {
    FVector Center{0, 0, 0};
    FVector Direction{0, 0, 1};
    FClippingPlaneParameters GetCurrentParameters() const { return FClippingPlaneParameters{Center, Direction}; }
};

// FTransform::Equals (tolerance KINDA_SMALL_NUMBER per component; a rotation equals its negated quaternion)
inline bool TransformEquals(const FTransform& a, const FTransform& b, double tol = 1e-4) {
    auto c3 = [&](const FVector& u, const FVector& v) { return std::fabs(u.X - v.X) <= tol && std::fabs(u.Y - v.Y) <= tol && std::fabs(u.Z - v.Z) <= tol; };
    auto c4 = [&](const FQuat& u, double s, const FQuat& v) {
        return std::fabs(u.X - s * v.X) <= tol && std::fabs(u.Y - s * v.Y) <= tol && std::fabs(u.Z - s * v.Z) <= tol && std::fabs(u.W - s * v.W) <= tol;
    };
    return c3(a.Translation, b.Translation) && (c4(a.Rotation, 1.0, b.Rotation) || c4(a.Rotation, -1.0, b.Rotation)) && c3(a.Scale3D, b.Scale3D);
}
// RaymarchTypes.h:145-148
inline bool WorldParametersEqual(const FRaymarchWorldParameters& a, const FRaymarchWorldParameters& b) {
    return TransformEquals(a.VolumeTransform, b.VolumeTransform) && a.ClippingPlaneParameters.Center == b.ClippingPlaneParameters.Center &&
           a.ClippingPlaneParameters.Direction == b.ClippingPlaneParameters.Direction;
}

struct FTickReport {
    enum Action { None, NotInitialized, Reset, Incremental } action = None;
    int lights_updated = 0;
    bool octree_rebuilt = false;
    std::vector<std::string> errors;
};

template <typename Ops = URaymarchUtils>
class ARaymarchVolume {
public:
    FBasicRaymarchRenderingResources RaymarchResources;
    std::vector<ARaymarchLight*> LightsArray;
    std::map<ARaymarchLight*, FDirLightParameters> LightParametersMap;
    ARaymarchClipPlane* ClippingPlane = nullptr;
    FTransform ComponentTransform;
    FRaymarchWorldParameters WorldParameters;
    ERaymarchMaterial SelectRaymarchMaterial = ERaymarchMaterial::Lit;
    bool bFastShader = true;  // RaymarchVolume.h:64-65
    bool bVisible = true;
    bool bRequestedRecompute = false;
    bool bRequestedOctreeRebuild = false;  // RaymarchVolume.h:168-169; requested by SetVolumeAsset (RaymarchVolume.cpp:553-554)
    uint32_t OctreeVolumeMip = 0;          // RaymarchVolume.h:191-193
    // The reference's ResetAllLights does not refresh LightParametersMap (RaymarchVolume.cpp:418-451): lights that had moved
    // when a reset ran are seen as changed again next tick. false reproduces that, true records what the reset used.
    bool bRefreshLightMapOnReset = false;

    void OnConstruction() {  // RaymarchVolume.cpp:161-173
        WorldParameters = GetWorldParameters();
        LightParametersMap.clear();
        for (ARaymarchLight* Light : LightsArray)
            if (Light && Light->LightIntensity > 0.0f) LightParametersMap[Light] = Light->GetCurrentParameters();
    }

    FRaymarchWorldParameters GetWorldParameters() const {  // RaymarchVolume.cpp:632-648
        FRaymarchWorldParameters r;
        if (ClippingPlane) r.ClippingPlaneParameters = ClippingPlane->GetCurrentParameters();
        r.VolumeTransform = ComponentTransform;
        return r;
    }

    FTickReport Tick(float /*DeltaTime*/ = 0.0f) {  // RaymarchVolume.cpp:326-416
        FTickReport rep;
        if (!RaymarchResources.bIsInitialized || !bVisible) {
            rep.action = FTickReport::NotInitialized;
            return rep;
        }
        if (!WorldParametersEqual(WorldParameters, GetWorldParameters())) {
            bRequestedRecompute = true;
            WorldParameters = GetWorldParameters();
        }
        if (bRequestedOctreeRebuild && SelectRaymarchMaterial == ERaymarchMaterial::Octree) {  // RaymarchVolume.cpp:358-363
            Ops::GenerateOctree(RaymarchResources);
            bRequestedOctreeRebuild = false;
            rep.octree_rebuilt = true;
        }
        if (SelectRaymarchMaterial != ERaymarchMaterial::Lit) return rep;
        if (bRequestedRecompute) {
            ResetAllLights(&rep);
            return rep;
        }
        std::vector<ARaymarchLight*> LightsToUpdate;
        for (ARaymarchLight* Light : LightsArray) {
            if (!Light) continue;
            auto it = LightParametersMap.find(Light);
            if (it == LightParametersMap.end()) {
                LightParametersMap[Light] = Light->GetCurrentParameters();
                LightsToUpdate.push_back(Light);
            } else if (Light->GetCurrentParameters() != it->second) {
                LightsToUpdate.push_back(Light);
            }
        }
        // More than half lights need update -> full reset is quicker
        if (LightsToUpdate.size() > 1 && LightsToUpdate.size() >= LightsArray.size() / 2) {
            ResetAllLights(&rep);
        } else {
            for (ARaymarchLight* UpdatedLight : LightsToUpdate) {
                UpdateSingleLight(UpdatedLight, &rep);
                LightParametersMap[UpdatedLight] = UpdatedLight->GetCurrentParameters();
            }
            if (!LightsToUpdate.empty()) {
                rep.action = FTickReport::Incremental;
                rep.lights_updated = (int) LightsToUpdate.size();
            }
        }
        return rep;
    }

    void ResetAllLights(FTickReport* rep = nullptr) {  // RaymarchVolume.cpp:418-451
        if (!RaymarchResources.bIsInitialized) return;
        Ops::ClearResourceLightVolumes(RaymarchResources, 0.0f);
        if (rep) rep->action = FTickReport::Reset;
        for (ARaymarchLight* Light : LightsArray) {
            if (!Light) continue;
            bool bResetWasSuccessful = true;
            Ops::AddDirLightToSingleVolume(RaymarchResources, Light->GetCurrentParameters(), true, WorldParameters, bResetWasSuccessful, bFastShader);
            if (!bResetWasSuccessful) {
                if (rep) rep->errors.push_back("Error. Could not add/remove light " + Light->Name + " in volume.");
                return;  // bRequestedRecompute stays set: retried next tick
            }
            if (rep) rep->lights_updated += 1;
            if (bRefreshLightMapOnReset) LightParametersMap[Light] = Light->GetCurrentParameters();
        }
        bRequestedRecompute = false;
    }

    void UpdateSingleLight(ARaymarchLight* UpdatedLight, FTickReport* rep = nullptr) {  // RaymarchVolume.cpp:453-465
        bool bLightAddWasSuccessful = false;
        Ops::ChangeDirLightInSingleVolume(RaymarchResources, LightParametersMap[UpdatedLight], UpdatedLight->GetCurrentParameters(), WorldParameters,
                                          bLightAddWasSuccessful, bFastShader);
        if (!bLightAddWasSuccessful && rep) rep->errors.push_back("Error. Could not change light " + UpdatedLight->Name + " in volume.");
    }
};

}  // namespace tbrm_ue
