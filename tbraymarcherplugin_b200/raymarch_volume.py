"""Host-side mirror of the caller of the hot path: ARaymarchVolume's light bookkeeping and per-tick update policy.

Reference (paths relative to the plugin root):
  ARaymarchVolume::OnConstruction    Source/Raymarcher/Private/Actor/RaymarchVolume.cpp:161-173   (LightParametersMap seeding)
  ARaymarchVolume::Tick              RaymarchVolume.cpp:326-416   (world change => full reset; reset-vs-incremental rule)
  ARaymarchVolume::ResetAllLights    RaymarchVolume.cpp:418-451
  ARaymarchVolume::UpdateSingleLight RaymarchVolume.cpp:453-465
  ARaymarchVolume::GetWorldParameters RaymarchVolume.cpp:632-648
  ARaymarchLight::GetCurrentParameters Source/Raymarcher/Private/Actor/RaymarchLight.cpp:28-31

This is SURVEY.md §8(f) row 1's host part (the scheduler that decides WHICH operator of the boundary runs each frame); the
actors' engine plumbing (components, materials, editor hooks) is out of scope. Everything that touches the GPU goes through an
"operator surface" object with URaymarchUtils' methods (the real one by default), so the policy is testable without a device.

Faithfulness notes:
* The reference compares world parameters with FTransform::Equals (tolerance 1e-4 per component, RaymarchTypes.h:145-148) and
  light parameters exactly (RaymarchTypes.h:31-34); so does this mirror.
* ResetAllLights does NOT refresh LightParametersMap (RaymarchVolume.cpp:418-451): after a reset triggered while lights had
  moved, the next tick still sees them as changed and either resets again or issues ChangeDirLight from the stale "old"
  parameters. ``bRefreshLightMapOnReset=False`` (default) reproduces that; ``True`` records the parameters a reset used.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Dict, List, Optional

from .raymarch_utils import (FBasicRaymarchRenderingResources, FClippingPlaneParameters, FDirLightParameters, FRaymarchWorldParameters,
                             FTransform, URaymarchUtils)

KINDA_SMALL_NUMBER = 1e-4  # FTransform::Equals default tolerance


class ERaymarchMaterial(enum.Enum):  # RaymarchVolume.h: ERaymarchMaterial
    Lit = 0
    Intensity = 1
    Octree = 2


@dataclass(eq=False)
class ARaymarchLight:
    """Directional light actor: forward vector + intensity (RaymarchLight.cpp:28-31). Identity (not value) keys the map."""

    ForwardVector: tuple = (0.0, 0.0, -1.0)
    LightIntensity: float = 1.0
    Name: str = "RaymarchLight"

    def GetCurrentParameters(self) -> FDirLightParameters:
        return FDirLightParameters(tuple(float(c) for c in self.ForwardVector), float(self.LightIntensity))


@dataclass(eq=False)
class ARaymarchClipPlane:
    Center: tuple = (0.0, 0.0, 0.0)
    Direction: tuple = (0.0, 0.0, 1.0)

    def GetCurrentParameters(self) -> FClippingPlaneParameters:
        return FClippingPlaneParameters(tuple(map(float, self.Center)), tuple(map(float, self.Direction)))


def _light_params_equal(a: FDirLightParameters, b: FDirLightParameters) -> bool:  # RaymarchTypes.h:31-34: exact
    return tuple(a.LightDirection) == tuple(b.LightDirection) and a.LightIntensity == b.LightIntensity


def _transform_equals(a: FTransform, b: FTransform, tol: float = KINDA_SMALL_NUMBER) -> bool:
    """FTransform::Equals: translation, rotation (q or -q) and scale within the tolerance."""
    close = lambda u, v: all(abs(x - y) <= tol for x, y in zip(u, v))  # noqa: E731
    rot = close(a.Rotation, b.Rotation) or close(a.Rotation, tuple(-c for c in b.Rotation))
    return close(a.Translation, b.Translation) and rot and close(a.Scale3D, b.Scale3D)


def _world_params_equal(a: FRaymarchWorldParameters, b: FRaymarchWorldParameters) -> bool:  # RaymarchTypes.h:145-148
    ca, cb = a.ClippingPlaneParameters, b.ClippingPlaneParameters
    return _transform_equals(a.VolumeTransform, b.VolumeTransform) and tuple(ca.Center) == tuple(cb.Center) and tuple(ca.Direction) == tuple(cb.Direction)


@dataclass
class FTickReport:
    """What one Tick did (for logs, tests and benchmarks)."""

    action: str = "none"  # "none" | "not_initialized" | "reset" | "incremental"
    lights_updated: int = 0
    octree_rebuilt: bool = False
    errors: List[str] = field(default_factory=list)


class ARaymarchVolume:
    def __init__(self, RaymarchResources: FBasicRaymarchRenderingResources, LightsArray: Optional[List[Optional[ARaymarchLight]]] = None,
                 ClippingPlane: Optional[ARaymarchClipPlane] = None, VolumeTransform: Optional[FTransform] = None, ops=URaymarchUtils,
                 bRefreshLightMapOnReset: bool = False):
        self.RaymarchResources = RaymarchResources
        self.LightsArray: List[Optional[ARaymarchLight]] = list(LightsArray or [])
        self.ClippingPlane = ClippingPlane
        self.ComponentTransform = VolumeTransform or FTransform()
        self.SelectRaymarchMaterial = ERaymarchMaterial.Lit
        self.bFastShader = True  # RaymarchVolume.h:64-65
        # a full reset adds all lights at once, same-face passes joined into one per-slice sweep (not in the
        # reference; pays where the sweep is launch-bound, i.e. with bFastShader off and several lights per face)
        self.bJoinSameAxisLights = False
        self.bVisible = True
        self.bRequestedRecompute = False
        self.bRequestedOctreeRebuild = False  # RaymarchVolume.h:168-169; set by SetVolumeAsset (RaymarchVolume.cpp:553-554)
        self.OctreeVolumeMip = 0              # RaymarchVolume.h:191-193
        self.RaymarchingSteps = 150.0         # RaymarchVolume.h:188-189
        self.bRefreshLightMapOnReset = bRefreshLightMapOnReset
        self.ops = ops
        self.WorldParameters = self.GetWorldParameters()
        self.LightParametersMap: Dict[ARaymarchLight, FDirLightParameters] = {}
        self.OnConstruction()

    # RaymarchVolume.cpp:161-173
    def OnConstruction(self) -> None:
        self.LightParametersMap.clear()
        for light in self.LightsArray:
            if light is not None and light.LightIntensity > 0.0:
                self.LightParametersMap[light] = light.GetCurrentParameters()

    # RaymarchVolume.cpp:632-648
    def GetWorldParameters(self) -> FRaymarchWorldParameters:
        clip = self.ClippingPlane.GetCurrentParameters() if self.ClippingPlane else FClippingPlaneParameters((0.0, 0.0, 100000.0), (0.0, 0.0, -1.0))
        return FRaymarchWorldParameters(self.ComponentTransform, clip)

    def UpdateWorldParameters(self) -> None:
        self.WorldParameters = self.GetWorldParameters()

    # RaymarchVolume.cpp:326-416
    def Tick(self, DeltaTime: float = 0.0) -> FTickReport:
        rep = FTickReport()
        if not self.RaymarchResources.bIsInitialized or not self.bVisible:
            rep.action = "not_initialized"
            return rep
        # Volume transform changed or clipping plane moved -> need full recompute.
        if not _world_params_equal(self.WorldParameters, self.GetWorldParameters()):
            self.bRequestedRecompute = True
            self.UpdateWorldParameters()
        # RaymarchVolume.cpp:358-363
        if self.bRequestedOctreeRebuild and self.SelectRaymarchMaterial == ERaymarchMaterial.Octree:
            self.ops.GenerateOctree(self.RaymarchResources)
            self.bRequestedOctreeRebuild = False
            rep.octree_rebuilt = True
        # Only check if we need to update lights if we're using the Lit raymarch material.
        if self.SelectRaymarchMaterial != ERaymarchMaterial.Lit:
            return rep
        if self.bRequestedRecompute:
            self.ResetAllLights(rep)
            return rep
        lights_to_update: List[ARaymarchLight] = []
        for light in self.LightsArray:
            if light is None:
                continue
            if light not in self.LightParametersMap:
                self.LightParametersMap[light] = light.GetCurrentParameters()
                lights_to_update.append(light)
            elif not _light_params_equal(light.GetCurrentParameters(), self.LightParametersMap[light]):
                lights_to_update.append(light)
        # More than half lights need update -> full reset is quicker (integer division as in the reference)
        if len(lights_to_update) > 1 and len(lights_to_update) >= len(self.LightsArray) // 2:
            self.ResetAllLights(rep)
        else:
            for light in lights_to_update:
                self.UpdateSingleLight(light, rep)
                self.LightParametersMap[light] = light.GetCurrentParameters()
            if lights_to_update:
                rep.action = "incremental"
                rep.lights_updated = len(lights_to_update)
        return rep

    # RaymarchVolume.cpp:418-451
    def ResetAllLights(self, rep: Optional[FTickReport] = None) -> None:
        rep = rep if rep is not None else FTickReport()
        if not self.RaymarchResources.bIsInitialized:
            return
        self.ops.ClearResourceLightVolumes(self.RaymarchResources, 0.0)
        rep.action = "reset"
        if self.bJoinSameAxisLights:
            lights = [l for l in self.LightsArray if l is not None]
            if not self.ops.AddDirLightsToSingleVolumeJoined(self.RaymarchResources, [l.GetCurrentParameters() for l in lights], True,
                                                             self.WorldParameters):
                rep.errors.append("Error. Could not add/remove lights in volume.")
                return
            rep.lights_updated = len(lights)
            if self.bRefreshLightMapOnReset:
                for l in lights:
                    self.LightParametersMap[l] = l.GetCurrentParameters()
            self.bRequestedRecompute = False
            return
        for light in self.LightsArray:
            if light is None:
                continue
            ok = self.ops.AddDirLightToSingleVolume(self.RaymarchResources, light.GetCurrentParameters(), True, self.WorldParameters,
                                                    bGPUSync=self.bFastShader)
            if not ok:
                rep.errors.append(f"Error. Could not add/remove light {light.Name} in volume.")
                return  # bRequestedRecompute stays set: the reset is retried next tick
            rep.lights_updated += 1
            if self.bRefreshLightMapOnReset:
                self.LightParametersMap[light] = light.GetCurrentParameters()
        self.bRequestedRecompute = False

    # RaymarchVolume.cpp:453-465
    def UpdateSingleLight(self, UpdatedLight: ARaymarchLight, rep: Optional[FTickReport] = None) -> None:
        ok = self.ops.ChangeDirLightInSingleVolume(self.RaymarchResources, self.LightParametersMap[UpdatedLight], UpdatedLight.GetCurrentParameters(),
                                                   self.WorldParameters, bGPUSync=self.bFastShader)
        if not ok and rep is not None:
            rep.errors.append(f"Error. Could not change light {UpdatedLight.Name} in volume.")

    # ARaymarchVolume::SetVolumeAsset, RaymarchVolume.cpp:467-560 (the part that concerns the path): new data => everything is stale
    def OnVolumeLoaded(self) -> None:
        self.UpdateWorldParameters()
        self.bRequestedRecompute = True
        self.bRequestedOctreeRebuild = True

    def SetVolumeAsset(self, RaymarchResources: FBasicRaymarchRenderingResources, ImageInfo=None, TransferFuncCurve=None,
                       DefaultWindowingParameters=None) -> bool:
        """SetVolumeAsset with the asset taken apart: resources whose data volume is the asset's texture (e.g. from
        UMHDLoader.CreateVolumeFromFile), its FVolumeInfo, TF curve (256 x RGBA; None: MakeDefaultTFTexture, :511-514) and default
        windowing. The mesh scale becomes WorldDimensions / 10 — "Unreal units are in cm, MHD and Dicoms both have sizes in mm" (:545-546) —
        and a full light recompute plus an octree rebuild are requested (:549-554)."""
        if RaymarchResources is None or RaymarchResources.handle is None:
            return False
        self.RaymarchResources = RaymarchResources
        if TransferFuncCurve is not None:
            self.ops.ColorCurveToTexture(RaymarchResources, TransferFuncCurve)
        else:
            self.ops.MakeDefaultTFTexture(RaymarchResources)
        RaymarchResources.bIsInitialized = True
        if DefaultWindowingParameters is not None:
            RaymarchResources.WindowingParameters = DefaultWindowingParameters
            self.ops.SetWindowingParameters(RaymarchResources, DefaultWindowingParameters)
        if ImageInfo is not None:
            t = self.ComponentTransform
            self.ComponentTransform = FTransform(t.Translation, t.Rotation, tuple(float(d) / 10.0 for d in ImageInfo.WorldDimensions))
        self.OnVolumeLoaded()
        return True

    # ---- setters: what changes the light volume requests a recompute (RaymarchVolume.cpp:562-577, 746-818) -------------------------
    def _set_windowing(self, **kw) -> None:
        w = self.RaymarchResources.WindowingParameters
        if all(getattr(w, k) == v for k, v in kw.items()):
            return  # unchanged: nothing happens (:748-749 ...)
        for k, v in kw.items():
            setattr(w, k, v)
        self.ops.SetWindowingParameters(self.RaymarchResources, w)  # SetMaterialWindowingParameters + the compute shaders' uniform
        self.bRequestedRecompute = True

    def SetWindowCenter(self, Center: float) -> None:
        self._set_windowing(Center=Center)

    def SetWindowWidth(self, Width: float) -> None:
        self._set_windowing(Width=Width)

    def SetLowCutoff(self, LowCutoff: bool) -> None:
        self._set_windowing(LowCutoff=LowCutoff)

    def SetHighCutoff(self, HighCutoff: bool) -> None:
        self._set_windowing(HighCutoff=HighCutoff)

    def GetWindowCenter(self) -> float:
        return self.RaymarchResources.WindowingParameters.Center

    def GetWindowWidth(self) -> float:
        return self.RaymarchResources.WindowingParameters.Width

    def SetTFCurve(self, InTFCurve) -> None:
        if InTFCurve is None:
            return
        self.ops.ColorCurveToTexture(self.RaymarchResources, InTFCurve)
        self.ops.FlushRenderingCommands(self.RaymarchResources)
        self.bRequestedRecompute = True

    def SwitchRenderer(self, InSelectRaymarchMaterial: ERaymarchMaterial) -> None:
        self.SelectRaymarchMaterial = InSelectRaymarchMaterial

    def SetRaymarchSteps(self, InRaymarchingSteps: float) -> None:
        self.RaymarchingSteps = float(InRaymarchingSteps)

    def LoadMHDFileIntoVolumeNormalized(self, FileName: str, loader=None, **kw) -> bool:
        """CreateVolumeFromFile(FileName, bNormalize = true, bConvertToFloat = false) + SetVolumeAsset (:613-628)."""
        return self._load_mhd(FileName, True, False, loader, **kw)

    def LoadMHDFileIntoVolumeTransientR32F(self, FileName: str, loader=None, **kw) -> bool:
        """CreateVolumeFromFile(FileName, false, true) + SetVolumeAsset (:596-611)."""
        return self._load_mhd(FileName, False, True, loader, **kw)

    def _load_mhd(self, FileName, normalize, to_float, loader, **kw) -> bool:
        if loader is None:
            from .raymarch_utils import UMHDLoader as loader
        try:
            res, info = loader.CreateVolumeFromFile(FileName, bNormalize=normalize, bConvertToFloat=to_float, **kw)
        except Exception:  # the reference returns false when no asset could be created
            return False
        old = self.RaymarchResources
        ok = self.SetVolumeAsset(res, info)
        if ok and old is not None and old is not res:
            old.release()
        return ok

    def Render(self, Camera, rows=None):
        """What UE's renderer does with the selected material (RaymarchVolume.cpp:144-152, 789-800): the entry point of M_Raymarch,
        M_Intensity_Raymarch or M_Octree_Raymarch for every covered pixel. Returns (rgba, executed_steps)."""
        if self.SelectRaymarchMaterial == ERaymarchMaterial.Lit:
            return self.ops.PerformWindowedLitRaymarch(self.RaymarchResources, Camera, self.WorldParameters, self.RaymarchingSteps, rows=rows)
        if self.SelectRaymarchMaterial == ERaymarchMaterial.Intensity:
            return self.ops.PerformWindowedIntensityRaymarch(self.RaymarchResources, Camera, self.WorldParameters, self.RaymarchingSteps, rows=rows)
        return self.ops.PerformWindowedRaymarchOctree(self.RaymarchResources, Camera, self.WorldParameters, self.RaymarchingSteps,
                                                      self.OctreeVolumeMip, rows=rows)
