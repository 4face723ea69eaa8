"""tbraymarcherplugin_b200 — B200-native hot path of TBRaymarcherPlugin: the AddDirLight / ChangeDirLight
illumination sweep, the windowed lit ray march and the Mandelbulb SDF march, as hand-written sm_100a CUDA behind the
plugin's own operator surface (URaymarchUtils). See DESIGN.md."""
from ._capi import FMT_G8, FMT_G16, FMT_R32F, TbrmError
from .raymarch_utils import (FBasicRaymarchRenderingResources, FCamera, FClippingPlaneParameters, FDirLightParameters,
                             FMandelbulbParameters, FRaymarchWorldParameters, FSweepStats, FTransform, FWindowingParameters,
                             URaymarchUtils, plan_dir_light)
from .raymarch_volume import ARaymarchClipPlane, ARaymarchLight, ARaymarchVolume, ERaymarchMaterial

__all__ = [
    "FMT_G8", "FMT_G16", "FMT_R32F", "TbrmError", "FBasicRaymarchRenderingResources", "FCamera", "FClippingPlaneParameters",
    "FDirLightParameters", "FMandelbulbParameters", "FRaymarchWorldParameters", "FSweepStats", "FTransform",
    "FWindowingParameters", "URaymarchUtils", "plan_dir_light", "ARaymarchVolume", "ARaymarchLight", "ARaymarchClipPlane",
    "ERaymarchMaterial",
]
