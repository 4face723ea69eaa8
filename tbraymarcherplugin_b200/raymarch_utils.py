"""Python mirror of the reference's operator surface, bound to libtbrm.so through the C ABI.

Names, argument meaning and error behaviour follow the plugin (paths relative to the plugin root):

* ``URaymarchUtils``                      Source/Raymarcher/Public/Util/RaymarchUtils.h:20-99
* ``FBasicRaymarchRenderingResources``    Source/Raymarcher/Public/Rendering/RaymarchTypes.h:87-129
* ``FDirLightParameters``                 RaymarchTypes.h:20-41
* ``FClippingPlaneParameters``            RaymarchTypes.h:45-71
* ``FRaymarchWorldParameters``            RaymarchTypes.h:136-153
* ``FWindowingParameters``                Source/VolumeTextureToolkit/Public/VolumeAsset/VolumeInfo.h:32-53

``bool& LightAdded`` out-parameters become return values. The C++ twin with the reference's exact signatures is
``tbraymarcherplugin_b200/csrc/RaymarchUtils.hpp``. This module is host-side glue only: every computation
happens in the CUDA library; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import FMT_G8, FMT_G16, FMT_R32F, TbrmError, check

Vec3 = Tuple[float, float, float]

_NP_OF_FMT = {FMT_G8: np.uint8, FMT_G16: np.uint16, FMT_R32F: np.float32}
_FMT_OF_NP = {np.dtype(np.uint8): FMT_G8, np.dtype(np.uint16): FMT_G16, np.dtype(np.float32): FMT_R32F}


# --------------------------------------------------------------------------------------------------------------
# parameter structs
# --------------------------------------------------------------------------------------------------------------
@dataclass
class FDirLightParameters:
    LightDirection: Vec3 = (0.0, 0.0, 0.0)
    LightIntensity: float = 0.0

    def to_c(self) -> _capi.DirLight:
        return _capi.DirLight((C.c_double * 3)(*map(float, self.LightDirection)), float(self.LightIntensity))


@dataclass
class FClippingPlaneParameters:
    # RaymarchVolume.cpp:640-641: "ridiculously far and facing away" when there is no clipping plane
    Center: Vec3 = (0.0, 0.0, 100000.0)
    Direction: Vec3 = (0.0, 0.0, -1.0)


@dataclass
class FTransform:
    Translation: Vec3 = (0.0, 0.0, 0.0)
    Rotation: Tuple[float, float, float, float] = (0.0, 0.0, 0.0, 1.0)  # quaternion x, y, z, w
    Scale3D: Vec3 = (1.0, 1.0, 1.0)

    @staticmethod
    def from_axis_angle(axis: Vec3, degrees: float, translation: Vec3 = (0, 0, 0), scale: Vec3 = (1, 1, 1)) -> "FTransform":
        n = math.sqrt(sum(a * a for a in axis))
        h = math.radians(degrees) / 2.0
        s = math.sin(h) / n
        return FTransform(tuple(map(float, translation)), (axis[0] * s, axis[1] * s, axis[2] * s, math.cos(h)), tuple(map(float, scale)))


@dataclass
class FRaymarchWorldParameters:
    VolumeTransform: FTransform = field(default_factory=FTransform)
    ClippingPlaneParameters: FClippingPlaneParameters = field(default_factory=FClippingPlaneParameters)

    def to_c(self) -> _capi.World:
        t = self.VolumeTransform
        c = self.ClippingPlaneParameters
        return _capi.World(
            (C.c_double * 3)(*map(float, t.Translation)),
            (C.c_double * 4)(*map(float, t.Rotation)),
            (C.c_double * 3)(*map(float, t.Scale3D)),
            _capi.ClipPlane((C.c_double * 3)(*map(float, c.Center)), (C.c_double * 3)(*map(float, c.Direction))),
        )


@dataclass
class FWindowingParameters:
    Center: float = 0.5
    Width: float = 1.0
    LowCutoff: bool = True
    HighCutoff: bool = True

    def to_c(self) -> _capi.Windowing:
        return _capi.Windowing(float(self.Center), float(self.Width), int(bool(self.LowCutoff)), int(bool(self.HighCutoff)))


@dataclass
class FCamera:
    """Explicit pinhole camera standing in for the UE view the material reads."""

    Eye: Vec3 = (-0.9, -0.5, 0.7)
    LookAt: Vec3 = (0.0, 0.0, 0.0)
    Up: Vec3 = (0.0, 0.0, 1.0)
    HFovDeg: float = 60.0
    Width: int = 512
    Height: int = 512
    SceneDepth: float = 0.0
    FrameIndex: int = 0
    Jitter: bool = True

    def to_c(self) -> _capi.Camera:
        return _capi.Camera(
            (C.c_double * 3)(*map(float, self.Eye)),
            (C.c_double * 3)(*map(float, self.LookAt)),
            (C.c_double * 3)(*map(float, self.Up)),
            float(self.HFovDeg),
            int(self.Width),
            int(self.Height),
            float(self.SceneDepth),
            int(self.FrameIndex),
            int(bool(self.Jitter)),
        )


@dataclass
class FMandelbulbParameters:
    """Arguments of PerformMandelbulbRaymarchReturnDistance (SDFMarcher.usf:61-72); defaults per SURVEY.md §8(d)."""

    VolumeCenter: Vec3 = (0.0, 0.0, 0.0)
    Extent: float = 2.4
    Power: float = 8.0
    MaxSteps: float = 1024.0
    MaxIterations: float = 16.0
    Bailout: float = 2.4
    HighPrecisionEps: float = 1e-4
    LowPrecisionEps: float = 1e-2

    def to_c(self) -> _capi.Mandelbulb:
        return _capi.Mandelbulb(
            (C.c_float * 3)(*map(float, self.VolumeCenter)),
            self.Extent,
            self.Power,
            self.MaxSteps,
            self.MaxIterations,
            self.Bailout,
            self.HighPrecisionEps,
            self.LowPrecisionEps,
        )


@dataclass
class FSweepStats:
    passes: int = 0
    fell_back: bool = False
    voxels: int = 0
    kernel_launches: int = 0
    faces: Tuple[int, ...] = ()
    impl: Tuple[int, ...] = ()  # 1 per-slice launches, 2 fused (generic), 3 fused (TMA-staged)


# --------------------------------------------------------------------------------------------------------------
# resources
# --------------------------------------------------------------------------------------------------------------
class FBasicRaymarchRenderingResources:
    """Per-volume GPU resources: data volume, TF texture, light volume, windowing, propagation buffers.

    Created by :meth:`URaymarchUtils.InitializeRaymarchResources` (ARaymarchVolume::InitializeRaymarchResources,
    RaymarchVolume.cpp:821-920). A default-constructed object is *uninitialised*: ops return LightAdded = False.
    """

    def __init__(self) -> None:
        self._h: Optional[C.c_void_p] = None
        self.bIsInitialized = False
        self.LightVolumeHalfResolution = False
        self.bLightVolume32Bit = False
        self.WindowingParameters = FWindowingParameters()
        self.DataDims: Tuple[int, int, int] = (0, 0, 0)
        self.LightDims: Tuple[int, int, int] = (0, 0, 0)
        self.DataFormat = FMT_G8
        self.LightFormat = FMT_G8
        self.Device = 0

    @property
    def handle(self):
        return self._h

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def release(self) -> None:  # FreeRaymarchResources, RaymarchVolume.cpp:922-949
        if self._h is not None:
            _capi.load().tbrm_destroy(self._h)
            self._h = None
        self.bIsInitialized = False


class URaymarchUtils:
    """Blueprint function library of the plugin, as static methods."""

    # ---- resource set-up ------------------------------------------------------------------------------------
    @staticmethod
    def InitializeRaymarchResources(
        data_dims: Sequence[int],
        data_format: int = FMT_G8,
        bLightVolume32Bit: bool = False,
        LightVolumeHalfResolution: bool = False,
        device: int = 0,
    ) -> FBasicRaymarchRenderingResources:
        lib = _capi.load()
        res = FBasicRaymarchRenderingResources()
        dims = (C.c_int32 * 3)(*map(int, data_dims))
        h = C.c_void_p()
        light_fmt = FMT_R32F if bLightVolume32Bit else FMT_G8  # RaymarchVolume.cpp:857-861
        check(lib.tbrm_create(device, dims, data_format, light_fmt, int(LightVolumeHalfResolution), C.byref(h)))
        res._h = h
        res.Device = device
        res.DataDims = tuple(int(d) for d in data_dims)
        ld = (C.c_int32 * 3)()
        check(lib.tbrm_light_volume_dims(h, ld))
        res.LightDims = (ld[0], ld[1], ld[2])
        res.DataFormat = data_format
        res.LightFormat = light_fmt
        res.bLightVolume32Bit = bLightVolume32Bit
        res.LightVolumeHalfResolution = LightVolumeHalfResolution
        return res

    @staticmethod
    def SetDataVolume(Resources: FBasicRaymarchRenderingResources, volume: np.ndarray) -> None:
        """Upload the data volume; ``volume`` is indexed [z, y, x] (x fastest, TextureUtilities.cpp:43-78)."""
        v = np.ascontiguousarray(volume)
        X, Y, Z = Resources.DataDims
        if v.shape != (Z, Y, X) or _FMT_OF_NP.get(v.dtype) != Resources.DataFormat:
            raise ValueError(f"volume must have shape {(Z, Y, X)} and dtype {_NP_OF_FMT[Resources.DataFormat]}")
        check(_capi.load().tbrm_upload_volume(Resources.handle, v.ctypes.data_as(C.c_void_p), 0))
        check(_capi.load().tbrm_flush(Resources.handle))
        Resources.bIsInitialized = True

    @staticmethod
    def SetDataVolumeAsync(Resources: FBasicRaymarchRenderingResources, volume: np.ndarray) -> None:
        """Streaming upload of the NEXT data volume into the back buffer (returns immediately; pinned memory overlaps with the
        render queue). ``volume`` must stay alive until PresentDataVolume + FlushRenderingCommands (or the next call)."""
        X, Y, Z = Resources.DataDims
        if volume.shape != (Z, Y, X) or _FMT_OF_NP.get(volume.dtype) != Resources.DataFormat or not volume.flags["C_CONTIGUOUS"]:
            raise ValueError(f"volume must be C-contiguous with shape {(Z, Y, X)} and dtype {_NP_OF_FMT[Resources.DataFormat]}")
        check(_capi.load().tbrm_upload_volume_async(Resources.handle, volume.ctypes.data_as(C.c_void_p)))

    @staticmethod
    def PresentDataVolume(Resources: FBasicRaymarchRenderingResources) -> None:
        """The render queue waits for the last SetDataVolumeAsync and switches to that volume."""
        check(_capi.load().tbrm_present_volume(Resources.handle))
        Resources.bIsInitialized = True

    @staticmethod
    def SetDataVolumeDevice(Resources: FBasicRaymarchRenderingResources, device_ptr: int, copy: bool = False) -> None:
        lib = _capi.load()
        if copy:
            check(lib.tbrm_upload_volume(Resources.handle, C.c_void_p(device_ptr), 1))
        else:
            check(lib.tbrm_bind_volume_device(Resources.handle, C.c_void_p(device_ptr)))
        Resources.bIsInitialized = True

    @staticmethod
    def SetWindowingParameters(Resources: FBasicRaymarchRenderingResources, w: FWindowingParameters) -> None:
        Resources.WindowingParameters = w
        cw = w.to_c()
        check(_capi.load().tbrm_set_windowing(Resources.handle, C.byref(cw)))

    @staticmethod
    def SetOptions(Resources: FBasicRaymarchRenderingResources, border_exact: bool = False, data_addr_wrap: bool = False, sweep_impl: int = 0,
                   debug_flags: Sequence[int] = ()) -> None:
        """Engine-semantics switches and kernel selection; ``debug_flags`` fills tbrm_options.reserved."""
        r = list(debug_flags)[:5] + [0] * (5 - min(len(debug_flags), 5))
        o = _capi.Options(int(border_exact), int(data_addr_wrap), int(sweep_impl), (C.c_int32 * 5)(*r))
        check(_capi.load().tbrm_set_options(Resources.handle, C.byref(o)))

    # ---- transfer functions (RaymarchUtils.h:57-64) ------------------------------------------------------------
    @staticmethod
    def MakeDefaultTFTexture(Resources: FBasicRaymarchRenderingResources) -> None:
        check(_capi.load().tbrm_make_default_tf(Resources.handle))

    @staticmethod
    def ColorCurveToTexture(Resources: FBasicRaymarchRenderingResources, curve_rgba: np.ndarray, texture_height: int = 16) -> None:
        """``curve_rgba``: (256, 4) float32 samples of the colour curve at i/255 (RaymarchUtils.cpp:153-162)."""
        s = np.ascontiguousarray(curve_rgba, dtype=np.float32)
        if s.shape != (256, 4):
            raise ValueError("curve samples must be (256, 4)")
        tex = np.ascontiguousarray(np.broadcast_to(s, (texture_height, 256, 4)))
        check(_capi.load().tbrm_set_transfer_function(Resources.handle, tex.ctypes.data_as(C.POINTER(C.c_float)), 256, texture_height))

    # ---- light volume ops (RaymarchUtils.h:31-49) ------------------------------------------------------------
    @staticmethod
    def ClearResourceLightVolumes(Resources: FBasicRaymarchRenderingResources, ClearValue: float) -> None:
        if Resources is None or Resources.handle is None:
            return  # RaymarchUtils.cpp:106-109
        check(_capi.load().tbrm_clear_light_volume(Resources.handle, float(ClearValue)))

    @staticmethod
    def AddDirLightToSingleVolume(
        Resources: FBasicRaymarchRenderingResources,
        LightParameters: FDirLightParameters,
        Added: bool,
        WorldParameters: FRaymarchWorldParameters,
        bGPUSync: bool = False,
        stats: Optional[FSweepStats] = None,
    ) -> bool:
        """Returns LightAdded (RaymarchUtils.cpp:35-68): False iff the resources are not initialised."""
        lib = _capi.load()
        light, world = LightParameters.to_c(), WorldParameters.to_c()
        added = C.c_int(0)
        st = _capi.SweepStats()
        status = lib.tbrm_add_dir_light_stats(Resources.handle, C.byref(light), int(Added), C.byref(world), C.byref(added), int(bGPUSync), C.byref(st))
        if status == _capi.TBRM_ERR_NOT_INITIALIZED:
            return False
        check(status)
        _fill_stats(stats, st)
        return bool(added.value)

    @staticmethod
    def ChangeDirLightInSingleVolume(
        Resources: FBasicRaymarchRenderingResources,
        OldLightParameters: FDirLightParameters,
        NewLightParameters: FDirLightParameters,
        WorldParameters: FRaymarchWorldParameters,
        bGPUSync: bool = False,
        stats: Optional[FSweepStats] = None,
    ) -> bool:
        """Returns LightAdded (RaymarchUtils.cpp:70-92)."""
        lib = _capi.load()
        old, new, world = OldLightParameters.to_c(), NewLightParameters.to_c(), WorldParameters.to_c()
        added = C.c_int(0)
        st = _capi.SweepStats()
        status = lib.tbrm_change_dir_light_stats(
            Resources.handle, C.byref(old), C.byref(new), C.byref(world), C.byref(added), int(bGPUSync), C.byref(st)
        )
        if status == _capi.TBRM_ERR_NOT_INITIALIZED:
            return False
        check(status)
        _fill_stats(stats, st)
        return bool(added.value)

    @staticmethod
    def AddDirLightsToSingleVolumeJoined(Resources: FBasicRaymarchRenderingResources, Lights: Sequence[FDirLightParameters], Added: bool,
                                         WorldParameters: FRaymarchWorldParameters, stats: Optional[FSweepStats] = None) -> bool:
        """Same-axis light joining (the optimisation the reference's Readme says it lacks): AddDirLight for all of
        ``Lights`` with the passes that propagate from the same cube face joined into one per-slice sweep. Returns LightAdded."""
        arr = (_capi.DirLight * max(len(Lights), 1))(*[l.to_c() for l in Lights])
        world = WorldParameters.to_c()
        n, st = C.c_int(0), _capi.SweepStats()
        status = _capi.load().tbrm_add_dir_lights_joined(Resources.handle, arr, len(Lights), int(Added), C.byref(world), C.byref(n), C.byref(st))
        if status == _capi.TBRM_ERR_NOT_INITIALIZED:
            return False
        check(status)
        _fill_stats(stats, st)
        return True

    @staticmethod
    def FlushRenderingCommands(Resources: FBasicRaymarchRenderingResources) -> None:
        check(_capi.load().tbrm_flush(Resources.handle))

    @staticmethod
    def ReadLightVolume(Resources: FBasicRaymarchRenderingResources) -> np.ndarray:
        X, Y, Z = Resources.LightDims
        out = np.empty((Z, Y, X), dtype=_NP_OF_FMT[Resources.LightFormat])
        check(_capi.load().tbrm_download_light_volume(Resources.handle, out.ctypes.data_as(C.c_void_p)))
        return out

    @staticmethod
    def WriteLightVolume(Resources: FBasicRaymarchRenderingResources, light: np.ndarray) -> None:
        X, Y, Z = Resources.LightDims
        v = np.ascontiguousarray(light, dtype=_NP_OF_FMT[Resources.LightFormat])
        if v.shape != (Z, Y, X):
            raise ValueError("light volume shape mismatch")
        check(_capi.load().tbrm_upload_light_volume(Resources.handle, v.ctypes.data_as(C.c_void_p)))

    # ---- material entry points ----------------------------------------------------------------------------
    @staticmethod
    def PerformRaymarchCubeSetup(Resources: FBasicRaymarchRenderingResources, Camera: FCamera, WorldParameters: FRaymarchWorldParameters) -> np.ndarray:
        cam, world = Camera.to_c(), WorldParameters.to_c()
        out = np.empty((Camera.Height, Camera.Width, 4), dtype=np.float32)
        check(_capi.load().tbrm_raymarch_cube_setup(Resources.handle, C.byref(cam), C.byref(world), out.ctypes.data_as(C.c_void_p), 0))
        return out

    @staticmethod
    def PerformWindowedLitRaymarch(
        Resources: FBasicRaymarchRenderingResources,
        Camera: FCamera,
        WorldParameters: FRaymarchWorldParameters,
        StepCount: float,
        rows: Optional[Tuple[int, int]] = None,
        out: Optional[np.ndarray] = None,
        device_out_ptr: Optional[int] = None,
        count_steps: bool = True,
    ):
        """Cube setup + lit march for image rows [rows[0], rows[1]). Returns (rgba, executed_steps)."""
        cam, world = Camera.to_c(), WorldParameters.to_c()
        r0, r1 = rows if rows is not None else (0, Camera.Height)
        steps = C.c_uint64(0)
        psteps = C.byref(steps) if count_steps else None
        if device_out_ptr is not None:
            check(_capi.load().tbrm_raymarch_lit(Resources.handle, C.byref(cam), C.byref(world), float(StepCount), r0, r1, C.c_void_p(device_out_ptr), 1, psteps))
            return None, int(steps.value)
        if out is None:
            out = np.empty((r1 - r0, Camera.Width, 4), dtype=np.float32)
        check(_capi.load().tbrm_raymarch_lit(Resources.handle, C.byref(cam), C.byref(world), float(StepCount), r0, r1, out.ctypes.data_as(C.c_void_p), 0, psteps))
        return out, int(steps.value)

    @staticmethod
    def PerformWindowedLitRaymarchAsync(Resources: FBasicRaymarchRenderingResources, Camera: FCamera, WorldParameters: FRaymarchWorldParameters,
                                        StepCount: float, out: np.ndarray) -> None:
        """Whole frame into ``out`` ((H, W, 4) float32, pinned for a true overlap) through the download stream; returns
        immediately. Read ``out`` after WaitForDownloads."""
        if out.shape != (Camera.Height, Camera.Width, 4) or out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous (H, W, 4) float32 array")
        cam, world = Camera.to_c(), WorldParameters.to_c()
        check(_capi.load().tbrm_raymarch_lit_to_host_async(Resources.handle, C.byref(cam), C.byref(world), float(StepCount), out.ctypes.data_as(C.c_void_p)))

    @staticmethod
    def WaitForDownloads(Resources: FBasicRaymarchRenderingResources) -> None:
        check(_capi.load().tbrm_download_wait(Resources.handle))

    @staticmethod
    def PerformMandelbulbRaymarchReturnDistance(
        Params: FMandelbulbParameters,
        Camera: FCamera,
        WorldParameters: FRaymarchWorldParameters,
        rows: Optional[Tuple[int, int]] = None,
        device: int = 0,
        device_out_ptr: Optional[int] = None,
    ):
        cam, world, mb = Camera.to_c(), WorldParameters.to_c(), Params.to_c()
        r0, r1 = rows if rows is not None else (0, Camera.Height)
        iters = C.c_uint64(0)
        if device_out_ptr is not None:
            check(_capi.load().tbrm_mandelbulb_march(device, C.byref(mb), C.byref(cam), C.byref(world), r0, r1, C.c_void_p(device_out_ptr), 1, C.byref(iters)))
            return None, int(iters.value)
        out = np.empty((r1 - r0, Camera.Width, 2), dtype=np.float32)
        check(_capi.load().tbrm_mandelbulb_march(device, C.byref(mb), C.byref(cam), C.byref(world), r0, r1, out.ctypes.data_as(C.c_void_p), 0, C.byref(iters)))
        return out, int(iters.value)


    # ---- the small helpers of the function library (RaymarchUtils.cpp:219-252) ----------------------------------------
    @staticmethod
    def GetVolumeTextureDimensions(Resources: Optional[FBasicRaymarchRenderingResources]) -> Tuple[int, int, int]:
        """(0, 0, 0) for a missing texture (RaymarchUtils.cpp:219-230)."""
        return tuple(Resources.DataDims) if Resources is not None and Resources.handle is not None else (0, 0, 0)

    @staticmethod
    def LocalToTextureCoords(LocalCoords: Vec3) -> Vec3:  # RaymarchUtils.cpp:244-247: the mesh is [-1,1]^3 there
        return tuple(c / 2.0 + 0.5 for c in LocalCoords)

    @staticmethod
    def TextureToLocalCoords(TextureCoords: Vec3) -> Vec3:  # RaymarchUtils.cpp:249-252
        return tuple((c - 0.5) * 2.0 for c in TextureCoords)

    @staticmethod
    def TransformToMatrix(Transform: FTransform, WithScaling: bool = True) -> np.ndarray:
        """FTransform::ToMatrixWithScale / ToMatrixNoScale (RaymarchUtils.cpp:232-242): 4x4, row-vector convention (rows = axes, row 3 =
        translation)."""
        x, y, z, w = (float(c) for c in Transform.Rotation)
        sx, sy, sz = (float(c) for c in Transform.Scale3D) if WithScaling else (1.0, 1.0, 1.0)
        x2, y2, z2 = x + x, y + y, z + z
        xx2, yy2, zz2 = x * x2, y * y2, z * z2
        yz2, wx2, xy2, wz2, xz2, wy2 = y * z2, w * x2, x * y2, w * z2, x * z2, w * y2
        m = np.zeros((4, 4))
        m[0, :3] = ((1.0 - (yy2 + zz2)) * sx, (xy2 + wz2) * sx, (xz2 - wy2) * sx)
        m[1, :3] = ((xy2 - wz2) * sy, (1.0 - (xx2 + zz2)) * sy, (yz2 + wx2) * sy)
        m[2, :3] = ((xz2 + wy2) * sz, (yz2 - wx2) * sz, (1.0 - (xx2 + yy2)) * sz)
        m[3, :3] = tuple(float(c) for c in Transform.Translation)
        m[3, 3] = 1.0
        return m

    # ---- the other materials and the octree ----------------------------------------------
    @staticmethod
    def GenerateOctree(Resources: FBasicRaymarchRenderingResources) -> None:
        """URaymarchUtils::GenerateOctree (RaymarchUtils.h:45): fills the 4-mip UNORM16 octree volume of the resource set."""
        check(_capi.load().tbrm_generate_octree(Resources.handle))

    @staticmethod
    def ReadOctreeMip(Resources: FBasicRaymarchRenderingResources, Mip: int) -> np.ndarray:
        d = (C.c_int32 * 3)()
        check(_capi.load().tbrm_octree_mip_dims(Resources.handle, int(Mip), d))
        out = np.empty((d[2], d[1], d[0]), dtype=np.uint16)
        check(_capi.load().tbrm_download_octree_mip(Resources.handle, int(Mip), out.ctypes.data_as(C.c_void_p)))
        return out

    @staticmethod
    def PerformWindowedIntensityRaymarch(Resources: FBasicRaymarchRenderingResources, Camera: FCamera, WorldParameters: FRaymarchWorldParameters,
                                         StepCount: float, rows: Optional[Tuple[int, int]] = None):
        """WindowedRaymarchMaterials.usf:187-242. Returns (rgba, executed_steps)."""
        cam, world = Camera.to_c(), WorldParameters.to_c()
        r0, r1 = rows if rows is not None else (0, Camera.Height)
        steps = C.c_uint64(0)
        out = np.empty((r1 - r0, Camera.Width, 4), dtype=np.float32)
        check(_capi.load().tbrm_raymarch_intensity(Resources.handle, C.byref(cam), C.byref(world), float(StepCount), r0, r1,
                                                   out.ctypes.data_as(C.c_void_p), 0, C.byref(steps)))
        return out, int(steps.value)

    @staticmethod
    def PerformWindowedRaymarchOctree(Resources: FBasicRaymarchRenderingResources, Camera: FCamera, WorldParameters: FRaymarchWorldParameters,
                                      StepCount: float, OctreeMip: int = 0, rows: Optional[Tuple[int, int]] = None):
        """WindowedRaymarchMaterials.usf:99-183 (needs GenerateOctree). Returns (rgba, executed_steps)."""
        cam, world = Camera.to_c(), WorldParameters.to_c()
        r0, r1 = rows if rows is not None else (0, Camera.Height)
        steps = C.c_uint64(0)
        out = np.empty((r1 - r0, Camera.Width, 4), dtype=np.float32)
        check(_capi.load().tbrm_raymarch_octree(Resources.handle, C.byref(cam), C.byref(world), float(StepCount), int(OctreeMip), r0, r1,
                                                out.ctypes.data_as(C.c_void_p), 0, C.byref(steps)))
        return out, int(steps.value)

    # ---- Mandelbulb variants ---------------------------------------------------------------
    @staticmethod
    def PerformMandelbulbRaymarchReturnNormal(Params: FMandelbulbParameters, DerivationDistance: float, Camera: FCamera,
                                              WorldParameters: FRaymarchWorldParameters, rows: Optional[Tuple[int, int]] = None, device: int = 0):
        """SDFMarcher.usf:117-188. Returns ((rows, W, 4) normal + alpha, SDF iterations)."""
        cam, world, mb = Camera.to_c(), WorldParameters.to_c(), Params.to_c()
        r0, r1 = rows if rows is not None else (0, Camera.Height)
        iters = C.c_uint64(0)
        out = np.empty((r1 - r0, Camera.Width, 4), dtype=np.float32)
        check(_capi.load().tbrm_mandelbulb_march_normal(device, C.byref(mb), float(DerivationDistance), C.byref(cam), C.byref(world), r0, r1,
                                                        out.ctypes.data_as(C.c_void_p), 0, C.byref(iters)))
        return out, int(iters.value)

    @staticmethod
    def CalculateMandelbulbSDF(Dimensions: Sequence[int], Center: Vec3 = (0.0, 0.0, 0.0), Extent: float = 2.0, Power: float = 8.0,
                               g16: bool = True, device: int = 0):
        """EnqueueRenderCommand_CalculateMandelbulbSDF (FractalShaders.cpp:26-70). Returns ((Z, Y, X) volume, SDF iterations); the volume is
        UNORM16 like the reference's PF_G16 texture, or float32."""
        out = np.zeros(tuple(int(d) for d in Dimensions)[::-1], dtype=np.uint16 if g16 else np.float32)
        iters = C.c_uint64(0)
        check(_capi.load().tbrm_mandelbulb_sdf(device, (C.c_int32 * 3)(*map(int, Dimensions)), (C.c_float * 3)(*map(float, Center)), float(Extent),
                                               float(Power), FMT_G16 if g16 else FMT_R32F, out.ctypes.data_as(C.c_void_p), 0, C.byref(iters)))
        return out, int(iters.value)


# --------------------------------------------------------------------------------------------------------------
# volume ingest: UMHDLoader / IVolumeLoader / UVolumeTextureToolkit of the VolumeTextureToolkit module
# --------------------------------------------------------------------------------------------------------------
# EVolumeVoxelFormat (VolumeInfo.h:12-27) <-> numpy
VOXEL_DTYPES = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.uint32, 5: np.int32, 6: np.float32}
_VOXEL_OF_NP = {np.dtype(v): k for k, v in VOXEL_DTYPES.items()}


class FVolumeInfo:
    """FVolumeInfo (VolumeInfo.h:56-141) over the C struct."""

    def __init__(self, c: Optional[_capi.VolumeInfo] = None):
        self.c = c if c is not None else _capi.VolumeInfo()

    bParseWasSuccessful = property(lambda self: bool(self.c.parse_ok))
    Dimensions = property(lambda self: tuple(self.c.dims))
    Spacing = property(lambda self: tuple(self.c.spacing))
    WorldDimensions = property(lambda self: tuple(self.c.world_dims))
    OriginalFormat = property(lambda self: int(self.c.original_format))
    ActualFormat = property(lambda self: int(self.c.actual_format))
    BytesPerVoxel = property(lambda self: int(self.c.bytes_per_voxel))
    bIsSigned = property(lambda self: bool(self.c.is_signed))
    bIsNormalized = property(lambda self: bool(self.c.is_normalized))
    MinValue = property(lambda self: float(self.c.min_value))
    MaxValue = property(lambda self: float(self.c.max_value))
    bIsCompressed = property(lambda self: bool(self.c.is_compressed))
    CompressedByteSize = property(lambda self: int(self.c.compressed_bytes))
    DataFileName = property(lambda self: self.c.data_file.decode())

    def NormalizeValue(self, v: float) -> float:
        return float(_capi.load().tbrm_volume_info_normalize_value(C.byref(self.c), float(v)))

    def DenormalizeValue(self, v: float) -> float:
        return float(_capi.load().tbrm_volume_info_denormalize_value(C.byref(self.c), float(v)))

    def NormalizeRange(self, v: float) -> float:
        return float(_capi.load().tbrm_volume_info_normalize_range(C.byref(self.c), float(v)))

    def DenormalizeRange(self, v: float) -> float:
        return float(_capi.load().tbrm_volume_info_denormalize_range(C.byref(self.c), float(v)))


class UMHDLoader:
    @staticmethod
    def ParseVolumeInfoFromHeaderText(text: str) -> FVolumeInfo:
        """UMHDLoader::ParseVolumeInfoFromHeader (MHDLoader.cpp:18-181) on the header's text; bParseWasSuccessful tells the outcome."""
        info = FVolumeInfo()
        _capi.load().tbrm_mhd_parse_header(text.encode(), C.byref(info.c))
        return info

    @staticmethod
    def CreateVolumeFromFile(FileName: str, bNormalize: bool = True, bConvertToFloat: bool = True, bLightVolume32Bit: bool = False,
                             LightVolumeHalfResolution: bool = False, device: int = 0):
        """UMHDLoader::CreateVolumeFromFile (MHDLoader.cpp:183-227) + InitializeRaymarchResources: returns (resources, FVolumeInfo)."""
        lib = _capi.load()
        info = FVolumeInfo()
        h = C.c_void_p()
        light_fmt = FMT_R32F if bLightVolume32Bit else FMT_G8
        check(lib.tbrm_load_mhd_volume(device, str(FileName).encode(), int(bNormalize), int(bConvertToFloat), light_fmt, int(LightVolumeHalfResolution),
                                       C.byref(info.c), C.byref(h)))
        return _wrap_loaded_resources(h, info, light_fmt, bLightVolume32Bit, LightVolumeHalfResolution, device), info


def _wrap_loaded_resources(h, info: "FVolumeInfo", light_fmt: int, bLightVolume32Bit: bool, LightVolumeHalfResolution: bool,
                           device: int) -> FBasicRaymarchRenderingResources:
    res = FBasicRaymarchRenderingResources()
    res._h = h
    res.Device = device
    res.DataDims = info.Dimensions
    ld = (C.c_int32 * 3)()
    check(_capi.load().tbrm_light_volume_dims(h, ld))
    res.LightDims = (ld[0], ld[1], ld[2])
    res.DataFormat = {0: FMT_G8, 2: FMT_G16, 6: FMT_R32F}.get(info.ActualFormat, FMT_G8 if info.BytesPerVoxel == 1 else FMT_G16)
    res.LightFormat = light_fmt
    res.bLightVolume32Bit = bLightVolume32Bit
    res.LightVolumeHalfResolution = LightVolumeHalfResolution
    return res


class UVolumeTextureToolkit:
    @staticmethod
    def LoadRawIntoNewVolume(RawFileName: str, Dimensions: Sequence[int], dtype, bNormalize: bool = True, bConvertToFloat: bool = False,
                             CompressedByteSize: int = 0, bLightVolume32Bit: bool = False, LightVolumeHalfResolution: bool = False, device: int = 0):
        """UVolumeTextureToolkit::LoadRawIntoNewVolumeTextureAsset (TextureUtilities.h:85-101) + InitializeRaymarchResources: a headerless raw
        (or zlib, CompressedByteSize > 0) file of Dimensions voxels of numpy type dtype. Returns (resources, FVolumeInfo)."""
        info = FVolumeInfo()
        h = C.c_void_p()
        light_fmt = FMT_R32F if bLightVolume32Bit else FMT_G8
        check(_capi.load().tbrm_load_raw_volume(device, str(RawFileName).encode(), (C.c_int32 * 3)(*map(int, Dimensions)), _VOXEL_OF_NP[np.dtype(dtype)],
                                                int(CompressedByteSize), int(bNormalize), int(bConvertToFloat), light_fmt,
                                                int(LightVolumeHalfResolution), C.byref(info.c), C.byref(h)))
        return _wrap_loaded_resources(h, info, light_fmt, bLightVolume32Bit, LightVolumeHalfResolution, device), info

    @staticmethod
    def NormalizeArrayByFormat(array: np.ndarray, device: int = 0):
        """TextureUtilities.cpp:304-327 on the GPU. Returns (normalised uint8 / uint16 array, original min, original max)."""
        a = np.ascontiguousarray(array)
        fmt = _VOXEL_OF_NP[a.dtype]
        out = np.empty(a.shape, np.uint8 if a.itemsize == 1 else np.uint16)
        lo, hi = C.c_float(), C.c_float()
        check(_capi.load().tbrm_normalize_volume(device, fmt, a.ctypes.data_as(C.c_void_p), 0, a.size, out.ctypes.data_as(C.c_void_p), 0,
                                                 C.byref(lo), C.byref(hi)))
        return out, lo.value, hi.value

    @staticmethod
    def ConvertArrayToFloat(array: np.ndarray, device: int = 0) -> np.ndarray:
        """TextureUtilities.cpp:329-350 on the GPU."""
        a = np.ascontiguousarray(array)
        out = np.empty(a.shape, np.float32)
        check(_capi.load().tbrm_convert_volume_to_float(device, _VOXEL_OF_NP[a.dtype], a.ctypes.data_as(C.c_void_p), 0, a.size,
                                                        out.ctypes.data_as(C.c_void_p), 0))
        return out


def _fill_stats(dst: Optional[FSweepStats], st: _capi.SweepStats) -> None:
    if dst is None:
        return
    dst.passes = int(st.passes)
    dst.fell_back = bool(st.fell_back)
    dst.voxels = int(st.voxels)
    dst.kernel_launches = int(st.kernel_launches)
    dst.faces = tuple(int(f) for f in st.faces if f >= 0)
    dst.impl = tuple(int(st.impl[i]) for i in range(len(dst.faces)))


def plan_dir_light(light_dims: Sequence[int], windowing: FWindowingParameters, light: FDirLightParameters,
                   world: FRaymarchWorldParameters, border_exact: bool = False) -> _capi.LightPlan:
    """Host parameter math of one light (pure host code inside libtbrm.so; needs no GPU)."""
    dims = (C.c_int32 * 3)(*map(int, light_dims))
    w, l, wo = windowing.to_c(), light.to_c(), world.to_c()
    o = _capi.Options(int(border_exact), 0, 0)
    out = _capi.LightPlan()
    check(_capi.load().tbrm_plan_dir_light(dims, C.byref(w), C.byref(o), C.byref(l), C.byref(wo), C.byref(out)))
    return out
