"""An ORACLE-BACKED stand-in for the operator surface (URaymarchUtils, UMHDLoader, UVolumeTextureToolkit). TEST INFRASTRUCTURE: it exists so
that the Python logic of the GPU tests (argument plumbing, shapes, tolerances, file handling) can be exercised on a machine without a GPU
(tests/test_gpu_test_logic_cpu.py) — a GPU test that has never run must not fail at round end because of a typo. It proves nothing about the
kernels and is never imported by the product."""
import pathlib
import zlib

import numpy as np

import oracle
from tbraymarcherplugin_b200 import raymarch_utils as RU
from tbraymarcherplugin_b200.raymarch_utils import FWindowingParameters

_REAL_LOADER = RU.UMHDLoader


class FakeResources:
    def __init__(self, dims, fmt, light32, half):
        self.DataDims, self.DataFormat, self.light32, self.half = tuple(dims), fmt, light32, half
        self.data, self.tf, self.win, self.vol, self.mips = None, oracle.default_tf(), FWindowingParameters(), None, None
        self.bIsInitialized, self.WindowingParameters, self.handle = False, FWindowingParameters(), object()

    def v(self) -> oracle.OracleVolume:
        if self.vol is None:
            self.vol = oracle.OracleVolume(self.data, self.tf, self.win, light32=self.light32, half_res=self.half)
        self.vol.tf, self.vol.windowing = self.tf, self.win
        return self.vol

    def release(self):
        pass


class FakeRaymarchUtils:
    @staticmethod
    def InitializeRaymarchResources(dims, fmt=0, bLightVolume32Bit=False, LightVolumeHalfResolution=False, device=0):
        return FakeResources(dims, fmt, bLightVolume32Bit, LightVolumeHalfResolution)

    @staticmethod
    def SetDataVolume(r, d):
        r.data, r.vol, r.mips = np.ascontiguousarray(d), None, None

    @staticmethod
    def ColorCurveToTexture(r, curve, texture_height=16):
        r.tf = oracle.prepare_tf(curve)

    @staticmethod
    def MakeDefaultTFTexture(r):
        r.tf = oracle.default_tf()

    @staticmethod
    def SetWindowingParameters(r, w):
        r.win = FWindowingParameters(w.Center, w.Width, w.LowCutoff, w.HighCutoff)

    @staticmethod
    def SetOptions(r, **_):
        pass

    @staticmethod
    def ClearResourceLightVolumes(r, value):
        r.v().clear(value)

    @staticmethod
    def AddDirLightToSingleVolume(r, light, added, world, bGPUSync=False, stats=None):
        n = r.v().add_dir_light(light, added, world)
        if stats is not None:
            stats.impl, stats.passes, stats.kernel_launches = (3, 3), n, 100 * n  # launch counts: only their order matters to the tests
        return True

    @staticmethod
    def AddDirLightsToSingleVolumeJoined(r, lights, added, world, stats=None):
        n = oracle.add_dir_lights_joined(r.v(), list(lights), added, world)
        if stats is not None:
            stats.passes, stats.impl, stats.kernel_launches = n, (4,) * min(n, 4), 10 * n
        return True

    @staticmethod
    def ChangeDirLightInSingleVolume(r, old, new, world, bGPUSync=False, stats=None):
        r.v().change_dir_light(old, new, world)
        return True

    @staticmethod
    def ReadLightVolume(r):
        return r.v().light.copy()

    @staticmethod
    def PerformWindowedLitRaymarch(r, cam, world, steps, rows=None, **_):
        return r.v().raymarch_lit(cam, world, steps, rows=rows)

    @staticmethod
    def PerformWindowedIntensityRaymarch(r, cam, world, steps, rows=None):
        return oracle.raymarch_intensity(r.v(), cam, world, steps, rows=rows)

    @staticmethod
    def GenerateOctree(r):
        r.mips = oracle.generate_octree(r.data)

    @staticmethod
    def ReadOctreeMip(r, mip):
        return r.mips[mip]

    @staticmethod
    def PerformWindowedRaymarchOctree(r, cam, world, steps, OctreeMip=0, rows=None):
        if r.mips is None or not 0 <= OctreeMip < 4:
            raise RU.TbrmError(2, "no octree / no such mip")
        return oracle.raymarch_octree(r.v(), cam, world, steps, r.mips, OctreeMip, rows=rows)

    @staticmethod
    def FlushRenderingCommands(r):
        pass

    @staticmethod
    def PerformMandelbulbRaymarchReturnDistance(mb, cam, world, **_):
        oracle.lib().tbo_set_mandelbulb_variant(1 if mb.Power == 8.0 else 0)
        try:
            return oracle.mandelbulb(mb, cam, world)
        finally:
            oracle.lib().tbo_set_mandelbulb_variant(0)

    @staticmethod
    def CalculateMandelbulbSDF(Dimensions, Center=(0.0, 0.0, 0.0), Extent=2.0, Power=8.0, g16=True, device=0):
        oracle.lib().tbo_set_mandelbulb_variant(1 if Power == 8.0 else 0)
        try:
            return oracle.mandelbulb_sdf(Dimensions, Center, Extent, Power, g16)
        finally:
            oracle.lib().tbo_set_mandelbulb_variant(0)


class FakeVolumeInfo:
    def __init__(self, dims, lo, hi, world_dims):
        self.Dimensions, self.MinValue, self.MaxValue, self.bIsNormalized, self.WorldDimensions = tuple(dims), lo, hi, True, world_dims


class FakeMHDLoader:
    @staticmethod
    def CreateVolumeFromFile(FileName, bNormalize=True, bConvertToFloat=True, bLightVolume32Bit=False, **_):
        info = _REAL_LOADER.ParseVolumeInfoFromHeaderText(open(FileName).read())  # the product's host parser (no GPU involved)
        raw = np.frombuffer((pathlib.Path(FileName).parent / info.DataFileName).read_bytes(), dtype=RU.VOXEL_DTYPES[info.OriginalFormat])
        n, lo, hi = oracle.normalize_array(info.OriginalFormat, raw.reshape(info.Dimensions[::-1]))
        r = FakeResources(info.Dimensions, 0 if n.dtype == np.uint8 else 1, bLightVolume32Bit, False)
        r.data = n
        return r, FakeVolumeInfo(info.Dimensions, lo, hi, info.WorldDimensions)


class FakeVolumeTextureToolkit:
    @staticmethod
    def LoadRawIntoNewVolume(RawFileName, Dimensions, dtype, CompressedByteSize=0, bLightVolume32Bit=False, **_):
        b = pathlib.Path(RawFileName).read_bytes()
        raw = np.frombuffer(zlib.decompress(b) if CompressedByteSize else b, dtype=dtype).reshape(tuple(Dimensions)[::-1])
        fmt = {np.dtype(v): k for k, v in RU.VOXEL_DTYPES.items()}[np.dtype(dtype)]
        n, lo, hi = oracle.normalize_array(fmt, raw)
        r = FakeResources(Dimensions, 0 if n.dtype == np.uint8 else 1, bLightVolume32Bit, False)
        r.data = n
        return r, FakeVolumeInfo(Dimensions, lo, hi, tuple(Dimensions))
