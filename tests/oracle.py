"""ctypes binding of the CPU oracle (oracle/libtbrm_oracle.so). Test infrastructure: only tests/, smoke() and the
cpu_baseline / --impl reference legs of bench.py import this."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from tbraymarcherplugin_b200 import _capi
from tbraymarcherplugin_b200.raymarch_utils import (FCamera, FDirLightParameters, FMandelbulbParameters, FRaymarchWorldParameters,
                                                    FWindowingParameters)

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "oracle" / "libtbrm_oracle.so"

ADDR_CLAMP, ADDR_WRAP, ADDR_BORDER = 0, 1, 2


class Pass(C.Structure):
    _fields_ = [
        ("face", C.c_int32), ("axis", C.c_int32), ("dirn", C.c_int32), ("td", C.c_int32 * 3), ("start", C.c_int32), ("stop", C.c_int32),
        ("weight", C.c_float), ("light_alpha", C.c_float), ("border", C.c_float), ("uv_offset", C.c_float * 2),
        ("uvw_offset", C.c_float * 3), ("step_size", C.c_float),
    ]


class LightPlan(C.Structure):
    _fields_ = [
        ("zero_direction", C.c_int32), ("add_passes", C.c_int32), ("passes", Pass * 2), ("clip_center", C.c_float * 3),
        ("clip_dir", C.c_float * 3), ("data_border", C.c_float), ("local_dir", C.c_double * 3),
    ]


class Volume(C.Structure):
    _fields_ = [
        ("data", C.c_void_p), ("ddims", C.c_int32 * 3), ("data_fmt", C.c_int32), ("light", C.c_void_p), ("ldims", C.c_int32 * 3),
        ("light_fmt", C.c_int32), ("tf", C.POINTER(C.c_float)), ("win", _capi.Windowing), ("border_exact", C.c_int32),
        ("data_addr_wrap", C.c_int32),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True)
        L = C.CDLL(str(LIB))
        L.tbo_det_pow.restype = C.c_float
        L.tbo_det_pow.argtypes = [C.c_float, C.c_float]
        L.tbo_round_to_half.restype = C.c_float
        L.tbo_round_to_half.argtypes = [C.c_float]
        L.tbo_sample_data.restype = C.c_float
        L.tbo_sample_data.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float]
        L.tbo_pcg16_x.restype = C.c_uint32
        L.tbo_sample_windowed_tf.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(_capi.Windowing), C.POINTER(C.c_float)]
        L.tbo_raymarch_lit.argtypes = [C.POINTER(Volume), C.POINTER(_capi.Camera), C.POINTER(_capi.World), C.c_float, C.c_int, C.c_int,
                                       C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]
        L.tbo_add_dir_light.argtypes = [C.POINTER(Volume), C.POINTER(_capi.DirLight), C.c_int, C.POINTER(_capi.World), C.c_void_p]
        L.tbo_change_dir_light.argtypes = [C.POINTER(Volume), C.POINTER(_capi.DirLight), C.POINTER(_capi.DirLight), C.POINTER(_capi.World), C.c_void_p]
        L.tbo_clear_light_volume.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_float]
        L.tbo_raymarch_cube_setup.argtypes = [C.POINTER(_capi.Camera), C.POINTER(_capi.World), C.c_void_p]
        L.tbo_mandelbulb_march.argtypes = [C.POINTER(_capi.Mandelbulb), C.POINTER(_capi.Camera), C.POINTER(_capi.World), C.c_int, C.c_int,
                                           C.c_void_p, C.POINTER(C.c_uint64)]
        L.tbo_plan_dir_light.argtypes = [C.POINTER(C.c_int32), C.POINTER(_capi.Windowing), C.c_int, C.POINTER(_capi.DirLight),
                                         C.POINTER(_capi.World), C.POINTER(LightPlan)]
        L.tbo_prepare_tf.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.tbo_default_tf.argtypes = [C.POINTER(C.c_float)]
        _lib = L
    return _lib


_FMT = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2}


def prepare_tf(curve_rgba: np.ndarray, height: int = 16) -> np.ndarray:
    tex = np.ascontiguousarray(np.broadcast_to(np.asarray(curve_rgba, np.float32), (height, 256, 4)))
    out = np.empty((256, 4), np.float32)
    lib().tbo_prepare_tf(tex.ctypes.data_as(C.POINTER(C.c_float)), 256, height, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def default_tf() -> np.ndarray:
    out = np.empty((256, 4), np.float32)
    lib().tbo_default_tf(out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


class OracleVolume:
    """Host-side twin of FBasicRaymarchRenderingResources for the oracle."""

    def __init__(self, data: np.ndarray, tf: np.ndarray, windowing: FWindowingParameters = None, light32: bool = True,
                 half_res: bool = False, border_exact: bool = False, data_addr_wrap: bool = False):
        self.data = np.ascontiguousarray(data)
        Z, Y, X = self.data.shape
        self.ddims = (X, Y, Z)
        self.ldims = tuple((d + 1) // 2 for d in self.ddims) if half_res else self.ddims
        self.light = np.zeros(self.ldims[::-1], np.float32 if light32 else np.uint8)
        self.tf = np.ascontiguousarray(tf, np.float32)
        self.windowing = windowing or FWindowingParameters()
        self.border_exact = border_exact
        self.data_addr_wrap = data_addr_wrap

    def c(self) -> Volume:
        return Volume(self.data.ctypes.data, (C.c_int32 * 3)(*self.ddims), _FMT[self.data.dtype], self.light.ctypes.data,
                      (C.c_int32 * 3)(*self.ldims), _FMT[self.light.dtype], self.tf.ctypes.data_as(C.POINTER(C.c_float)),
                      self.windowing.to_c(), int(self.border_exact), int(self.data_addr_wrap))

    def clear(self, value: float = 0.0):
        lib().tbo_clear_light_volume(self.light.ctypes.data, (C.c_int32 * 3)(*self.ldims), _FMT[self.light.dtype], value)

    def add_dir_light(self, light: FDirLightParameters, added: bool, world: FRaymarchWorldParameters, near_gate: np.ndarray = None) -> int:
        v, l, w = self.c(), light.to_c(), world.to_c()
        return lib().tbo_add_dir_light(C.byref(v), C.byref(l), int(added), C.byref(w), near_gate.ctypes.data if near_gate is not None else None)

    def change_dir_light(self, old: FDirLightParameters, new: FDirLightParameters, world: FRaymarchWorldParameters, near_gate=None) -> int:
        v, o, n, w = self.c(), old.to_c(), new.to_c(), world.to_c()
        return lib().tbo_change_dir_light(C.byref(v), C.byref(o), C.byref(n), C.byref(w), near_gate.ctypes.data if near_gate is not None else None)

    def raymarch_lit(self, cam: FCamera, world: FRaymarchWorldParameters, steps: float, rows=None, want_gate: bool = False):
        r0, r1 = rows if rows else (0, cam.Height)
        out = np.empty((r1 - r0, cam.Width, 4), np.float32)
        gate = np.zeros((r1 - r0, cam.Width), np.uint8) if want_gate else None
        n = C.c_uint64(0)
        v, c, w = self.c(), cam.to_c(), world.to_c()
        lib().tbo_raymarch_lit(C.byref(v), C.byref(c), C.byref(w), float(steps), r0, r1, out.ctypes.data, C.byref(n),
                               gate.ctypes.data if gate is not None else None)
        return (out, int(n.value), gate) if want_gate else (out, int(n.value))


def plan_dir_light(ldims, windowing: FWindowingParameters, light: FDirLightParameters, world: FRaymarchWorldParameters,
                   border_exact: bool = False) -> LightPlan:
    out = LightPlan()
    w, l, wo = windowing.to_c(), light.to_c(), world.to_c()
    lib().tbo_plan_dir_light((C.c_int32 * 3)(*ldims), C.byref(w), int(border_exact), C.byref(l), C.byref(wo), C.byref(out))
    return out


def cube_setup(cam: FCamera, world: FRaymarchWorldParameters) -> np.ndarray:
    out = np.empty((cam.Height, cam.Width, 4), np.float32)
    c, w = cam.to_c(), world.to_c()
    lib().tbo_raymarch_cube_setup(C.byref(c), C.byref(w), out.ctypes.data)
    return out


def mandelbulb(params: FMandelbulbParameters, cam: FCamera, world: FRaymarchWorldParameters, rows=None):
    r0, r1 = rows if rows else (0, cam.Height)
    out = np.empty((r1 - r0, cam.Width, 2), np.float32)
    n = C.c_uint64(0)
    m, c, w = params.to_c(), cam.to_c(), world.to_c()
    lib().tbo_mandelbulb_march(C.byref(m), C.byref(c), C.byref(w), r0, r1, out.ctypes.data, C.byref(n))
    return out, int(n.value)


def synth_volume(kind: str, dims, seed: int = 0x5EED1234) -> np.ndarray:
    """Synthetic volume generated by the oracle library (OpenMP); bit-identical to synth.py's numpy generators."""
    out = np.empty(tuple(dims)[::-1], np.uint8)
    lib().tbo_synth_volume_u8(0 if kind == "sphere" else 1, (C.c_int32 * 3)(*dims), C.c_uint32(seed & 0xFFFFFFFF), out.ctypes.data_as(C.c_void_p))
    return out


def det_pow(x: float, y: float) -> float:
    return float(lib().tbo_det_pow(x, y))


def sample_data(data: np.ndarray, u, v, w, mode=ADDR_CLAMP, border=0.0) -> float:
    d = np.ascontiguousarray(data)
    Z, Y, X = d.shape
    return float(lib().tbo_sample_data(d.ctypes.data, (C.c_int32 * 3)(X, Y, Z), _FMT[d.dtype], u, v, w, mode, border))


def sample_windowed_tf(value, step, tf, windowing: FWindowingParameters) -> np.ndarray:
    out = np.empty(4, np.float32)
    t = np.ascontiguousarray(tf, np.float32)
    w = windowing.to_c()
    lib().tbo_sample_windowed_tf(value, step, t.ctypes.data_as(C.POINTER(C.c_float)), C.byref(w), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


# ---- SURVEY.md §8(f) rows 2-4 -------------------------------------------------------------------------------------------------------
def _octree_args(mips):
    return (C.c_void_p * 4)(*[m.ctypes.data for m in mips]), (C.c_int32 * 3)(*mips[0].shape[::-1])


def octree_dims(ddims):
    """FMath::RoundUpToPowerOfTwo per side (RaymarchVolume.cpp:876-877)."""
    return tuple(1 << max(0, int(d) - 1).bit_length() for d in ddims)


def generate_octree(data: np.ndarray):
    """GenerateOctreeShader.usf: returns the 4 UNORM16 mips (z, y, x ordered arrays)."""
    d = np.ascontiguousarray(data)
    Z, Y, X = d.shape
    od = octree_dims((X, Y, Z))
    mips = [np.zeros(tuple(max(1, s >> m) for s in od[::-1]), np.uint16) for m in range(4)]
    ptrs, odims = _octree_args(mips)
    f = lib().tbo_generate_octree
    f.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_void_p]
    f(d.ctypes.data, (C.c_int32 * 3)(X, Y, Z), _FMT[d.dtype], odims, ptrs)
    return mips


def raymarch_intensity(vol: OracleVolume, cam: FCamera, world: FRaymarchWorldParameters, steps: float, rows=None):
    r0, r1 = rows if rows else (0, cam.Height)
    out = np.empty((r1 - r0, cam.Width, 4), np.float32)
    n = C.c_uint64(0)
    v, c, w = vol.c(), cam.to_c(), world.to_c()
    f = lib().tbo_raymarch_intensity
    f.argtypes = [C.POINTER(Volume), C.POINTER(_capi.Camera), C.POINTER(_capi.World), C.c_float, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
    f(C.byref(v), C.byref(c), C.byref(w), float(steps), r0, r1, out.ctypes.data, C.byref(n))
    return out, int(n.value)


def raymarch_octree(vol: OracleVolume, cam: FCamera, world: FRaymarchWorldParameters, steps: float, mips, octree_mip: int = 0, rows=None):
    r0, r1 = rows if rows else (0, cam.Height)
    out = np.empty((r1 - r0, cam.Width, 4), np.float32)
    n = C.c_uint64(0)
    v, c, w = vol.c(), cam.to_c(), world.to_c()
    ptrs, odims = _octree_args(mips)
    f = lib().tbo_raymarch_octree
    f.argtypes = [C.POINTER(Volume), C.POINTER(_capi.Camera), C.POINTER(_capi.World), C.c_float, C.c_int, C.c_int, C.c_void_p,
                  C.POINTER(C.c_int32), C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
    f(C.byref(v), C.byref(c), C.byref(w), float(steps), r0, r1, ptrs, odims, int(octree_mip), out.ctypes.data, C.byref(n))
    return out, int(n.value)


def mandelbulb_normal(params: FMandelbulbParameters, derivation_distance: float, cam: FCamera, world: FRaymarchWorldParameters, rows=None):
    r0, r1 = rows if rows else (0, cam.Height)
    out = np.empty((r1 - r0, cam.Width, 4), np.float32)
    n = C.c_uint64(0)
    m, c, w = params.to_c(), cam.to_c(), world.to_c()
    f = lib().tbo_mandelbulb_march_normal
    f.argtypes = [C.POINTER(_capi.Mandelbulb), C.c_float, C.POINTER(_capi.Camera), C.POINTER(_capi.World), C.c_int, C.c_int, C.c_void_p,
                  C.POINTER(C.c_uint64)]
    f(C.byref(m), float(derivation_distance), C.byref(c), C.byref(w), r0, r1, out.ctypes.data, C.byref(n))
    return out, int(n.value)


def mandelbulb_sdf(dims, center=(0.0, 0.0, 0.0), extent: float = 2.0, power: float = 8.0, g16: bool = True):
    """CalculateMandelbulbSDF.usf over a dims = (X, Y, Z) volume: UNORM16 like the reference's PF_G16 texture, or raw float32."""
    out = np.zeros(tuple(dims)[::-1], np.uint16 if g16 else np.float32)
    n = C.c_uint64(0)
    f = lib().tbo_mandelbulb_sdf
    f.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
    f((C.c_int32 * 3)(*dims), (C.c_float * 3)(*center), float(extent), float(power), 1 if g16 else 2, out.ctypes.data, C.byref(n))
    return out, int(n.value)


# EVolumeVoxelFormat (VolumeInfo.h:12-27) -> numpy
VOXEL_DTYPES = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.uint32, 5: np.int32, 6: np.float32}


def normalize_array(fmt: int, arr: np.ndarray):
    a = np.ascontiguousarray(arr, VOXEL_DTYPES[fmt])
    out = np.empty(a.shape, np.uint8 if a.itemsize == 1 else np.uint16)
    lo, hi = C.c_float(), C.c_float()
    f = lib().tbo_normalize_array
    f.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    assert f(fmt, a.ctypes.data, a.size, out.ctypes.data, C.byref(lo), C.byref(hi)) == out.itemsize
    return out, lo.value, hi.value


def convert_to_float(fmt: int, arr: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(arr, VOXEL_DTYPES[fmt])
    out = np.empty(a.shape, np.float32)
    f = lib().tbo_convert_to_float
    f.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    assert f(fmt, a.ctypes.data, a.size, out.ctypes.data) == 0
    return out


def add_dir_lights_joined(vol: OracleVolume, lights, added: bool, world: FRaymarchWorldParameters) -> int:
    """CPU twin of tbrm_add_dir_lights_joined: same-face passes of several lights in one sweep. Returns the number of sweeps."""
    arr = (_capi.DirLight * len(lights))(*[l.to_c() for l in lights])
    v, w = vol.c(), world.to_c()
    f = lib().tbo_add_dir_lights_joined
    f.argtypes = [C.POINTER(Volume), C.POINTER(_capi.DirLight), C.c_int, C.c_int, C.POINTER(_capi.World)]
    return f(C.byref(v), arr, len(lights), int(added), C.byref(w))
