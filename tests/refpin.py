"""ctypes binding of oracle/_ref/libtbrm_ref.so: the reference's OWN host code (LightingShaderUtils.cpp, VolumeInfo.cpp, the
TextureUtilities.h templates) compiled from /root/reference against the engine-type shim of oracle/ue_shim (oracle/ref.mk).
Test infrastructure. The library is prebuilt in the development container (build() of __graft_entry__.py) and travels to the GPU
box as a file; where it is absent the tests that need it skip and the committed vectors of tests/golden/ref_*.npz stand in."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

import oracle
from tbraymarcherplugin_b200 import _capi

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "oracle" / "_ref" / "libtbrm_ref.so"
REFERENCE = Path("/root/reference")

# EVolumeVoxelFormat (Source/VolumeTextureToolkit/Public/VolumeAsset/VolumeInfo.h:12-27)
VOXEL_DTYPES = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.uint32, 5: np.int32, 6: np.float32}

_lib = None


def available() -> bool:
    """The reference build can be used here: it is prebuilt and loads, or /root/reference is present to build it from."""
    if not (LIB.exists() or REFERENCE.exists()):
        return False
    try:
        lib()
        return True
    except Exception:  # a prebuilt file from another machine that does not load: the committed golden vectors stand in
        return False


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            if not REFERENCE.exists():
                raise FileNotFoundError(f"{LIB} is not built and {REFERENCE} is not present")
            subprocess.run(["make", "-C", str(ROOT / "oracle"), "-f", "ref.mk"], check=True)
        L = C.CDLL(str(LIB))
        L.tbref_plan_dir_light.argtypes = [C.POINTER(C.c_int32), C.POINTER(_capi.DirLight), C.POINTER(_capi.World), C.POINTER(oracle.LightPlan)]
        L.tbref_permutation_rows.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.tbref_normalize_array.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.tbref_convert_to_float.argtypes = [C.c_int, C.c_void_p, C.c_int32, C.c_void_p]
        L.tbref_volume_info_map.restype = C.c_float
        L.tbref_volume_info_map.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        _lib = L
    return _lib


def plan_dir_light(ldims, light, world) -> oracle.LightPlan:
    out = oracle.LightPlan()
    l, w = light.to_c(), world.to_c()
    lib().tbref_plan_dir_light((C.c_int32 * 3)(*ldims), C.byref(l), C.byref(w), C.byref(out))
    return out


def permutation_rows(face: int) -> np.ndarray:
    rows = (C.c_double * 9)()
    lib().tbref_permutation_rows(face, rows)
    return np.array(rows).reshape(3, 3)


def normalize_array(fmt: int, arr: np.ndarray):
    """UVolumeTextureToolkit::NormalizeArrayByFormat: returns (normalised array, original min, original max)."""
    a = np.ascontiguousarray(arr, VOXEL_DTYPES[fmt])
    out = np.empty(a.size, np.uint8 if a.itemsize == 1 else np.uint16)
    lo, hi = C.c_float(), C.c_float()
    nb = lib().tbref_normalize_array(fmt, a.ctypes.data, a.nbytes, out.ctypes.data, C.byref(lo), C.byref(hi))
    assert nb == out.itemsize
    return out.reshape(a.shape), lo.value, hi.value


def convert_to_float(fmt: int, arr: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(arr, VOXEL_DTYPES[fmt])
    out = np.empty(a.size, np.float32)
    assert lib().tbref_convert_to_float(fmt, a.ctypes.data, a.size, out.ctypes.data) == 0
    return out.reshape(a.shape)


def volume_info_map(what: int, is_normalized: bool, lo: float, hi: float, v: float) -> float:
    return float(lib().tbref_volume_info_map(what, int(is_normalized), lo, hi, v))


PLAN_PASS_FIELDS = ["face", "axis", "dirn", "td", "start", "stop", "weight", "light_alpha", "border", "uv_offset", "uvw_offset", "step_size"]


def plan_to_vector(p: oracle.LightPlan) -> np.ndarray:
    """Every field the reference computes, as float64 (exact for the int32 / float32 / float64 members)."""
    v = [p.zero_direction, p.add_passes, *p.clip_center, *p.clip_dir, *p.local_dir]
    for ps in p.passes:
        for f in PLAN_PASS_FIELDS:
            x = getattr(ps, f)
            v.extend(list(x) if hasattr(x, "__len__") else [x])
    return np.array(v, np.float64)


# ---- the reference's shaders run on the CPU (oracle/ref_shaders.cpp) ----------------------------------------------------------
class CameraUniforms(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("fwd", C.c_float * 3), ("rt", C.c_float * 3), ("ut", C.c_float * 3), ("inv_w2", C.c_float),
                ("inv_h2", C.c_float), ("m", C.c_float * 12), ("depth", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
                ("frame_mod8", C.c_int32), ("jitter", C.c_int32)]


def camera_uniforms(cam, world) -> CameraUniforms:
    """The fp32 uniforms of the stand-in camera, computed by the oracle library (an INPUT of the pixel shaders, not part of them)."""
    out = CameraUniforms()
    c, w = cam.to_c(), world.to_c()
    f = oracle.lib().tbo_make_camera_uniforms
    f.argtypes = [C.POINTER(_capi.Camera), C.POINTER(_capi.World), C.POINTER(CameraUniforms)]
    f.restype = None
    f(C.byref(c), C.byref(w), C.byref(out))
    return out


class RefVolume(oracle.OracleVolume):
    """OracleVolume whose ops run the reference's own shaders (AddDirLightShader.usf, ChangeDirLightShader.usf,
    WindowedRaymarchMaterials.usf ... compiled for the CPU) driven by the reference's own host math."""

    def add_dir_light(self, light, added, world, near_gate=None) -> int:
        v, l, w = self.c(), light.to_c(), world.to_c()
        f = lib().tbref_add_dir_light
        f.argtypes = [C.POINTER(oracle.Volume), C.POINTER(_capi.DirLight), C.c_int, C.POINTER(_capi.World)]
        return f(C.byref(v), C.byref(l), int(added), C.byref(w))

    def change_dir_light(self, old, new, world, near_gate=None) -> int:
        v, o, n, w = self.c(), old.to_c(), new.to_c(), world.to_c()
        f = lib().tbref_change_dir_light
        f.argtypes = [C.POINTER(oracle.Volume), C.POINTER(_capi.DirLight), C.POINTER(_capi.DirLight), C.POINTER(_capi.World)]
        return f(C.byref(v), C.byref(o), C.byref(n), C.byref(w))

    def raymarch(self, material: int, cam, world, steps: float, rows=None, octree=None, octree_mip: int = 0) -> np.ndarray:
        """material: -1 cube setup, 0 lit, 1 intensity, 2 octree (octree = list of 4 uint16 mip arrays, z-y-x ordered)."""
        r0, r1 = rows if rows else (0, cam.Height)
        out = np.empty((r1 - r0, cam.Width, 4), np.float32)
        v, w, cu = self.c(), world.to_c(), camera_uniforms(cam, world)
        mips = (C.c_void_p * 4)(*[m.ctypes.data for m in octree]) if octree else None
        odims = (C.c_int32 * 3)(*octree[0].shape[::-1]) if octree else None
        f = lib().tbref_raymarch
        f.argtypes = [C.c_int, C.POINTER(oracle.Volume), C.POINTER(CameraUniforms), C.POINTER(_capi.World), C.c_float, C.c_int, C.c_int,
                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        f(material, C.byref(v), C.byref(cu), C.byref(w), float(steps), r0, r1, mips, odims, int(octree_mip), out.ctypes.data)
        return out


def generate_octree(data: np.ndarray):
    """The reference's GenerateOctreeShader.usf: returns the 4 UNORM16 mips."""
    d = np.ascontiguousarray(data)
    Z, Y, X = d.shape
    od = oracle.octree_dims((X, Y, Z))
    mips = [np.zeros(tuple(max(1, s >> m) for s in od[::-1]), np.uint16) for m in range(4)]
    f = lib().tbref_generate_octree
    f.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_void_p]
    f(d.ctypes.data, (C.c_int32 * 3)(X, Y, Z), oracle._FMT[d.dtype], (C.c_int32 * 3)(*od), (C.c_void_p * 4)(*[m.ctypes.data for m in mips]))
    return mips


def mandelbulb_march(variant: int, params, cam, world, derivation_distance: float = 0.0, rows=None) -> np.ndarray:
    """variant 0: PerformMandelbulbRaymarchReturnDistance (2 floats / pixel), 1: ...ReturnNormal (4 floats / pixel)."""
    r0, r1 = rows if rows else (0, cam.Height)
    out = np.empty((r1 - r0, cam.Width, 2 if variant == 0 else 4), np.float32)
    m, cu = params.to_c(), camera_uniforms(cam, world)
    f = lib().tbref_mandelbulb_march
    f.argtypes = [C.c_int, C.POINTER(_capi.Mandelbulb), C.c_float, C.POINTER(CameraUniforms), C.c_int, C.c_int, C.c_void_p]
    f(variant, C.byref(m), float(derivation_distance), C.byref(cu), r0, r1, out.ctypes.data)
    return out


def mandelbulb_sdf(dims, center=(0.0, 0.0, 0.0), extent: float = 2.0, power: float = 8.0, g16: bool = True) -> np.ndarray:
    out = np.zeros(tuple(dims)[::-1], np.uint16 if g16 else np.float32)
    f = lib().tbref_mandelbulb_sdf
    f.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int, C.c_void_p]
    f((C.c_int32 * 3)(*dims), (C.c_float * 3)(*center), float(extent), float(power), 1 if g16 else 2, out.ctypes.data)
    return out


# ---- the reference's own volume loaders (oracle/ref_loaders.cpp: MHDLoader.cpp + VolumeLoader.cpp compiled from /root/reference) -----------------
def mhd_parse_file(path) -> _capi.VolumeInfo:
    """UMHDLoader::ParseVolumeInfoFromHeader of the reference on a header file."""
    out = _capi.VolumeInfo()
    f = lib().tbref_mhd_parse_file
    f.argtypes = [C.c_char_p, C.POINTER(_capi.VolumeInfo)]
    f(str(path).encode(), C.byref(out))
    return out


def mhd_create_volume(path, normalize: bool, convert_to_float: bool):
    """UMHDLoader::CreateVolumeFromFile of the reference: (FVolumeInfo of the asset, texture format as tbrm_format or -1, bulk data bytes),
    or None when the reference produces no asset."""
    info, fmt, nbytes = _capi.VolumeInfo(), C.c_int(), C.c_uint64()
    buf = np.empty(1 << 24, np.uint8)
    f = lib().tbref_mhd_create_volume
    f.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(_capi.VolumeInfo), C.POINTER(C.c_int), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    if f(str(path).encode(), int(normalize), int(convert_to_float), C.byref(info), C.byref(fmt), buf.ctypes.data, buf.nbytes, C.byref(nbytes)) != 0:
        return None
    assert nbytes.value <= buf.nbytes
    return info, fmt.value, buf[: nbytes.value].copy()
