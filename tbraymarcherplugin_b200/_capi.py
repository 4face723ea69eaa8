"""ctypes binding of include/tbrm.h (the C ABI of libtbrm.so).

The library is the product: if it is missing or fails to load this module raises — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libtbrm.so"

TBRM_OK = 0
TBRM_ERR_INVALID_ARGUMENT = 1
TBRM_ERR_NOT_INITIALIZED = 2
TBRM_ERR_CUDA = 3
TBRM_ERR_UNSUPPORTED = 4
TBRM_ERR_NO_DEVICE = 5

FMT_G8, FMT_G16, FMT_R32F = 0, 1, 2
SYNTH_SPHERE, SYNTH_PERLIN_CT = 0, 1


class DirLight(C.Structure):
    _fields_ = [("direction", C.c_double * 3), ("intensity", C.c_float)]


class ClipPlane(C.Structure):
    _fields_ = [("center", C.c_double * 3), ("direction", C.c_double * 3)]


class World(C.Structure):
    _fields_ = [("translation", C.c_double * 3), ("rotation", C.c_double * 4), ("scale", C.c_double * 3), ("clip", ClipPlane)]


class Windowing(C.Structure):
    _fields_ = [("center", C.c_float), ("width", C.c_float), ("low_cutoff", C.c_int32), ("high_cutoff", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [
        ("eye", C.c_double * 3),
        ("look_at", C.c_double * 3),
        ("up", C.c_double * 3),
        ("hfov_deg", C.c_double),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("scene_depth", C.c_float),
        ("frame_index", C.c_int32),
        ("jitter", C.c_int32),
    ]


class Mandelbulb(C.Structure):
    _fields_ = [
        ("center", C.c_float * 3),
        ("extent", C.c_float),
        ("power", C.c_float),
        ("max_steps", C.c_float),
        ("max_iterations", C.c_float),
        ("bailout", C.c_float),
        ("high_precision_eps", C.c_float),
        ("low_precision_eps", C.c_float),
    ]


class Options(C.Structure):
    _fields_ = [("border_exact", C.c_int32), ("data_addr_wrap", C.c_int32), ("sweep_impl", C.c_int32), ("reserved", C.c_int32 * 5)]


class SweepStats(C.Structure):
    _fields_ = [
        ("passes", C.c_int32),
        ("fell_back", C.c_int32),
        ("voxels", C.c_int64),
        ("kernel_launches", C.c_int32),
        ("faces", C.c_int32 * 4),
        ("impl", C.c_int32 * 4),
    ]


class Slab(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("z_begin", C.c_int32), ("z_end", C.c_int32)]


class PassPlan(C.Structure):
    _fields_ = [
        ("face", C.c_int32),
        ("axis", C.c_int32),
        ("dirn", C.c_int32),
        ("td", C.c_int32 * 3),
        ("start", C.c_int32),
        ("stop", C.c_int32),
        ("weight", C.c_float),
        ("light_alpha", C.c_float),
        ("border", C.c_float),
        ("uv_offset", C.c_float * 2),
        ("uvw_offset", C.c_float * 3),
        ("step_size", C.c_float),
    ]


class LightPlan(C.Structure):
    _fields_ = [
        ("zero_direction", C.c_int32),
        ("add_passes", C.c_int32),
        ("passes", PassPlan * 2),
        ("clip_center", C.c_float * 3),
        ("clip_dir", C.c_float * 3),
        ("data_border", C.c_float),
        ("local_dir", C.c_double * 3),
    ]


class VolumeInfo(C.Structure):  # tbrm_volume_info = FVolumeInfo
    _fields_ = [
        ("parse_ok", C.c_int32),
        ("dims", C.c_int32 * 3),
        ("spacing", C.c_double * 3),
        ("world_dims", C.c_double * 3),
        ("original_format", C.c_int32),
        ("actual_format", C.c_int32),
        ("bytes_per_voxel", C.c_int32),
        ("is_signed", C.c_int32),
        ("is_normalized", C.c_int32),
        ("min_value", C.c_float),
        ("max_value", C.c_float),
        ("is_compressed", C.c_int32),
        ("compressed_bytes", C.c_int64),
        ("data_file", C.c_char * 512),
    ]


# every symbol include/tbrm.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_I = C.c_int
PROTOTYPES = {
    "tbrm_abi_version": (_I, []),
    "tbrm_status_string": (C.c_char_p, [_I]),
    "tbrm_last_error": (C.c_char_p, []),
    "tbrm_device_count": (_I, []),
    "tbrm_kernel_launch_count": (C.c_int64, []),
    "tbrm_create": (_I, [_I, C.POINTER(C.c_int32), _I, _I, _I, C.POINTER(_P)]),
    "tbrm_destroy": (_I, [_P]),
    "tbrm_set_options": (_I, [_P, C.POINTER(Options)]),
    "tbrm_upload_volume": (_I, [_P, _P, _I]),
    "tbrm_bind_volume_device": (_I, [_P, _P]),
    "tbrm_upload_volume_async": (_I, [_P, _P]),
    "tbrm_present_volume": (_I, [_P]),
    "tbrm_raymarch_lit_to_host_async": (_I, [_P, C.POINTER(Camera), C.POINTER(World), C.c_float, _P]),
    "tbrm_download_wait": (_I, [_P]),
    "tbrm_set_transfer_function": (_I, [_P, C.POINTER(C.c_float), _I, _I]),
    "tbrm_make_default_tf": (_I, [_P]),
    "tbrm_set_windowing": (_I, [_P, C.POINTER(Windowing)]),
    "tbrm_clear_light_volume": (_I, [_P, C.c_float]),
    "tbrm_add_dir_light": (_I, [_P, C.POINTER(DirLight), _I, C.POINTER(World), C.POINTER(_I), _I]),
    "tbrm_change_dir_light": (_I, [_P, C.POINTER(DirLight), C.POINTER(DirLight), C.POINTER(World), C.POINTER(_I), _I]),
    "tbrm_add_dir_light_stats": (_I, [_P, C.POINTER(DirLight), _I, C.POINTER(World), C.POINTER(_I), _I, C.POINTER(SweepStats)]),
    "tbrm_change_dir_light_stats": (
        _I,
        [_P, C.POINTER(DirLight), C.POINTER(DirLight), C.POINTER(World), C.POINTER(_I), _I, C.POINTER(SweepStats)],
    ),
    "tbrm_add_dir_lights_joined": (_I, [_P, C.POINTER(DirLight), _I, _I, C.POINTER(World), C.POINTER(_I), C.POINTER(SweepStats)]),
    "tbrm_plan_dir_light": (
        _I,
        [C.POINTER(C.c_int32), C.POINTER(Windowing), C.POINTER(Options), C.POINTER(DirLight), C.POINTER(World), C.POINTER(LightPlan)],
    ),
    "tbrm_light_volume_dims": (_I, [_P, C.POINTER(C.c_int32)]),
    "tbrm_download_light_volume": (_I, [_P, _P]),
    "tbrm_upload_light_volume": (_I, [_P, _P]),
    "tbrm_light_volume_device_ptr": (_P, [_P]),
    "tbrm_data_volume_device_ptr": (_P, [_P]),
    "tbrm_bind_light_volume_device": (_I, [_P, _P]),
    "tbrm_slab_partition": (None, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "tbrm_slab_configure": (_I, [_P, C.POINTER(Slab)]),
    "tbrm_slab_arena": (_I, [_P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "tbrm_slab_ipc_handle": (_I, [_P, _P]),
    "tbrm_slab_open_peer": (_I, [_P, _I, _P]),
    "tbrm_slab_set_peer": (_I, [_P, _I, _P]),
    "tbrm_slab_reset_comm": (_I, [_P]),
    "tbrm_slab_check": (_I, [_P]),
    "tbrm_slab_set_timeout_ms": (_I, [_P, _I]),
    "tbrm_add_dir_light_pass": (_I, [_P, C.POINTER(DirLight), _I, C.POINTER(World), _I, _I, C.POINTER(SweepStats)]),
    "tbrm_slab_pass_order": (_I, [_P, C.POINTER(DirLight), C.POINTER(World), _I, C.POINTER(_I)]),
    "tbrm_raymarch_cube_setup": (_I, [_P, C.POINTER(Camera), C.POINTER(World), _P, _I]),
    "tbrm_raymarch_lit": (_I, [_P, C.POINTER(Camera), C.POINTER(World), C.c_float, _I, _I, _P, _I, C.POINTER(C.c_uint64)]),
    "tbrm_raymarch_lit_interleaved": (_I, [_P, C.POINTER(Camera), C.POINTER(World), C.c_float, _I, _I, _I, _P, _I, C.POINTER(C.c_uint64)]),
    "tbrm_raymarch_interleaved_rows": (_I, [_I, _I, _I, _I]),
    "tbrm_mandelbulb_march": (_I, [_I, C.POINTER(Mandelbulb), C.POINTER(Camera), C.POINTER(World), _I, _I, _P, _I, C.POINTER(C.c_uint64)]),
    "tbrm_generate_octree": (_I, [_P]),
    "tbrm_octree_mip_dims": (_I, [_P, _I, C.POINTER(C.c_int32)]),
    "tbrm_download_octree_mip": (_I, [_P, _I, _P]),
    "tbrm_raymarch_intensity": (_I, [_P, C.POINTER(Camera), C.POINTER(World), C.c_float, _I, _I, _P, _I, C.POINTER(C.c_uint64)]),
    "tbrm_raymarch_octree": (_I, [_P, C.POINTER(Camera), C.POINTER(World), C.c_float, _I, _I, _I, _P, _I, C.POINTER(C.c_uint64)]),
    "tbrm_mandelbulb_march_normal": (
        _I,
        [_I, C.POINTER(Mandelbulb), C.c_float, C.POINTER(Camera), C.POINTER(World), _I, _I, _P, _I, C.POINTER(C.c_uint64)],
    ),
    "tbrm_debug_mandelbulb_sdf_p8": (C.c_float, [C.POINTER(C.c_float), C.c_float, _I, C.POINTER(C.c_uint32)]),
    "tbrm_debug_download_derived": (_I, [_P, _I, _P, C.c_size_t]),
    "tbrm_mandelbulb_sdf": (_I, [_I, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_float, C.c_float, _I, _P, _I, C.POINTER(C.c_uint64)]),
    "tbrm_mhd_parse_header": (_I, [C.c_char_p, C.POINTER(VolumeInfo)]),
    "tbrm_volume_info_normalize_value": (C.c_float, [C.POINTER(VolumeInfo), C.c_float]),
    "tbrm_volume_info_denormalize_value": (C.c_float, [C.POINTER(VolumeInfo), C.c_float]),
    "tbrm_volume_info_normalize_range": (C.c_float, [C.POINTER(VolumeInfo), C.c_float]),
    "tbrm_volume_info_denormalize_range": (C.c_float, [C.POINTER(VolumeInfo), C.c_float]),
    "tbrm_converted_format": (_I, [C.POINTER(VolumeInfo), _I, _I, C.POINTER(_I), C.POINTER(_I)]),
    "tbrm_normalize_volume": (_I, [_I, _I, _P, _I, C.c_uint64, _P, _I, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "tbrm_convert_volume_to_float": (_I, [_I, _I, _P, _I, C.c_uint64, _P, _I]),
    "tbrm_load_mhd_volume": (_I, [_I, C.c_char_p, _I, _I, _I, _I, C.POINTER(VolumeInfo), C.POINTER(_P)]),
    "tbrm_load_raw_volume": (_I, [_I, C.c_char_p, C.POINTER(C.c_int32), _I, C.c_int64, _I, _I, _I, _I, C.POINTER(VolumeInfo), C.POINTER(_P)]),
    "tbrm_flush": (_I, [_P]),
    "tbrm_stream": (_P, [_P]),
    "tbrm_set_stream": (_I, [_P, _P]),
    "tbrm_slab_light_ipc_handle": (_I, [_P, _P]),
    "tbrm_slab_open_peer_light": (_I, [_P, C.c_int, _P]),
    "tbrm_slab_set_peer_light": (_I, [_P, C.c_int, _P]),
    "tbrm_slab_push_light": (_I, [_P, C.c_int]),
    "tbrm_timer_begin": (_I, [_P]),
    "tbrm_timer_end": (_I, [_P, C.POINTER(C.c_float)]),
    "tbrm_synth_volume_u8": (_I, [_I, _I, C.POINTER(C.c_int32), C.c_uint32, _P, _I]),
}

_lib = None


class TbrmError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"tbrm status {status}: {message}")
        self.status = status


def load() -> C.CDLL:
    """Load libtbrm.so; raises if it has not been built (python -m tbraymarcherplugin_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -m tbraymarcherplugin_b200.build). "
            "There is no CPU fallback for the hot path."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != TBRM_OK:
        lib = load()
        msg = lib.tbrm_last_error().decode() or lib.tbrm_status_string(status).decode()
        raise TbrmError(status, msg)
