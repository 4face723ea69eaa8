"""Host side of volume ingest in the PRODUCT library — no GPU needed: the MetaImage header parser that mirrors
UMHDLoader::ParseVolumeInfoFromHeader (MHDLoader.cpp:18-181), FVolumeInfo's value / range mappings against the reference's own
VolumeInfo.cpp (tests/golden/ref_ingest.npz), and the error behaviour of the GPU entry points on a machine without a device."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from tbraymarcherplugin_b200 import _capi
from tbraymarcherplugin_b200.raymarch_utils import FVolumeInfo, UMHDLoader, UVolumeTextureToolkit

GOLDEN = Path(__file__).resolve().parent / "golden"

HEADER = """ObjectType = Image
NDims = 3
BinaryData = True
BinaryDataByteOrderMSB = False
CompressedData = False
TransformMatrix = 1 0 0 0 1 0 0 0 1
Offset = 0 0 0
CenterOfRotation = 0 0 0
ElementSpacing = 0.5 0.5 1.25
DimSize = 64 48 20
AnatomicalOrientation = ???
ElementType = MET_SHORT
ElementDataFile = ct_head.raw
"""


def test_mhd_header_fields():
    info = UMHDLoader.ParseVolumeInfoFromHeaderText(HEADER)
    assert info.bParseWasSuccessful and info.Dimensions == (64, 48, 20) and info.Spacing == (0.5, 0.5, 1.25)
    assert info.WorldDimensions == (32.0, 24.0, 25.0)  # Spacing * Dimensions (MHDLoader.cpp:74)
    assert info.OriginalFormat == 3 and info.BytesPerVoxel == 2 and info.bIsSigned and not info.bIsCompressed
    assert info.DataFileName == "ct_head.raw" and not info.bIsNormalized
    assert (info.MinValue, info.MaxValue) == (-1000.0, 3000.0)  # FVolumeInfo defaults until the data has been scanned


@pytest.mark.parametrize("name,fmt,nbytes,signed", [("MET_UCHAR", 0, 1, False), ("MET_CHAR", 1, 1, True), ("MET_USHORT", 2, 2, False),
                                                    ("MET_SHORT", 3, 2, True), ("MET_UINT", 4, 4, False), ("MET_INT", 5, 4, True),
                                                    ("MET_FLOAT", 6, 4, True)])
def test_mhd_element_types(name, fmt, nbytes, signed):
    info = UMHDLoader.ParseVolumeInfoFromHeaderText(HEADER.replace("MET_SHORT", name))
    assert info.bParseWasSuccessful and (info.OriginalFormat, info.BytesPerVoxel, info.bIsSigned) == (fmt, nbytes, signed)


def test_mhd_variants_and_failures():
    # ElementSize is accepted in place of ElementSpacing; CompressedDataSize switches zlib loading on (MHDLoader.cpp:59,140-153)
    info = UMHDLoader.ParseVolumeInfoFromHeaderText(HEADER.replace("ElementSpacing", "ElementSize") + "CompressedDataSize = 12345\n")
    assert info.bParseWasSuccessful and info.bIsCompressed and info.CompressedByteSize == 12345
    for missing in ("DimSize", "ElementSpacing", "ElementType", "ElementDataFile"):
        bad = "\n".join(l for l in HEADER.splitlines() if not l.startswith(missing))
        assert not UMHDLoader.ParseVolumeInfoFromHeaderText(bad).bParseWasSuccessful, missing
    assert not UMHDLoader.ParseVolumeInfoFromHeaderText(HEADER.replace("MET_SHORT", "MET_DOUBLE")).bParseWasSuccessful
    assert not UMHDLoader.ParseVolumeInfoFromHeaderText("").bParseWasSuccessful
    # keys are matched as whole words anywhere in the file, in any order
    shuffled = "ElementDataFile = a.raw\nElementType = MET_UCHAR\nDimSize = 1 2 3\nElementSpacing = 1 1 1\n"
    assert UMHDLoader.ParseVolumeInfoFromHeaderText(shuffled).Dimensions == (1, 2, 3)
    lib = _capi.load()
    assert lib.tbrm_mhd_parse_header(None, None) == _capi.TBRM_ERR_INVALID_ARGUMENT


def test_volume_info_mappings_equal_the_reference():
    g = np.load(GOLDEN / "ref_ingest.npz")  # FVolumeInfo::{Normalize,Denormalize}{Value,Range} of the reference on [-1000, 3000]
    info = FVolumeInfo()
    info.c.min_value, info.c.max_value = -1000.0, 3000.0
    for normalized, key in ((1, "info_maps"), (0, "info_maps_raw")):
        info.c.is_normalized = normalized
        fns = (info.NormalizeValue, info.DenormalizeValue, info.NormalizeRange, info.DenormalizeRange)
        got = np.array([[f(float(v)) for v in g["info_values"]] for f in fns], np.float32)
        assert np.array_equal(got, g[key]), key


def test_gpu_entry_points_validate_and_fail_loudly_without_a_device():
    lib = _capi.load()
    a = np.arange(16, dtype=np.int16)
    out = np.empty(16, np.uint16)
    lo, hi = C.c_float(), C.c_float()
    args = (a.ctypes.data_as(C.c_void_p), 0, 16, out.ctypes.data_as(C.c_void_p), 0, C.byref(lo), C.byref(hi))
    assert lib.tbrm_normalize_volume(0, 99, *args) == _capi.TBRM_ERR_INVALID_ARGUMENT  # unknown voxel format
    assert lib.tbrm_normalize_volume(0, 3, None, 0, 16, out.ctypes.data_as(C.c_void_p), 0, None, None) == _capi.TBRM_ERR_INVALID_ARGUMENT
    assert lib.tbrm_convert_volume_to_float(0, 6, a.ctypes.data_as(C.c_void_p), 0, 16, out.ctypes.data_as(C.c_void_p), 0) == _capi.TBRM_ERR_INVALID_ARGUMENT
    if lib.tbrm_device_count() > 0:
        pytest.skip("a GPU is present: the no-device behaviour cannot be observed")
    assert lib.tbrm_normalize_volume(0, 3, *args) == _capi.TBRM_ERR_NO_DEVICE  # no CPU fallback
    with pytest.raises(_capi.TbrmError):
        UVolumeTextureToolkit.NormalizeArrayByFormat(a)
    h = C.c_void_p()
    info = _capi.VolumeInfo()
    assert lib.tbrm_load_mhd_volume(0, b"/nonexistent/x.mhd", 1, 0, 0, 0, C.byref(info), C.byref(h)) == _capi.TBRM_ERR_INVALID_ARGUMENT
    assert lib.tbrm_generate_octree(None) == _capi.TBRM_ERR_NOT_INITIALIZED
    d = (C.c_int32 * 3)(8, 8, 8)
    c = (C.c_float * 3)(0, 0, 0)
    buf = np.zeros(512, np.uint16)
    assert lib.tbrm_mandelbulb_sdf(0, d, c, 2.0, 8.0, 0, buf.ctypes.data_as(C.c_void_p), 0, None) == _capi.TBRM_ERR_INVALID_ARGUMENT  # G8 output
    assert lib.tbrm_mandelbulb_sdf(0, d, c, 0.0, 8.0, 1, buf.ctypes.data_as(C.c_void_p), 0, None) == _capi.TBRM_OK  # Extent <= 0: no-op
    assert lib.tbrm_mandelbulb_sdf(0, d, c, 2.0, 8.0, 1, buf.ctypes.data_as(C.c_void_p), 0, None) == _capi.TBRM_ERR_NO_DEVICE


def _build_example(tmp_path, libdir=None, libname="tbrm"):
    """examples/mhd_to_frame.cpp against libtbrm.so — or, for tests/test_kernels_emulated_cpu.py, against the emulated build of the same ABI"""
    import subprocess

    root = Path(__file__).resolve().parents[1]
    libdir = libdir or root / "tbraymarcherplugin_b200"
    exe = tmp_path / "mhd_to_frame"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-I", str(root), str(root / "examples" / "mhd_to_frame.cpp"),
                    "-L", str(libdir), f"-l{libname}", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    return exe


def test_cpp_example_builds_against_the_c_abi_and_fails_loudly(tmp_path):
    """examples/mhd_to_frame.cpp: plain C++17 over include/tbrm.h, no CUDA headers. Without a readable header it reports the library's error."""
    import subprocess

    exe = _build_example(tmp_path)
    out = subprocess.run([str(exe), str(tmp_path / "missing.mhd")], capture_output=True, text=True)
    assert out.returncode == 1 and "cannot read" in out.stderr
    (tmp_path / "bad.mhd").write_text("DimSize = 4 4 4\nElementType = MET_UCHAR\nElementDataFile = x.raw\n")  # no ElementSpacing
    out = subprocess.run([str(exe), str(tmp_path / "bad.mhd")], capture_output=True, text=True)
    assert out.returncode == 1 and "required" in out.stderr


def test_raw_loader_validates_before_touching_a_device(tmp_path):
    lib = _capi.load()
    h, info = C.c_void_p(), _capi.VolumeInfo()
    dims = (C.c_int32 * 3)(4, 4, 4)
    args = (1, 0, 0, 0, C.byref(info), C.byref(h))
    assert lib.tbrm_load_raw_volume(0, str(tmp_path / "missing.raw").encode(), dims, 0, 0, *args) == _capi.TBRM_ERR_INVALID_ARGUMENT
    assert b"cannot open" in lib.tbrm_last_error()
    assert lib.tbrm_load_raw_volume(0, b"x.raw", dims, 42, 0, *args) == _capi.TBRM_ERR_INVALID_ARGUMENT  # unknown voxel format
    (tmp_path / "short.raw").write_bytes(b"\\0" * 10)
    assert lib.tbrm_load_raw_volume(0, str(tmp_path / "short.raw").encode(), dims, 0, 0, *args) == _capi.TBRM_ERR_INVALID_ARGUMENT
    assert b"fewer bytes" in lib.tbrm_last_error()
    # unnormalised 32-bit integers have no texture format the path samples (PF_R32_SINT in the reference: "experimental")
    (tmp_path / "i32.raw").write_bytes(b"\\0" * 256)
    assert lib.tbrm_load_raw_volume(0, str(tmp_path / "i32.raw").encode(), dims, 5, 0, 0, 0, 0, 0, C.byref(info), C.byref(h)) == _capi.TBRM_ERR_UNSUPPORTED


def test_hostile_headers_come_back_as_a_status_not_as_a_terminated_process(tmp_path):
    """Header values come from a file: a negative CompressedDataSize used to become std::vector((size_t) -1) -> std::length_error across the C
    ABI, a huge DimSize std::bad_alloc — both terminate the host process (Python or UE). They must come back as TBRM_ERR_INVALID_ARGUMENT."""
    lib = _capi.load()
    h, info = C.c_void_p(), _capi.VolumeInfo()
    args = (1, 0, 0, 0, C.byref(info), C.byref(h))
    (tmp_path / "x.raw").write_bytes(b"\0" * 64)
    base = "ElementSpacing = 1 1 1\nElementType = MET_UCHAR\nElementDataFile = x.raw\n"
    bad = {
        "negative_compressed.mhd": "DimSize = 4 4 4\nCompressedDataSize = -1\n" + base,
        "huge_compressed.mhd": "DimSize = 4 4 4\nCompressedDataSize = 4611686018427387904\n" + base,
        "huge_dims.mhd": "DimSize = 60000 60000 60000\n" + base,
        "overflowing_dims.mhd": "DimSize = 2000000000 2000000000 2000000000\n" + base,
        "unparsed_dims.mhd": "DimSize = four 4 4\n" + base,
        "zero_dims.mhd": "DimSize = 0 4 4\n" + base,
    }
    for name, text in bad.items():
        (tmp_path / name).write_text(text)
        assert lib.tbrm_load_mhd_volume(0, str(tmp_path / name).encode(), *args) == _capi.TBRM_ERR_INVALID_ARGUMENT, name
        assert not h.value
    dims = (C.c_int32 * 3)(4, 4, 4)
    assert lib.tbrm_load_raw_volume(0, str(tmp_path / "x.raw").encode(), dims, 0, -5, *args) == _capi.TBRM_ERR_INVALID_ARGUMENT
    assert lib.tbrm_load_raw_volume(0, str(tmp_path / "x.raw").encode(), dims, 0, 1 << 50, *args) == _capi.TBRM_ERR_INVALID_ARGUMENT
    zero = (C.c_int32 * 3)(4, 0, 4)
    assert lib.tbrm_load_raw_volume(0, str(tmp_path / "x.raw").encode(), zero, 0, 0, *args) == _capi.TBRM_ERR_INVALID_ARGUMENT


# ---- against the reference's own loaders (MHDLoader.cpp + VolumeLoader.cpp compiled from /root/reference, tests/golden/ref_loaders.npz) ---------
import importlib.util  # noqa: E402

import refpin  # noqa: E402

_spec = importlib.util.spec_from_file_location("make_golden_ref", GOLDEN / "make_golden_ref.py")
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)
needs_ref = pytest.mark.skipif(not refpin.available(), reason="oracle/_ref/libtbrm_ref.so not built and /root/reference absent")


def test_header_parser_equals_the_reference_parser_on_the_golden_headers():
    g = np.load(GOLDEN / "ref_loaders.npz")
    assert len(g["parsed"]) == len(mk.LOADER_HEADERS)
    for text, want, name in zip(mk.LOADER_HEADERS, g["parsed"], g["data_files"]):
        info = UMHDLoader.ParseVolumeInfoFromHeaderText(text)
        got = mk.info_to_vector(info.c)
        if want[0] == 0:  # a failed parse: the reference returns a half-filled FVolumeInfo; only the verdict is contractual
            assert not info.bParseWasSuccessful, text
            continue
        assert np.array_equal(got, want), (text, got, want)
        assert info.DataFileName == str(name)
    assert (g["parsed"][:, 0] == 0).sum() == 5 and (g["parsed"][:, 0] == 1).sum() == 6


def test_conversion_decision_equals_the_reference_table():
    """IVolumeLoader::ConvertData + FVolumeInfo::VoxelFormatToPixelFormat of the reference for every element type x (normalize, to float)."""
    g = np.load(GOLDEN / "ref_loaders.npz")["conversion_table"]
    lib = _capi.load()
    assert len(g) == 28
    for fmt, nrm, flt, tex, actual, *_ in g:
        info = UMHDLoader.ParseVolumeInfoFromHeaderText(
            f"DimSize = 4 3 2\nElementSpacing = 1 1 1\nElementType = {['MET_UCHAR', 'MET_CHAR', 'MET_USHORT', 'MET_SHORT', 'MET_UINT', 'MET_INT', 'MET_FLOAT'][fmt]}\nElementDataFile = v.raw\n")
        t, a = C.c_int(), C.c_int()
        assert lib.tbrm_converted_format(C.byref(info.c), int(nrm), int(flt), C.byref(t), C.byref(a)) == _capi.TBRM_OK
        assert (t.value, a.value) == (tex, actual), (fmt, nrm, flt)


@needs_ref
def test_loader_golden_is_what_the_reference_build_produces_today():
    want = np.load(GOLDEN / "ref_loaders.npz")
    got = mk.loaders_case()
    for k in want.files:
        assert np.array_equal(want[k], got[k]), k


@needs_ref
@pytest.mark.parametrize("compressed", [False, True])
def test_reference_loader_output_equals_the_oracle_conversions(tmp_path, compressed):
    """UMHDLoader::CreateVolumeFromFile of the reference, end to end (header, raw / zlib data file, ConvertData, texture): the texture's bulk
    data equals the oracle's normalisation / float conversion of the same voxels — what the GPU loader is held to (test_gpu_zz_materials.py)."""
    import zlib

    import oracle
    from tbraymarcherplugin_b200 import synth

    dims = (12, 10, 8)
    base = synth.perlin_ct_volume(dims)
    mets = ["MET_UCHAR", "MET_CHAR", "MET_USHORT", "MET_SHORT", "MET_UINT", "MET_INT", "MET_FLOAT"]
    for fmt, dt in refpin.VOXEL_DTYPES.items():
        raw = (base.astype(np.float32) * 0.4 - 30).astype(dt) if np.dtype(dt).kind != "u" else (base.astype(np.uint32) * (1 if fmt == 0 else 200)).astype(dt)
        payload = raw.tobytes()
        header = f"NDims = 3\nDimSize = {dims[0]} {dims[1]} {dims[2]}\nElementSpacing = 0.5 0.5 2\nElementType = {mets[fmt]}\n"
        if compressed:
            payload = zlib.compress(payload)
            header += f"CompressedData = True\nCompressedDataSize = {len(payload)}\n"
        (tmp_path / "v.bin").write_bytes(payload)
        (tmp_path / "v.mhd").write_text(header + "ElementDataFile = v.bin\n")
        for nrm, flt in ((1, 0), (0, 1), (0, 0)):
            info, tex, bulk = refpin.mhd_create_volume(tmp_path / "v.mhd", nrm, flt)
            if nrm:
                want, lo, hi = oracle.normalize_array(fmt, raw)
                assert (info.min_value, info.max_value) == (lo, hi) and info.is_normalized
            elif flt and fmt != 6:
                want = oracle.convert_to_float(fmt, raw)
            else:
                want = raw
            assert bulk.tobytes() == want.tobytes(), (fmt, nrm, flt)
            assert info.is_compressed == int(compressed)
