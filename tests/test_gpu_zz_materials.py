"""GPU parity of the rows next to the hot path, through the C ABI, against the CPU oracle AND against the
vectors produced by the reference's own code (tests/golden/ref_materials.npz, ref_ingest.npz — see tests/test_ref_materials_cpu.py):
octree generation, intensity march, octree march, volume normalisation / float conversion, the MHD loader (bit-exact: integer / byte work
and the shared fp32 contract), and the Mandelbulb variants (CUDA vs libm transcendentals: tolerance + a small budget, like the distance
march in test_gpu_parity.py)."""
import importlib.util
import zlib
from pathlib import Path

import numpy as np
import pytest

import oracle
from tbraymarcherplugin_b200 import FMT_G8, synth
from tbraymarcherplugin_b200.raymarch_utils import (FWindowingParameters, UMHDLoader, URaymarchUtils, UVolumeTextureToolkit, VOXEL_DTYPES)

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
_spec = importlib.util.spec_from_file_location("make_golden_ref", GOLDEN / "make_golden_ref.py")
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)


def make_res(data, windowing):
    Z, Y, X = data.shape
    fmt = FMT_G8 if data.dtype == np.uint8 else (1 if data.dtype == np.uint16 else 2)
    res = URaymarchUtils.InitializeRaymarchResources((X, Y, Z), fmt, bLightVolume32Bit=True)
    URaymarchUtils.SetDataVolume(res, data)
    URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(res, windowing)
    return res


@pytest.mark.parametrize("dims", mk.MATERIAL_DIMS)
def test_octree_and_materials_equal_the_reference_golden_outputs(dims):
    want = np.load(GOLDEN / "ref_materials.npz")
    tag = "x".join(map(str, dims))
    data = synth.perlin_ct_volume(dims)
    for wname, wv in mk.MATERIAL_WINDOWS.items():
        res = make_res(data, FWindowingParameters(*wv))
        URaymarchUtils.GenerateOctree(res)
        for m in range(4):
            got = URaymarchUtils.ReadOctreeMip(res, m)
            assert got.shape == want[f"octree_{tag}_mip{m}"].shape and np.array_equal(got, want[f"octree_{tag}_mip{m}"]), f"mip {m}"
        for world_name, mkw in mk.MATERIAL_WORLDS.items():
            rgba, steps = URaymarchUtils.PerformWindowedIntensityRaymarch(res, mk.material_camera(), mkw(), 40.0)
            assert np.array_equal(rgba, want[f"intensity_{tag}_{wname}_{world_name}"]) and steps > 0
            for mip in (0, 2):
                rgba, _ = URaymarchUtils.PerformWindowedRaymarchOctree(res, mk.material_camera(), mkw(), 40.0, mip)
                assert np.array_equal(rgba, want[f"octree_march_{tag}_{wname}_{world_name}_mip{mip}"]), (wname, world_name, mip)


@pytest.mark.parametrize("dims,dtype", [((64, 48, 80), np.uint8), ((33, 17, 70), np.uint8), ((16, 24, 8), np.uint16), ((40, 40, 40), np.float32),
                                        ((144, 80, 96), np.uint8)])
def test_octree_and_materials_match_oracle(dims, dtype):
    base = synth.perlin_ct_volume(dims)
    data = base if dtype == np.uint8 else (base.astype(np.uint16) * 257 if dtype == np.uint16 else (base / np.float32(200)).astype(np.float32))
    win = FWindowingParameters(0.45, 0.5, True, False)
    res = make_res(data, win)
    URaymarchUtils.GenerateOctree(res)
    mips = oracle.generate_octree(data)
    for m in range(4):
        assert np.array_equal(URaymarchUtils.ReadOctreeMip(res, m), mips[m]), f"mip {m}"
    vol = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), win)
    for world in (synth.identity_world(), synth.scaled_rotated_world(), synth.clipped_world()):
        for jitter in (False, True):
            cam = synth.benchmark_camera(96, 64, jitter=jitter, frame=5)
            rgba, steps = URaymarchUtils.PerformWindowedIntensityRaymarch(res, cam, world, 72.0)
            ref, ref_steps = oracle.raymarch_intensity(vol, cam, world, 72.0)
            assert steps == ref_steps and np.array_equal(rgba, ref)
            for mip in range(4):
                rgba, steps = URaymarchUtils.PerformWindowedRaymarchOctree(res, cam, world, 72.0, mip)
                ref, ref_steps = oracle.raymarch_octree(vol, cam, world, 72.0, mips, mip)
                assert steps == ref_steps and np.array_equal(rgba, ref), mip
    # rows: a sub-range renders the same pixels
    cam = synth.benchmark_camera(96, 64, jitter=True)
    full, _ = URaymarchUtils.PerformWindowedRaymarchOctree(res, cam, synth.identity_world(), 72.0, 1)
    part, _ = URaymarchUtils.PerformWindowedRaymarchOctree(res, cam, synth.identity_world(), 72.0, 1, rows=(13, 41))
    assert np.array_equal(part, full[13:41])


def test_octree_is_invalidated_by_a_new_data_volume():
    data = synth.perlin_ct_volume((32, 32, 32))
    res = make_res(data, FWindowingParameters())
    cam = synth.benchmark_camera(32, 32)
    from tbraymarcherplugin_b200 import TbrmError
    with pytest.raises(TbrmError):  # no octree yet
        URaymarchUtils.PerformWindowedRaymarchOctree(res, cam, synth.identity_world(), 32.0, 0)
    URaymarchUtils.GenerateOctree(res)
    URaymarchUtils.PerformWindowedRaymarchOctree(res, cam, synth.identity_world(), 32.0, 0)
    URaymarchUtils.SetDataVolume(res, data)  # bRequestedOctreeRebuild (RaymarchVolume.cpp:553-554)
    with pytest.raises(TbrmError):
        URaymarchUtils.PerformWindowedRaymarchOctree(res, cam, synth.identity_world(), 32.0, 0)
    with pytest.raises(TbrmError):
        URaymarchUtils.PerformWindowedRaymarchOctree(res, cam, synth.identity_world(), 32.0, 4)  # mips 0..3


# ---- volume ingest ------------------------------------------------------------------------------------------------------------------
def test_normalisation_equals_the_reference_golden_outputs():
    want = np.load(GOLDEN / "ref_ingest.npz")
    for key in (0, 1, 2, 3, 4, 5, 6, 16):
        a = want[f"in_{key}"]
        n, lo, hi = UVolumeTextureToolkit.NormalizeArrayByFormat(a)
        assert n.dtype == want[f"normalized_{key}"].dtype and np.array_equal(n, want[f"normalized_{key}"]), key
        assert np.array_equal(np.array([lo, hi], np.float32), want[f"minmax_{key}"]), key
        if key % 10 != 6:
            assert np.array_equal(UVolumeTextureToolkit.ConvertArrayToFloat(a), want[f"float_{key}"]), key


@pytest.mark.parametrize("fmt", sorted(VOXEL_DTYPES))
@pytest.mark.parametrize("count", [1, 31, 4096, 1 << 20, (1 << 22) + 13])
def test_normalisation_matches_oracle(fmt, count):
    rng = np.random.default_rng(fmt * 1000 + count % 997)
    dt = VOXEL_DTYPES[fmt]
    if fmt == 6:
        a = (rng.standard_normal(count) * 1200.0 - 300.0).astype(np.float32)
    else:
        info = np.iinfo(dt)
        a = rng.integers(max(info.min, -2000000), min(info.max, 3000000), count, endpoint=True).astype(dt)
    n, lo, hi = UVolumeTextureToolkit.NormalizeArrayByFormat(a)
    rn, rlo, rhi = oracle.normalize_array(fmt, a)
    assert (lo, hi) == (rlo, rhi) and np.array_equal(n, rn)
    if count > 1:
        assert n.max() == np.iinfo(n.dtype).max and n.min() == 0
    if fmt != 6:
        assert np.array_equal(UVolumeTextureToolkit.ConvertArrayToFloat(a), oracle.convert_to_float(fmt, a))
    # an unaligned source (scalar path of the kernels) gives the same numbers
    if count > 64:
        b = a[1:]
        n2, lo2, hi2 = UVolumeTextureToolkit.NormalizeArrayByFormat(b)
        r2 = oracle.normalize_array(fmt, b)
        assert np.array_equal(n2, r2[0]) and (lo2, hi2) == r2[1:]


def test_constant_volume_normalises_to_zero():
    n, lo, hi = UVolumeTextureToolkit.NormalizeArrayByFormat(np.full(1000, 7, np.int16))  # 0 / 0 -> NaN -> 0 (as on x86)
    assert (lo, hi) == (7.0, 7.0) and not n.any()


@pytest.mark.parametrize("met,dtype,compressed", [("MET_SHORT", np.int16, False), ("MET_UCHAR", np.uint8, True), ("MET_FLOAT", np.float32, False),
                                                  ("MET_USHORT", np.uint16, True)])
def test_mhd_file_to_resources_to_frame(tmp_path, met, dtype, compressed):
    """UMHDLoader::CreateVolumeFromFile -> InitializeRaymarchResources -> a lit frame: equals the oracle fed with the oracle's own
    normalisation of the same voxels."""
    dims = (48, 40, 32)
    base = synth.perlin_ct_volume(dims).astype(np.float32)
    raw = (base * 12.0 - 1000.0).astype(dtype) if np.dtype(dtype).kind != "u" else (base * (1 if dtype == np.uint8 else 200)).astype(dtype)
    payload = raw.tobytes()
    header = f"ObjectType = Image\nNDims = 3\nDimSize = {dims[0]} {dims[1]} {dims[2]}\nElementSpacing = 0.7 0.7 1.5\nElementType = {met}\n"
    if compressed:
        payload = zlib.compress(payload, 6)
        header += f"CompressedData = True\nCompressedDataSize = {len(payload)}\nElementDataFile = vol.zraw\n"
        (tmp_path / "vol.zraw").write_bytes(payload)
    else:
        header += "ElementDataFile = vol.raw\n"
        (tmp_path / "vol.raw").write_bytes(payload)
    (tmp_path / "vol.mhd").write_text(header)
    res, info = UMHDLoader.CreateVolumeFromFile(str(tmp_path / "vol.mhd"), bNormalize=True, bLightVolume32Bit=True)
    fmt = {np.dtype(v): k for k, v in VOXEL_DTYPES.items()}[np.dtype(dtype)]
    want, lo, hi = oracle.normalize_array(fmt, raw)
    assert info.bParseWasSuccessful and info.Dimensions == dims and info.bIsNormalized and (info.MinValue, info.MaxValue) == (lo, hi)
    assert info.bIsCompressed == compressed and res.DataDims == dims
    win = FWindowingParameters(info.NormalizeValue(lo + 0.45 * (hi - lo)), info.NormalizeRange(0.5 * (hi - lo)), True, False)
    URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(res, win)
    vol = oracle.OracleVolume(want, oracle.prepare_tf(synth.soft_ct_curve()), win)
    world = synth.identity_world()
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    for l in synth.LIGHTS[:2]:
        assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True)
        vol.add_dir_light(l, True, world)
    assert np.array_equal(URaymarchUtils.ReadLightVolume(res), vol.light)
    cam = synth.benchmark_camera(64, 48)
    rgba, steps = URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, 64.0)
    ref, ref_steps = vol.raymarch_lit(cam, world, 64.0)
    assert steps == ref_steps and np.array_equal(rgba, ref)
    res.release()
    # as stored / converted to float
    if dtype == np.int16:
        res_f, info_f = UMHDLoader.CreateVolumeFromFile(str(tmp_path / "vol.mhd"), bNormalize=False, bConvertToFloat=True, bLightVolume32Bit=True)
        assert info_f.ActualFormat == 6 and not info_f.bIsNormalized and res_f.DataFormat == 2
        res_f.release()


# ---- Mandelbulb variants (tolerance: CUDA vs libm transcendentals; the iteration is chaotic next to the surface) --------------------------
def test_mandelbulb_normal_march_close_to_oracle():
    from tbraymarcherplugin_b200.raymarch_utils import FMandelbulbParameters
    cam = synth.benchmark_camera(96, 64, jitter=False)
    mb = FMandelbulbParameters(MaxSteps=64.0, MaxIterations=8.0)
    got, iters = URaymarchUtils.PerformMandelbulbRaymarchReturnNormal(mb, 0.01, cam, synth.identity_world())
    ref, ref_iters = oracle.mandelbulb_normal(mb, 0.01, cam, synth.identity_world())
    assert (got[..., 3] != ref[..., 3]).mean() <= 0.01, "hit / miss differs on more than 1 % of the pixels"
    both = (got[..., 3] == 1) & (ref[..., 3] == 1) & (np.abs(ref[..., :3]).sum(-1) > 0) & (np.abs(got[..., :3]).sum(-1) > 0)
    assert both.sum() > 500
    cos = np.clip((got[..., :3][both] * ref[..., :3][both]).sum(-1), -1, 1)
    assert (np.arccos(cos) < 0.1).mean() >= 0.8, "normals differ"
    assert abs(iters - ref_iters) / ref_iters < 0.05
    # rows and the golden vector of the reference's own shader
    want = np.load(GOLDEN / "ref_materials.npz")["mandelbulb_normal"]
    small = synth.benchmark_camera(*mk.MANDELBULB_VIEW, jitter=False)
    g, _ = URaymarchUtils.PerformMandelbulbRaymarchReturnNormal(mk.mandelbulb_params(), 0.01, small, synth.identity_world())
    assert (g[..., 3] != want[..., 3]).mean() <= 0.03


@pytest.mark.parametrize("g16", [True, False])
def test_mandelbulb_sdf_bake_close_to_oracle(g16):
    dims, center, extent, power = (48, 40, 32), (0.1, 0.0, -0.05), 2.4, 8.0
    got, iters = URaymarchUtils.CalculateMandelbulbSDF(dims, center, extent, power, g16=g16)
    ref, ref_iters = oracle.mandelbulb_sdf(dims, center, extent, power, g16)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    d = np.abs(np.nan_to_num(got.astype(np.float64)) - np.nan_to_num(ref.astype(np.float64)))
    tol = 8 if g16 else 1e-4  # 8 LSB of UNORM16 = 1.2e-4
    assert (d > tol).mean() <= 0.005, f"{(d > tol).mean():.4f} of the voxels differ by more than {tol}"
    assert abs(iters - ref_iters) / ref_iters < 2e-3
    want = np.load(GOLDEN / "ref_materials.npz")["mandelbulb_sdf_g16" if g16 else "mandelbulb_sdf_r32f"]
    small, _ = URaymarchUtils.CalculateMandelbulbSDF(g16=g16, **{"Dimensions": mk.SDF_CASE["dims"], "Center": mk.SDF_CASE["center"],
                                                                 "Extent": mk.SDF_CASE["extent"], "Power": mk.SDF_CASE["power"]})
    d = np.abs(np.nan_to_num(small.astype(np.float64)) - np.nan_to_num(want.astype(np.float64)))
    assert (d > tol).mean() <= 0.005
    # Extent <= 0: the reference enqueues nothing
    untouched, n = URaymarchUtils.CalculateMandelbulbSDF((8, 8, 8), Extent=0.0, g16=g16)
    assert n == 0 and not untouched.any()
