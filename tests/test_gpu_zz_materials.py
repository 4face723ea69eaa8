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
    assert (got[..., 3] != ref[..., 3]).mean() <= 0.02, "hit / miss differs on more than 2 % of the pixels"
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
    assert (d > tol).mean() <= 0.02, f"{(d > tol).mean():.4f} of the voxels differ by more than {tol}"
    assert abs(iters - ref_iters) / ref_iters < 0.02
    want = np.load(GOLDEN / "ref_materials.npz")["mandelbulb_sdf_g16" if g16 else "mandelbulb_sdf_r32f"]
    small, _ = URaymarchUtils.CalculateMandelbulbSDF(g16=g16, **{"Dimensions": mk.SDF_CASE["dims"], "Center": mk.SDF_CASE["center"],
                                                                 "Extent": mk.SDF_CASE["extent"], "Power": mk.SDF_CASE["power"]})
    d = np.abs(np.nan_to_num(small.astype(np.float64)) - np.nan_to_num(want.astype(np.float64)))
    assert (d > tol).mean() <= 0.02
    # Extent <= 0: the reference enqueues nothing
    untouched, n = URaymarchUtils.CalculateMandelbulbSDF((8, 8, 8), Extent=0.0, g16=g16)
    assert n == 0 and not untouched.any()


def test_cpp_example_runs_end_to_end(tmp_path):
    """examples/mhd_to_frame.cpp (plain C++ over the C ABI): MetaImage file -> resources -> sweep -> octree -> the three materials."""
    import subprocess

    from test_ingest_cpu import _build_example

    dims = (48, 40, 32)
    raw = (synth.perlin_ct_volume(dims).astype(np.int16) * 12 - 1000)
    (tmp_path / "v.raw").write_bytes(raw.tobytes())
    (tmp_path / "v.mhd").write_text(f"NDims = 3\nDimSize = {dims[0]} {dims[1]} {dims[2]}\nElementSpacing = 1 1 1\nElementType = MET_SHORT\nElementDataFile = v.raw\n")
    out = subprocess.run([str(_build_example(tmp_path)), str(tmp_path / "v.mhd")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "48 x 40 x 32 voxels" in out.stdout and "normalised to G16" in out.stdout
    steps = [int(l.split(":")[1].split()[0]) for l in out.stdout.splitlines() if "march:" in l]
    assert len(steps) == 3 and all(s > 0 for s in steps) and steps[1] < steps[0]  # the intensity march stops at its first sample


def test_mandelbulb_power8_kernels_match_their_cpu_twin():
    """Power == 8 runs the transcendental-free iteration (mandelbulb_sdf_p8): only +, -, *, /, sqrt and one log, so the oracle's variant 1
    (the same arithmetic on the CPU) must agree far more tightly than the reference formulation does; another power takes the
    transcendental path and is compared with the usual budget."""
    from tbraymarcherplugin_b200.raymarch_utils import FMandelbulbParameters

    L = oracle.lib()
    world, cam = synth.identity_world(), synth.benchmark_camera(240, 135, jitter=False)
    mb = FMandelbulbParameters(MaxSteps=256.0, MaxIterations=16.0)
    got, iters = URaymarchUtils.PerformMandelbulbRaymarchReturnDistance(mb, cam, world)
    gsdf, _ = URaymarchUtils.CalculateMandelbulbSDF((40, 36, 32), (0.1, 0.0, -0.05), 2.4, 8.0, g16=False)
    try:
        L.tbo_set_mandelbulb_variant(1)
        twin, twin_iters = oracle.mandelbulb(mb, cam, world)
        tsdf, _ = oracle.mandelbulb_sdf((40, 36, 32), (0.1, 0.0, -0.05), 2.4, 8.0, False)
    finally:
        L.tbo_set_mandelbulb_variant(0)
    ref, _ = oracle.mandelbulb(mb, cam, world)
    bad_twin = (np.abs(got - twin).max(-1) > 1e-4).mean()
    bad_ref = (np.abs(got - ref).max(-1) > 1e-4).mean()
    assert bad_twin <= 0.01 and bad_ref <= 0.02, (bad_twin, bad_ref)  # vs the twin only the final log differs (1 ulp, amplified near the surface)
    assert abs(iters - twin_iters) / twin_iters < 5e-3
    assert (np.abs(gsdf - tsdf) > 1e-5).mean() <= 0.01
    mb6 = FMandelbulbParameters(MaxSteps=64.0, MaxIterations=8.0, Power=6.0)
    got6, _ = URaymarchUtils.PerformMandelbulbRaymarchReturnDistance(mb6, cam, world)
    ref6, _ = oracle.mandelbulb(mb6, cam, world)
    assert (np.abs(got6 - ref6).max(-1) > 1e-4).mean() <= 0.02


def test_headerless_raw_file_loads_like_the_mhd_path(tmp_path):
    from tbraymarcherplugin_b200.raymarch_utils import UVolumeTextureToolkit as T

    dims = (40, 24, 16)
    raw = (synth.perlin_ct_volume(dims).astype(np.int16) * 9 - 700)
    (tmp_path / "v.raw").write_bytes(raw.tobytes())
    (tmp_path / "v.zraw").write_bytes(zlib.compress(raw.tobytes(), 6))
    want, lo, hi = oracle.normalize_array(3, raw)
    for name, packed in (("v.raw", 0), ("v.zraw", (tmp_path / "v.zraw").stat().st_size)):
        res, info = T.LoadRawIntoNewVolume(str(tmp_path / name), dims, np.int16, CompressedByteSize=packed, bLightVolume32Bit=True)
        assert info.Dimensions == dims and (info.MinValue, info.MaxValue) == (lo, hi) and info.bIsNormalized and res.DataFormat == 1
        URaymarchUtils.GenerateOctree(res)  # mip 0 of the octree is the (G16) data volume itself
        assert np.array_equal(URaymarchUtils.ReadOctreeMip(res, 0)[:dims[2], :dims[1], :dims[0]], want)
        res.release()


@pytest.mark.parametrize("gpu_sync", [False, True])
@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 3), (1, 7, 1), (16, 1, 1), (7, 3, 1), (5, 4, 6)])
def test_degenerate_and_ragged_sizes_match_oracle(dims, gpu_sync):
    """Edge cases through the C ABI: one-voxel and one-voxel-thick volumes, odd sizes (the oracle equals the reference's shaders on the same
    cases, tests/test_ref_shaders_cpu.py): sweep incl. axis-aligned lights and a ChangeDirLight, the three materials, the octree."""
    from tbraymarcherplugin_b200.raymarch_utils import FDirLightParameters

    rng = np.random.default_rng(sum(dims))
    data = rng.integers(0, 256, dims[::-1]).astype(np.uint8)
    win = FWindowingParameters(0.45, 0.5, True, False)
    cam = synth.benchmark_camera(24, 16, jitter=True, frame=1)
    lights = synth.LIGHTS + [FDirLightParameters((1, 0, 0), 0.7), FDirLightParameters((0, 1, 0), 0.3)]
    for world in (synth.identity_world(), synth.clipped_world()):
        res = make_res(data, win)
        vol = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), win)
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
        for l in lights:
            assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=gpu_sync)
            vol.add_dir_light(l, True, world)
        assert np.array_equal(URaymarchUtils.ReadLightVolume(res), vol.light)
        n = synth.rotate_about_z(synth.LIGHTS[0], 20.0)
        assert URaymarchUtils.ChangeDirLightInSingleVolume(res, synth.LIGHTS[0], n, world, bGPUSync=gpu_sync)
        vol.change_dir_light(synth.LIGHTS[0], n, world)
        assert np.array_equal(URaymarchUtils.ReadLightVolume(res), vol.light)
        rgba, steps = URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, 17.0)
        ref, ref_steps = vol.raymarch_lit(cam, world, 17.0)
        assert steps == ref_steps and np.array_equal(rgba, ref)
        assert np.array_equal(URaymarchUtils.PerformWindowedIntensityRaymarch(res, cam, world, 17.0)[0], oracle.raymarch_intensity(vol, cam, world, 17.0)[0])
        URaymarchUtils.GenerateOctree(res)
        mips = oracle.generate_octree(data)
        for mip in range(4):
            assert np.array_equal(URaymarchUtils.ReadOctreeMip(res, mip), mips[mip])
            assert np.array_equal(URaymarchUtils.PerformWindowedRaymarchOctree(res, cam, world, 17.0, mip)[0],
                                  oracle.raymarch_octree(vol, cam, world, 17.0, mips, mip)[0])
        res.release()


def test_raymarch_volume_actor_from_mhd_file_ticks_and_renders_every_material(tmp_path):
    """The caller of the boundary end to end: ARaymarchVolume.LoadMHDFileIntoVolumeNormalized -> Tick (full reset; octree rebuild under the
    octree material) -> Render with each material; the lit frame equals the oracle's for the same normalised voxels and world."""
    from tbraymarcherplugin_b200 import ARaymarchLight, ARaymarchVolume, ERaymarchMaterial
    from tbraymarcherplugin_b200.raymarch_utils import FBasicRaymarchRenderingResources

    dims = (32, 32, 16)
    raw = (synth.perlin_ct_volume(dims).astype(np.int16) * 7 - 500)
    (tmp_path / "v.raw").write_bytes(raw.tobytes())
    (tmp_path / "v.mhd").write_text(f"DimSize = {dims[0]} {dims[1]} {dims[2]}\nElementSpacing = 1 1 2\nElementType = MET_SHORT\nElementDataFile = v.raw\n")
    lights = [ARaymarchLight(tuple(l.LightDirection), l.LightIntensity, f"L{i}") for i, l in enumerate(synth.LIGHTS[:2])]
    vol = ARaymarchVolume(FBasicRaymarchRenderingResources(), lights)
    assert vol.Tick().action == "not_initialized"
    assert vol.LoadMHDFileIntoVolumeNormalized(str(tmp_path / "v.mhd"), bLightVolume32Bit=True)
    assert vol.ComponentTransform.Scale3D == (3.2, 3.2, 3.2)  # WorldDimensions / 10
    vol.SetWindowCenter(0.45), vol.SetWindowWidth(0.5), vol.SetHighCutoff(False), vol.SetRaymarchSteps(48)
    rep = vol.Tick()
    assert rep.action == "reset" and rep.lights_updated == 2 and not rep.errors
    cam = synth.benchmark_camera(48, 32)
    cam.Eye = tuple(3.2 * c for c in cam.Eye)  # the mesh is 3.2 units wide now
    lit, steps = vol.Render(cam)
    want, _, _ = oracle.normalize_array(3, raw)
    ora = oracle.OracleVolume(want, oracle.default_tf(), vol.RaymarchResources.WindowingParameters)
    for l in lights:
        ora.add_dir_light(l.GetCurrentParameters(), True, vol.WorldParameters)
    assert np.array_equal(URaymarchUtils.ReadLightVolume(vol.RaymarchResources), ora.light)
    ref, ref_steps = ora.raymarch_lit(cam, vol.WorldParameters, 48.0)
    assert steps == ref_steps and np.array_equal(lit, ref)
    vol.SwitchRenderer(ERaymarchMaterial.Octree)
    assert vol.Tick().octree_rebuilt
    assert np.array_equal(vol.Render(cam)[0], oracle.raymarch_octree(ora, cam, vol.WorldParameters, 48.0, oracle.generate_octree(want), 0)[0])
    vol.SwitchRenderer(ERaymarchMaterial.Intensity)
    assert np.array_equal(vol.Render(cam)[0], oracle.raymarch_intensity(ora, cam, vol.WorldParameters, 48.0)[0])
    lights[0].ForwardVector = tuple(synth.rotate_about_z(synth.LIGHTS[0], 5.0).LightDirection)
    vol.SwitchRenderer(ERaymarchMaterial.Lit)
    assert vol.Tick().action == "incremental"
    ora.change_dir_light(synth.LIGHTS[0], lights[0].GetCurrentParameters(), vol.WorldParameters)
    assert np.array_equal(URaymarchUtils.ReadLightVolume(vol.RaymarchResources), ora.light)
    vol.RaymarchResources.release()


@pytest.mark.parametrize("light32", [True, False])
@pytest.mark.parametrize("dims", [(40, 32, 24), (64, 64, 64), (33, 17, 9)])
def test_joined_same_axis_sweeps_match_their_cpu_twin(dims, light32):
    """tbrm_add_dir_lights_joined against the oracle's twin (bit-exact), against consecutive AddDirLight calls (equal up
    to summation order), and the launch count it saves in the per-slice schedule."""
    from tbraymarcherplugin_b200.raymarch_utils import FDirLightParameters, FSweepStats

    data = synth.perlin_ct_volume(dims)
    win = FWindowingParameters(0.45, 0.5, True, False)
    lights = synth.LIGHTS + [synth.rotate_about_z(synth.LIGHTS[0], 7.0), synth.rotate_about_z(synth.LIGHTS[2], -9.0), FDirLightParameters((0, 0, 0), 1.0)]
    for world in (synth.identity_world(), synth.clipped_world()):
        Z, Y, X = data.shape
        res = URaymarchUtils.InitializeRaymarchResources((X, Y, Z), FMT_G8, bLightVolume32Bit=light32)
        URaymarchUtils.SetDataVolume(res, data)
        URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
        URaymarchUtils.SetWindowingParameters(res, win)
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
        st = FSweepStats()
        assert URaymarchUtils.AddDirLightsToSingleVolumeJoined(res, lights, True, world, stats=st)
        twin = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), win, light32=light32)
        n_twin = oracle.add_dir_lights_joined(twin, lights, True, world)
        assert st.passes == n_twin and np.array_equal(URaymarchUtils.ReadLightVolume(res), twin.light)
        # consecutive per-light adds (the reference's schedule): same volume up to summation order, more launches
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
        launches = 0
        for l in lights:
            s1 = FSweepStats()
            assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=False, stats=s1)
            launches += s1.kernel_launches
        seq = URaymarchUtils.ReadLightVolume(res)
        d = np.abs(seq.astype(np.float64) - twin.light.astype(np.float64))
        assert d.max() <= (4e-6 if light32 else 1.0)
        assert st.kernel_launches < launches
        res.release()
