"""GPU tests written after round 1's GPU budget was spent — they have run against an oracle-backed stand-in only (tests/test_gpu_test_logic_cpu.py)
and sort after every test that has already passed on a B200 (file name) and, within the file, from the everyday paths to the edge cases: the
Power == 8 Mandelbulb fast path against its CPU twin, the headerless raw loader, the ARaymarchVolume mirror from an MHD file, same-axis light
joining, the C++ example end to end, and degenerate / ragged volume sizes. Same bars as tests/test_gpu_zz_materials.py."""
import zlib

import numpy as np
import pytest

import oracle
from test_gpu_zz_materials import make_res
from tbraymarcherplugin_b200 import FMT_G8, synth
from tbraymarcherplugin_b200.raymarch_utils import FWindowingParameters, URaymarchUtils

# never run on a GPU yet: a hang (e.g. a barrier that never completes) must end the run instead of holding the box
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]

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
    # measured on a B200 (profiles/r2_mandelbulb_mismatch_p8.json): 0.009 % vs the twin (only the final log differs: 1 ulp, amplified next to
    # the surface), 0.12 % vs the reference's formulation
    print(f"mandelbulb p8: {100 * bad_twin:.4f} % of the pixels beyond 1e-4 vs the CPU twin, {100 * bad_ref:.4f} % vs the reference formulation")
    assert bad_twin <= 0.001 and bad_ref <= 0.005, (bad_twin, bad_ref)
    assert abs(iters - twin_iters) / twin_iters < 5e-3
    assert (np.abs(gsdf - tsdf) > 1e-5).mean() <= 0.01
    mb6 = FMandelbulbParameters(MaxSteps=64.0, MaxIterations=8.0, Power=6.0)
    got6, _ = URaymarchUtils.PerformMandelbulbRaymarchReturnDistance(mb6, cam, world)
    ref6, _ = oracle.mandelbulb(mb6, cam, world)
    assert (np.abs(got6 - ref6).max(-1) > 1e-4).mean() <= 0.005


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


def test_cpp_example_runs_end_to_end(tmp_path, libdir=None, libname="tbrm", view=()):
    """examples/mhd_to_frame.cpp (plain C++ over the C ABI): MetaImage file -> resources -> sweep -> octree -> the three materials."""
    import subprocess

    from test_ingest_cpu import _build_example

    dims = (48, 40, 32)
    raw = (synth.perlin_ct_volume(dims).astype(np.int16) * 12 - 1000)
    (tmp_path / "v.raw").write_bytes(raw.tobytes())
    (tmp_path / "v.mhd").write_text(f"NDims = 3\nDimSize = {dims[0]} {dims[1]} {dims[2]}\nElementSpacing = 1 1 1\nElementType = MET_SHORT\nElementDataFile = v.raw\n")
    out = subprocess.run([str(_build_example(tmp_path, libdir, libname)), str(tmp_path / "v.mhd"), *view], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "48 x 40 x 32 voxels" in out.stdout and "normalised to G16" in out.stdout
    steps = [int(l.split(":")[1].split()[0]) for l in out.stdout.splitlines() if "march:" in l]
    assert len(steps) == 3 and all(s > 0 for s in steps) and steps[1] < steps[0]  # the intensity march stops at its first sample


@pytest.mark.parametrize("dims", [(1, 1, 1), (1, 7, 1), (16, 1, 1), (2, 2, 2), (5, 4, 6), (40, 24, 56)])
def test_second_generation_raymarch_on_small_and_degenerate_volumes(dims):
    """raymarch_fast2_kernel (opt-in, reserved[1] = 3) against the oracle on sizes where a side has no interior (1 or 2 voxels) — its interior
    test once accepted tap index -1 for a one-voxel side — and on an ordinary small volume, with and without a clip plane."""
    rng = np.random.default_rng(sum(dims))
    data = rng.integers(0, 256, dims[::-1]).astype(np.uint8)
    win = FWindowingParameters(0.45, 0.5, True, False)
    for world in (synth.identity_world(), synth.clipped_world(), synth.scaled_rotated_world()):
        res = make_res(data, win)
        vol = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), win)
        for l in synth.LIGHTS[:2]:
            assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world)
            vol.add_dir_light(l, True, world)
        URaymarchUtils.SetOptions(res, debug_flags=(0, 3))
        for jitter in (False, True):
            cam = synth.benchmark_camera(40, 24, jitter=jitter, frame=2)
            rgba, steps = URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, 33.0)
            ref, ref_steps = vol.raymarch_lit(cam, world, 33.0)
            assert steps == ref_steps and np.array_equal(rgba, ref), (dims, jitter)
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


def derived_structures_match_numpy(dims):
    """brick grid and (y,z,x) replica through tbrm_debug_download_derived against their definitions in numpy"""
    import ctypes as C

    from tbraymarcherplugin_b200 import _capi

    X, Y, Z = dims
    data = np.random.default_rng(sum(dims)).integers(0, 256, (Z, Y, X)).astype(np.uint8)
    data[data < 200] //= 4  # mostly small values with rare large ones: maxima that depend on single voxels
    res = make_res(data, FWindowingParameters(0.45, 0.5, True, False))
    lib = _capi.load()
    B = [(d + 7) // 8 for d in dims]
    bricks = np.zeros((B[2], B[1], B[0]), np.uint8)
    _capi.check(lib.tbrm_debug_download_derived(res.handle, 0, bricks.ctypes.data_as(C.c_void_p), bricks.size))
    padded = np.zeros((8 * B[2] + 1, 8 * B[1] + 1, 8 * B[0] + 1), np.uint8)
    padded[:Z, :Y, :X] = data
    want = np.zeros_like(bricks)
    for dz in range(9):
        for dy in range(9):
            for dx in range(9):
                want = np.maximum(want, padded[dz:dz + 8 * B[2]:8, dy:dy + 8 * B[1]:8, dx:dx + 8 * B[0]:8])
    assert np.array_equal(bricks, want), dims
    replica = np.zeros((X, Z, Y), np.uint8)
    _capi.check(lib.tbrm_debug_download_derived(res.handle, 1, replica.ctypes.data_as(C.c_void_p), replica.size))
    assert np.array_equal(replica, data.transpose(2, 0, 1)), dims
    res.release()


@pytest.mark.parametrize("dims", [(16, 16, 8), (144, 80, 40), (64, 64, 64), (48, 33, 17), (40, 24, 16), (256, 128, 72)])
def test_brick_grid_and_axis_replica_match_their_definitions(dims):
    """brick_parts_kernel + brick_combine_kernel / permute_yzx_vec_kernel (X % 16 == 0 [and Y % 16 == 0]) and the byte-wise kernels they
    replace on other sizes."""
    derived_structures_match_numpy(dims)


@pytest.mark.parametrize("dims", [(40, 24, 56), (1, 7, 1)])
def test_lit_march_with_64_bit_tap_addressing_still_matches_oracle(dims):
    """The default lit march forms tap addresses from 32-bit offsets (ADDR32, csrc/raymarch.cu); reserved[1] = 2 selects the 64-bit pointer
    form that round 1 measured. Both must give the oracle's frame."""
    data = np.random.default_rng(sum(dims)).integers(0, 256, dims[::-1]).astype(np.uint8)
    win = FWindowingParameters(0.45, 0.5, True, False)
    for world in (synth.identity_world(), synth.clipped_world()):
        res = make_res(data, win)
        vol = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), win)
        for l in synth.LIGHTS[:2]:
            assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world)
            vol.add_dir_light(l, True, world)
        cam = synth.benchmark_camera(40, 24, jitter=True, frame=2)
        ref, ref_steps = vol.raymarch_lit(cam, world, 33.0)
        for flag in (0, 2, 3, 4, 5):  # default (second generation without leaps), 64-bit addressing, leaps, no leaps, first generation (ADDR32)
            URaymarchUtils.SetOptions(res, debug_flags=(0, flag))
            rgba, steps = URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, 33.0)
            assert steps == ref_steps and np.array_equal(rgba, ref), (dims, flag)
        res.release()


@pytest.mark.parametrize("light32", [True, False])
@pytest.mark.parametrize("dims", [(64, 64, 40), (96, 32, 64)])  # (a G8 light volume's TMA strides: X and Y of the LIGHT volume multiples of 16)
def test_half_resolution_light_volume_through_the_tma_staged_sweep(dims, light32):
    """A half-resolution light volume (RaymarchVolume.h: LightVolumeHalfResolution; two data voxels per light voxel along every axis) runs the
    TMA-staged sweep in its one-pixel form: the data box of a tile starts at twice the tile's origin and spans twice its extent. R32F and G8,
    AddDirLight (oblique and axis-aligned lights), ChangeDirLight, with and without a clip plane — bit-exact against the oracle."""
    from tbraymarcherplugin_b200.raymarch_utils import FDirLightParameters, FSweepStats

    data = synth.perlin_ct_volume(dims)
    win = FWindowingParameters(0.45, 0.5, True, False)
    for world in (synth.identity_world(), synth.clipped_world()):
        res = URaymarchUtils.InitializeRaymarchResources(dims, FMT_G8, bLightVolume32Bit=light32, LightVolumeHalfResolution=True)
        URaymarchUtils.SetDataVolume(res, data)
        URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
        URaymarchUtils.SetWindowingParameters(res, win)
        URaymarchUtils.SetOptions(res, sweep_impl=2)
        vol = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), win, light32=light32, half_res=True)
        assert tuple(res.LightDims) == tuple(vol.ldims) == tuple(d // 2 for d in dims)
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
        used = []
        for l in synth.LIGHTS + [FDirLightParameters((1, 0, 0), 0.7), FDirLightParameters((0.1, 0.2, 1.0), 0.4)]:
            st = FSweepStats()
            assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True, stats=st)
            vol.add_dir_light(l, True, world)
            used.append(tuple(st.impl))
            assert np.array_equal(URaymarchUtils.ReadLightVolume(res), vol.light), (dims, light32, l.LightDirection, st.impl)
        if dims == (64, 64, 40):  # (on the flat volume a second-axis pass reads further across the plane than the kernel's footprint holds)
            assert all(set(i) == {3} for i in used[:3]), f"the TMA-staged sweep must have taken every pass of the oblique lights: {used}"
        assert all(3 in i for i in used[:3]), used
        n = synth.rotate_about_z(synth.LIGHTS[0], 20.0)
        assert URaymarchUtils.ChangeDirLightInSingleVolume(res, synth.LIGHTS[0], n, world, bGPUSync=True)
        vol.change_dir_light(synth.LIGHTS[0], n, world)
        assert np.array_equal(URaymarchUtils.ReadLightVolume(res), vol.light), (dims, light32, "change")
        res.release()


@pytest.mark.parametrize("height", [64, 70, 33])
def test_interleaved_rows_of_a_frame_equal_the_whole_frame(height):
    """What the ranks of a sharded volume render (tbrm_raymarch_lit_interleaved: blocks of 8 image rows dealt round-robin, compacted per
    rank) assembled back into the frame, for 2 / 3 / 4 ranks on ONE GPU: equal to the frame marched in one piece, step counts included —
    the march's thread blocks (8 x 16 pixels) span two 8-row blocks of a rank's compacted rows."""
    import ctypes as C

    from tbraymarcherplugin_b200 import _capi, sharding

    lib = _capi.load()
    dims = (48, 40, 56)
    data = synth.perlin_ct_volume(dims)
    win = FWindowingParameters(0.45, 0.5, True, False)
    res = make_res(data, win)
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    for l in synth.LIGHTS[:2]:
        URaymarchUtils.AddDirLightToSingleVolume(res, l, True, synth.identity_world(), bGPUSync=True)
    cam = synth.benchmark_camera(100, height)
    for world in (synth.identity_world(), synth.clipped_world()):
        whole, whole_steps = URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, 48.0)
        c, w = cam.to_c(), world.to_c()
        for nranks in (2, 3, 4):
            parts, total = [], 0
            for rank in range(nranks):
                rows = int(lib.tbrm_raymarch_interleaved_rows(height, 8, rank, nranks))
                assert rows == len(sharding.rows_of_rank(height, rank, nranks, 8))
                out = np.full((max(rows, 1), cam.Width, 4), -1.0, np.float32)
                steps = C.c_uint64(0)
                _capi.check(lib.tbrm_raymarch_lit_interleaved(res.handle, C.byref(c), C.byref(w), 48.0, 8, rank, nranks,
                                                              out.ctypes.data_as(C.c_void_p), 0, C.byref(steps)))
                parts.append([out[i:i + (e - b)] for i, (b, e) in zip(np.cumsum([0] + [e - b for b, e in sharding.row_blocks_of_rank(height, rank, nranks, 8)])[:-1],
                                                                          sharding.row_blocks_of_rank(height, rank, nranks, 8))])
                total += steps.value
            frame = sharding.assemble_rows(height, nranks, parts, 8)
            assert np.array_equal(frame, whole), (height, nranks)
            assert total == whole_steps, (height, nranks, total, whole_steps)
    res.release()


def test_a_512_squared_plane_takes_seven_row_tiles_and_stays_bit_exact():
    """What the headline workload runs, on a thin volume: a sweep along Z over a 512 x 512 plane is 592 tiles of 64 x 7 pixels = four blocks on
    each of 148 SMs (512 tiles of 64 x 8 would leave SMs with three), chosen by the host, bit-exact against the oracle; a 256 x 512 plane
    (256 tiles of 8 rows: two blocks on most SMs, one on the rest) takes 7 rows as well (296 = 2 x 148), a 128 x 128 plane keeps 8."""
    import ctypes as C

    from tbraymarcherplugin_b200 import _capi
    from tbraymarcherplugin_b200.raymarch_utils import FDirLightParameters, FSweepStats

    lib = _capi.load()
    win = FWindowingParameters(0.45, 0.5, True, False)
    for dims, want in (((512, 512, 8), (7, 2, 592, 1)), ((512, 256, 8), (7, 2, 296, 1)), ((128, 128, 8), (8, 1, 64, 1))):
        data = synth.perlin_ct_volume(dims)
        res = make_res(data, win)
        URaymarchUtils.SetOptions(res, sweep_impl=2)
        vol = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), win)
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
        for l in (FDirLightParameters((0.0, 0.0, -1.0), 0.9), FDirLightParameters((0.0, 0.0, 1.0), 0.4)):  # one sweep along Z each, both directions
            st = FSweepStats()
            assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, synth.identity_world(), bGPUSync=True, stats=st)
            vol.add_dir_light(l, True, synth.identity_world())
            assert tuple(st.impl) == (3,), st.impl
            geom = (C.c_int32 * 4)()
            _capi.check(lib.tbrm_debug_download_derived(res.handle, 4, geom, C.sizeof(geom)))
            assert tuple(geom) == want, (dims, tuple(geom))
        assert np.array_equal(URaymarchUtils.ReadLightVolume(res), vol.light), dims
        res.release()


# bits 4-5 of tbrm_options.reserved[0]: one / two pixels per thread, automatic; bit 6: the second kernel generation (occlusion kernel + chain kernel);
# bits 8-9: tiles of 6 / 7 / 8 rows (256 / 512 / 768; 0 = chosen per launch)
@pytest.mark.parametrize("px_flag", [16, 32, 48, 64 + 16, 64 + 32, 256 + 16, 256 + 32, 512 + 16, 512 + 32, 768 + 48])
@pytest.mark.parametrize("dims", [(64, 48, 40), (80, 24, 16), (128, 16, 8)])
def test_tma_sweep_with_one_and_two_pixels_per_thread(dims, px_flag):
    """The TMA-staged sweep has a one-pixel-per-thread form (tile 32 x 8) for launches that cannot fill the SMs next to the two-pixel form
    (tile 64 x 8; the default until the other has run on a GPU); both forced here on the same volumes, and the automatic choice, incl. planes that end inside a tile: AddDirLight (oblique and axis-aligned lights),
    ChangeDirLight, with and without a clip plane — bit-exact against the oracle."""
    from tbraymarcherplugin_b200.raymarch_utils import FDirLightParameters, FSweepStats

    data = np.random.default_rng(sum(dims)).integers(0, 256, dims[::-1]).astype(np.uint8)
    win = FWindowingParameters(0.45, 0.5, True, False)
    used = set()
    for world in (synth.identity_world(), synth.clipped_world()):
        res = make_res(data, win)
        URaymarchUtils.SetOptions(res, sweep_impl=2, debug_flags=(px_flag, 0))
        vol = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), win)
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
        for l in synth.LIGHTS + [FDirLightParameters((1, 0, 0), 0.7), FDirLightParameters((0, -1, 0), 0.3)]:
            st = FSweepStats()
            assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True, stats=st)
            vol.add_dir_light(l, True, world)
            used |= set(st.impl)
        assert np.array_equal(URaymarchUtils.ReadLightVolume(res), vol.light), (dims, px_flag)
        n = synth.rotate_about_z(synth.LIGHTS[0], 20.0)
        assert URaymarchUtils.ChangeDirLightInSingleVolume(res, synth.LIGHTS[0], n, world, bGPUSync=True)
        vol.change_dir_light(synth.LIGHTS[0], n, world)
        assert np.array_equal(URaymarchUtils.ReadLightVolume(res), vol.light), (dims, px_flag)
        res.release()
    assert 3 in used  # the TMA-staged sweep took passes


CPP_ACTOR_PROGRAM = r'''
#include <array>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>
#include "tbraymarcherplugin_b200/csrc/RaymarchVolume.hpp"
using namespace tbrm_ue;
template <typename T>
static std::vector<T> slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    std::vector<char> b((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return std::vector<T>((const T*) b.data(), (const T*) (b.data() + b.size()));
}
template <typename T>
static void dump(const std::string& path, const std::vector<T>& v) {
    std::ofstream(path, std::ios::binary).write((const char*) v.data(), (std::streamsize) (v.size() * sizeof(T)));
}
#define CHECK(c) do { if (!(c)) { std::printf("failed: %s (line %d): %s\n", #c, __LINE__, tbrm_last_error()); return 1; } } while (0)
int main(int argc, char** argv) {
    const std::string dir = argv[1];
    const int32_t dims[3] = {std::atoi(argv[2]), std::atoi(argv[3]), std::atoi(argv[4])};
    const std::vector<uint8_t> data = slurp<uint8_t>(dir + "/data.raw");
    const std::vector<float> curve = slurp<float>(dir + "/curve.raw");
    const std::vector<double> lights = slurp<double>(dir + "/lights.raw");  // 3 lights x (direction, intensity): two initial, one moved
    const std::vector<tbrm_camera> cam = slurp<tbrm_camera>(dir + "/camera.raw");
    CHECK(data.size() == (size_t) dims[0] * dims[1] * dims[2] && curve.size() == 1024 && lights.size() == 12 && cam.size() == 1);
    ARaymarchVolume<> vol;  // the real operator surface: URaymarchUtils over the C ABI
    vol.RaymarchResources.WindowingParameters = FWindowingParameters{0.45f, 0.5f, true, false};
    CHECK(URaymarchUtils::InitializeRaymarchResources(vol.RaymarchResources, dims, TBRM_FMT_G8, data.data(), /*bLightVolume32Bit*/ true));
    std::array<float, 1024> c;
    for (int i = 0; i < 1024; ++i) c[i] = curve[i];
    URaymarchUtils::ColorCurveToTexture(c, vol.RaymarchResources);
    ARaymarchLight L[2];
    for (int i = 0; i < 2; ++i) L[i].ForwardVector = FVector(lights[4 * i], lights[4 * i + 1], lights[4 * i + 2]), L[i].LightIntensity = (float) lights[4 * i + 3];
    vol.LightsArray = {&L[0], &L[1]};
    vol.OnConstruction();
    vol.bRequestedRecompute = true;
    CHECK(vol.Tick().action == FTickReport::Reset);
    std::vector<float> light((size_t) dims[0] * dims[1] * dims[2]);
    CHECK(tbrm_download_light_volume(vol.RaymarchResources.Handle, light.data()) == TBRM_OK);
    dump(dir + "/light_reset.raw", light);
    vol.LightParametersMap[&L[0]] = L[0].GetCurrentParameters();  // what the reset used (the reference leaves the map stale)
    vol.LightParametersMap[&L[1]] = L[1].GetCurrentParameters();
    L[0].ForwardVector = FVector(lights[8], lights[9], lights[10]);
    const FTickReport rep = vol.Tick();
    CHECK(rep.action == FTickReport::Incremental && rep.lights_updated == 1 && rep.errors.empty());
    CHECK(tbrm_download_light_volume(vol.RaymarchResources.Handle, light.data()) == TBRM_OK);
    dump(dir + "/light_changed.raw", light);
    std::vector<float> frame((size_t) cam[0].width * cam[0].height * 4);
    uint64_t steps = 0;
    CHECK(URaymarchUtils::PerformWindowedLitRaymarch(vol.RaymarchResources, cam[0], vol.WorldParameters, 40.0f, frame.data(), &steps));
    dump(dir + "/frame.raw", frame);
    std::printf("steps %llu\n", (unsigned long long) steps);
    URaymarchUtils::FreeRaymarchResources(vol.RaymarchResources);
    return 0;
}
'''


def test_cpp_actor_mirror_end_to_end(tmp_path, libdir=None, libname="tbrm"):
    """csrc/RaymarchVolume.hpp + csrc/RaymarchUtils.hpp (the C++ host mirror of ARaymarchVolume / URaymarchUtils) driving the C ABI for real:
    resources from a host volume, a full reset and an incremental ChangeDirLight decided by Tick, a lit frame — all three bit-identical to
    the oracle."""
    import ctypes as C
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    libdir = libdir or root / "tbraymarcherplugin_b200"
    dims = (48, 32, 24)
    data = synth.perlin_ct_volume(dims)
    data.tofile(tmp_path / "data.raw")
    synth.soft_ct_curve().astype(np.float32).tofile(tmp_path / "curve.raw")
    moved = synth.rotate_about_z(synth.LIGHTS[0], 5.0)
    lights = [synth.LIGHTS[0], synth.LIGHTS[1], moved]
    np.array([[*l.LightDirection, l.LightIntensity] for l in lights], np.float64).tofile(tmp_path / "lights.raw")
    cam = synth.benchmark_camera(48, 32, jitter=True, frame=1)
    (tmp_path / "camera.raw").write_bytes(bytes(cam.to_c()))
    (tmp_path / "p.cpp").write_text(CPP_ACTOR_PROGRAM)
    exe = tmp_path / "p"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-I", str(root), str(tmp_path / "p.cpp"), "-L", str(libdir), f"-l{libname}",
                    f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), str(tmp_path), *map(str, dims)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    win = FWindowingParameters(0.45, 0.5, True, False)
    world = synth.identity_world()
    vol = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), win)
    for l in lights[:2]:
        vol.add_dir_light(l, True, world)
    shape = dims[::-1]
    assert np.array_equal(np.fromfile(tmp_path / "light_reset.raw", np.float32).reshape(shape), vol.light)
    vol.change_dir_light(lights[0], moved, world)
    assert np.array_equal(np.fromfile(tmp_path / "light_changed.raw", np.float32).reshape(shape), vol.light)
    ref, ref_steps = vol.raymarch_lit(cam, world, 40.0)
    assert np.array_equal(np.fromfile(tmp_path / "frame.raw", np.float32).reshape(ref.shape), ref)
    assert int(out.stdout.split("steps")[1]) == ref_steps
