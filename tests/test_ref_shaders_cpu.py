"""Pins the oracle to the REFERENCE'S OWN SHADER CODE. oracle/ref.mk streams AddDirLightShader.usf, ChangeDirLightShader.usf,
WindowedRaymarchMaterials.usf (+ RaymarchMaterialCommon / WindowedSampling / RaymarcherCommon), GenerateOctreeShader.usf, SDFMarcher.usf and
CalculateMandelbulbSDF.usf from /root/reference through syntactic rewrites (oracle/hlsl2cpp.py) and compiles them for the CPU against
oracle/hlsl_shim; oracle/ref_shaders.cpp dispatches them with the uniforms the reference's own host code computes.

  * tests/golden/ref_shaders.npz holds outputs of that build (tests/golden/make_golden_ref.py): the oracle must reproduce them
    bit for bit — this part runs everywhere;
  * with oracle/_ref/libtbrm_ref.so present the comparison also runs live over more formats, sizes and random lights.

A match proves that the oracle restates the shaders' logic (gating, thresholds, operation order, addressing, the schedule of the
host drivers). The engine semantics under the shaders (samplers, UNORM conversions, pow) are the shim's — the same policies the
oracle states — and stay unpinned."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

import oracle
import refpin
from tbraymarcherplugin_b200 import synth
from tbraymarcherplugin_b200.raymarch_utils import FDirLightParameters, FWindowingParameters

GOLDEN = Path(__file__).resolve().parent / "golden"
_spec = importlib.util.spec_from_file_location("make_golden_ref", GOLDEN / "make_golden_ref.py")
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)

needs_ref = pytest.mark.skipif(not refpin.available(), reason="oracle/_ref/libtbrm_ref.so not built and /root/reference absent")


@pytest.mark.parametrize("world", list(mk.SHADER_WORLDS))
@pytest.mark.parametrize("light32", [True, False])
def test_oracle_reproduces_the_reference_shaders_golden_outputs(world, light32):
    want = np.load(GOLDEN / "ref_shaders.npz")
    data, tf, win = mk.shader_inputs()
    vol = oracle.OracleVolume(data, tf, win, light32=light32)
    got = {}
    tag = f"{world}_{'r32f' if light32 else 'g8'}"
    mk.shader_sequence(vol, mk.SHADER_WORLDS[world](), got, tag)
    for k, v in got.items():
        assert v.dtype == want[k].dtype and np.array_equal(v, want[k]), f"{k}: oracle differs from the reference shader"
    if light32:
        cam = synth.benchmark_camera(40, 24, jitter=True, frame=3)
        w = mk.SHADER_WORLDS[world]()
        assert np.array_equal(oracle.cube_setup(cam, w), want[f"{tag}_setup"])
        rgba, _ = vol.raymarch_lit(cam, w, 48.0)
        assert np.array_equal(rgba, want[f"{tag}_lit"])


def test_reference_shader_golden_outputs_are_not_trivial():
    g = np.load(GOLDEN / "ref_shaders.npz")
    assert g["identity_r32f_reset"].max() > 1.5 and g["identity_g8_reset"].max() == 255
    for a, b in (("reset", "removed"), ("removed", "changed"), ("changed", "changed_fallback")):
        assert not np.array_equal(g[f"identity_r32f_{a}"], g[f"identity_r32f_{b}"])
    assert not np.array_equal(g["identity_r32f_reset"], g["clipped_r32f_reset"])
    assert g["identity_r32f_lit"][..., 3].max() > 0.3 and (g["identity_r32f_setup"][..., 3] > 0).mean() > 0.3


@needs_ref
def test_golden_outputs_are_what_the_reference_build_produces_today():
    want = np.load(GOLDEN / "ref_shaders.npz")
    got = mk.shaders_case()
    assert set(want.files) == set(got)
    for k in want.files:
        assert np.array_equal(want[k], got[k]), k


@needs_ref
@pytest.mark.parametrize("dtype,half_res,border_exact", [(np.uint8, False, False), (np.uint16, False, False), (np.float32, False, True),
                                                         (np.uint8, True, False)])
def test_oracle_equals_live_reference_shaders(dtype, half_res, border_exact):
    rng = np.random.default_rng(11)
    base = synth.perlin_ct_volume((20, 28, 36))
    data = base if dtype == np.uint8 else (base.astype(np.uint16) * 257 if dtype == np.uint16 else (base / np.float32(255)).astype(np.float32))
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    for win, world in ((FWindowingParameters(0.45, 0.5, True, False), synth.scaled_rotated_world()),
                       (FWindowingParameters(0.5, 1.0, True, True), synth.clipped_world())):
        kw = dict(light32=True, half_res=half_res, border_exact=border_exact)
        a, b = oracle.OracleVolume(data, tf, win, **kw), refpin.RefVolume(data, tf, win, **kw)
        lights = [FDirLightParameters(tuple(rng.standard_normal(3)), float(rng.uniform(0.3, 1.2))) for _ in range(3)]
        lights.append(FDirLightParameters((0, 0, -1), 0.5))  # axis-aligned: a single pass
        for l in lights:
            a.add_dir_light(l, True, world), b.add_dir_light(l, True, world)
            assert np.array_equal(a.light, b.light)
        for l in lights[:3]:
            n = synth.rotate_about_z(l, 7.0)
            a.change_dir_light(l, n, world), b.change_dir_light(l, n, world)
            assert np.array_equal(a.light, b.light)
        cam = synth.benchmark_camera(36, 28, jitter=True, frame=5)
        rgba, _ = a.raymarch_lit(cam, world, 40.0)
        assert np.array_equal(rgba, b.raymarch(0, cam, world, 40.0))


@needs_ref
def test_the_oracle_written_goldens_of_the_gpu_tests_are_reference_shader_output_too():
    """tests/golden/sweep_32.npz and raymarch_32.npz were written by the oracle (make_golden.py); tests/test_gpu_golden.py holds the
    TMA-staged sweep (the dominant kernel; 32^3 is eligible) and the fast lit march to them. The reference's own shaders produce the
    same arrays — so those GPU tests compare the CUDA path with reference output."""
    spec = importlib.util.spec_from_file_location("make_golden", GOLDEN / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want = np.load(GOLDEN / "sweep_32.npz")
    data = synth.perlin_ct_volume(mg.SWEEP_DIMS)
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    for name, mkw in mg.WORLDS.items():
        vol, world = refpin.RefVolume(data, tf, mg.CT_WINDOW), mkw()
        for l in synth.LIGHTS:
            vol.add_dir_light(l, True, world)
        assert np.array_equal(vol.light, want[f"{name}_reset"])
        vol.add_dir_light(synth.LIGHTS[1], False, world)
        assert np.array_equal(vol.light, want[f"{name}_removed"])
        vol.change_dir_light(synth.LIGHTS[0], synth.rotate_about_z(synth.LIGHTS[0], 5.0), world)
        assert np.array_equal(vol.light, want[f"{name}_changed"])
    frames = np.load(GOLDEN / "raymarch_32.npz")
    vol, world = refpin.RefVolume(data, tf, mg.CT_WINDOW), synth.identity_world()
    for l in synth.LIGHTS[:2]:
        vol.add_dir_light(l, True, world)
    cam = synth.benchmark_camera(48, 32, jitter=True, frame=3)  # the shader always jitters: only the jitter = 1 frame has a counterpart
    assert np.array_equal(vol.raymarch(0, cam, world, 64.0), frames["rgba_jitter1"])
    assert np.array_equal(vol.raymarch(-1, cam, world, 64.0), frames["setup_jitter1"])


@needs_ref
@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 3), (1, 7, 1), (16, 1, 1), (7, 3, 1), (5, 4, 6)])
def test_degenerate_and_ragged_sizes_equal_the_reference_shaders(dims):
    """Edge cases: one-voxel and one-voxel-thick volumes, odd sizes — random data, three windows, G8 / R32F and half-resolution light volumes,
    axis-aligned lights, a ChangeDirLight, all three materials and the octree."""
    rng = np.random.default_rng(sum(dims))
    data = rng.integers(0, 256, dims[::-1]).astype(np.uint8)
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    lights = synth.LIGHTS + [FDirLightParameters((1, 0, 0), 0.7), FDirLightParameters((0, 1, 0), 0.3)]
    cam = synth.benchmark_camera(24, 16, jitter=True, frame=1)
    for win in (FWindowingParameters(0.45, 0.5, True, False), FWindowingParameters(0.5, 1.0, True, True), FWindowingParameters(0.2, 0.1, False, False)):
        for world in (synth.identity_world(), synth.clipped_world()):
            for light32, half in ((True, False), (False, False), (True, True)):
                a = oracle.OracleVolume(data, tf, win, light32=light32, half_res=half)
                b = refpin.RefVolume(data, tf, win, light32=light32, half_res=half)
                for l in lights:
                    a.add_dir_light(l, True, world), b.add_dir_light(l, True, world)
                    assert np.array_equal(a.light, b.light)
                n = synth.rotate_about_z(synth.LIGHTS[0], 20.0)
                a.change_dir_light(synth.LIGHTS[0], n, world), b.change_dir_light(synth.LIGHTS[0], n, world)
                assert np.array_equal(a.light, b.light)
                if light32:
                    assert np.array_equal(a.raymarch_lit(cam, world, 17.0)[0], b.raymarch(0, cam, world, 17.0))
                    assert np.array_equal(oracle.raymarch_intensity(a, cam, world, 17.0)[0], b.raymarch(1, cam, world, 17.0))
    ma, mb = oracle.generate_octree(data), refpin.generate_octree(data)
    assert all(np.array_equal(x, y) for x, y in zip(ma, mb))
    a, b = oracle.OracleVolume(data, tf, FWindowingParameters()), refpin.RefVolume(data, tf, FWindowingParameters())
    for mip in range(4):
        assert np.array_equal(oracle.raymarch_octree(a, cam, synth.identity_world(), 17.0, ma, mip)[0],
                              b.raymarch(2, cam, synth.identity_world(), 17.0, octree=mb, octree_mip=mip))
