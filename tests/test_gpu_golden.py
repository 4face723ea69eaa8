"""The CUDA path (through the C ABI) against the committed vectors of tests/golden/ — the same inputs and operations as
tests/golden/make_golden.py ran through the oracle. Bit-exact for the sweep and the lit ray march, tolerance for
the Mandelbulb march."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

from tbraymarcherplugin_b200 import FMT_G8, synth
from tbraymarcherplugin_b200.raymarch_utils import FMandelbulbParameters, FSweepStats, URaymarchUtils

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
_spec = importlib.util.spec_from_file_location("make_golden", GOLDEN / "make_golden.py")
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


def make_res(sweep_impl):
    data = synth.perlin_ct_volume(make_golden.SWEEP_DIMS)
    res = URaymarchUtils.InitializeRaymarchResources(make_golden.SWEEP_DIMS, FMT_G8, bLightVolume32Bit=True)
    URaymarchUtils.SetDataVolume(res, data)
    URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(res, make_golden.CT_WINDOW)
    URaymarchUtils.SetOptions(res, sweep_impl=sweep_impl)
    return res


@pytest.mark.parametrize("impl", [1, 2, 3])  # per-slice launches, fused (TMA-staged when eligible), generic fused
@pytest.mark.parametrize("world_name", list(make_golden.WORLDS))
def test_sweep_equals_golden(world_name, impl):
    want = np.load(GOLDEN / "sweep_32.npz")
    world = make_golden.WORLDS[world_name]()
    res = make_res(impl)
    sync = impl >= 2
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    used = set()
    for l in synth.LIGHTS:
        st = FSweepStats()
        assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=sync, stats=st)
        used |= set(st.impl)
    if impl == 2:
        assert 3 in used, f"the TMA-staged sweep never ran: {used}"
    assert np.array_equal(URaymarchUtils.ReadLightVolume(res), want[f"{world_name}_reset"])
    URaymarchUtils.AddDirLightToSingleVolume(res, synth.LIGHTS[1], False, world, bGPUSync=sync)
    assert np.array_equal(URaymarchUtils.ReadLightVolume(res), want[f"{world_name}_removed"])
    assert URaymarchUtils.ChangeDirLightInSingleVolume(res, synth.LIGHTS[0], synth.rotate_about_z(synth.LIGHTS[0], 5.0), world, bGPUSync=sync)
    assert np.array_equal(URaymarchUtils.ReadLightVolume(res), want[f"{world_name}_changed"])


@pytest.mark.parametrize("jitter", [0, 1])
def test_lit_raymarch_and_cube_setup_equal_golden(jitter):
    want = np.load(GOLDEN / "raymarch_32.npz")
    res = make_res(0)
    world = synth.identity_world()
    for l in synth.LIGHTS[:2]:
        URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world)
    cam = synth.benchmark_camera(48, 32, jitter=bool(jitter), frame=3)
    rgba, steps = URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, 64.0)
    assert steps == int(want[f"steps_jitter{jitter}"][0])
    assert np.array_equal(rgba, want[f"rgba_jitter{jitter}"])
    assert np.array_equal(URaymarchUtils.PerformRaymarchCubeSetup(res, cam, world), want[f"setup_jitter{jitter}"])


def test_mandelbulb_close_to_golden():
    want = np.load(GOLDEN / "mandelbulb_32x24.npz")
    cam = synth.benchmark_camera(32, 24, jitter=False)
    out, iters = URaymarchUtils.PerformMandelbulbRaymarchReturnDistance(FMandelbulbParameters(MaxSteps=64.0, MaxIterations=8.0), cam,
                                                                        synth.identity_world())
    # libm vs CUDA transcendentals: 1e-4 per channel, up to 2 % of the pixels may land one sphere-tracing step earlier / later
    bad = np.abs(out - want["out"]).max(axis=-1) > 1e-4
    assert bad.mean() <= 0.02, f"{bad.sum()} of {bad.size} pixels differ"
    assert abs(iters - int(want["iterations"][0])) / int(want["iterations"][0]) < 0.02

