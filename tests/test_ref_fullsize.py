"""Parity at BASELINE.json's FULL sizes, against the reference's own code. tests/golden/ref_fullsize_hashes.json holds SHA-256 checksums
(per block of 64 Z-slices / image rows, and the checksum of the checksums) of the light volume and the frame that the REFERENCE'S OWN
shaders produce (compiled for the CPU, oracle/ref.mk; tests/golden/make_golden_ref_fullsize.py) for
  cfg1 = configs[0]: 256^3 sphere R8, 1 light, 512 x 512, 256 steps;
  cfg2 = configs[1]: 512^3 CT-like Perlin R8, 2 lights, 1920 x 1080, 512 steps, windowing on — the bench.py workload;
  cfg3 = configs[2]: 512^3, 4 lights, 16 incremental ChangeDirLight updates (light volume after the reset and after updates 4, 8, 16);
  cfg4 = configs[3]: 1024^3, 3 lights, 3840 x 2160, 768 steps (GPU only; one GPU — N GPUs give the same bits, tests/test_gpu_multi.py).
CPU: the oracle reproduces them (bit-exact). GPU: the CUDA path (TMA-staged fused sweep + fast lit march, through the C ABI) reproduces
them — at the size the benchmark runs."""
import ctypes as C
import importlib.util
import json
from pathlib import Path

import numpy as np
import pytest

import oracle
from tbraymarcherplugin_b200 import FMT_G8, _capi, synth
from tbraymarcherplugin_b200.raymarch_utils import FSweepStats, FWindowingParameters, URaymarchUtils

GOLDEN = Path(__file__).resolve().parent / "golden"
_spec = importlib.util.spec_from_file_location("make_golden_ref_fullsize", GOLDEN / "make_golden_ref_fullsize.py")
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)
WANT = json.loads((GOLDEN / "ref_fullsize_hashes.json").read_text())


def assert_digests(got: np.ndarray, want: dict, what: str):
    d = mk.digests(got)
    assert d["shape"] == want["shape"] and d["dtype"] == want["dtype"], what
    bad = [i for i, (a, b) in enumerate(zip(d["blocks"], want["blocks"])) if a != b]
    assert not bad, f"{what}: blocks {bad} of {len(want['blocks'])} (64 slices / rows each) differ from the reference shaders' output"
    assert d["all"] == want["all"], what


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_oracle_equals_the_reference_shaders_at_full_size(name):
    light, frame = mk.run(mk.CONFIGS[name], oracle.OracleVolume, lambda v, cam, w, s: v.raymarch_lit(cam, w, s)[0])
    assert_digests(light, WANT[name]["light"], f"{name} light volume")
    assert_digests(frame, WANT[name]["frame"], f"{name} frame")
    assert abs(float(light.max()) - WANT[name]["light_max"]) == 0 and 0.05 < WANT[name]["frame_alpha_mean"] < 0.95  # not a trivial scene


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg4"])
def test_cuda_path_equals_the_reference_shaders_at_full_size(name):
    if name not in WANT:
        pytest.skip(f"ref_fullsize_hashes.json has no {name} entry")
    cfg = mk.CFG4 if name == "cfg4" else mk.CONFIGS[name]
    n = cfg["n"]
    res = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=True)
    data = oracle.synth_volume(cfg["volume"], (n, n, n))  # bit-identical to the device generator (test_gpu_parity.py)
    URaymarchUtils.SetDataVolume(res, data)
    URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(res, FWindowingParameters(*cfg["window"]))
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    world = synth.identity_world()
    impls = set()
    for i in cfg["lights"]:
        st = FSweepStats()
        assert URaymarchUtils.AddDirLightToSingleVolume(res, synth.LIGHTS[i], True, world, bGPUSync=True, stats=st)
        impls |= set(st.impl)
    if name == "cfg2":  # the bench.py workload: the dominant kernel is the one under test
        assert impls == {3}, f"the TMA-staged fused sweep must have run every pass: {impls}"
    assert_digests(URaymarchUtils.ReadLightVolume(res), WANT[name]["light"], f"{name} light volume")
    frame, steps = URaymarchUtils.PerformWindowedLitRaymarch(res, synth.benchmark_camera(*cfg["view"]), world, cfg["steps"])
    assert_digests(frame, WANT[name]["frame"], f"{name} frame")
    assert steps > 1e7
    res.release()


@pytest.mark.gpu
def test_cuda_change_dir_light_sequence_equals_the_reference_shaders_at_full_size():
    """BASELINE.json configs[2]: 512^3, 4 lights, 16 incremental ChangeDirLight updates (update k turns light k % 4 by a further 5 degrees
    about +Z): the light volume after the reset and after updates 4, 8 and 16 against the reference's ChangeDirLightShader.usf output."""
    if "cfg3" not in WANT:
        pytest.skip("ref_fullsize_hashes.json has no cfg3 entry")
    cfg = mk.CFG3
    n = cfg["n"]
    res = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=True)
    URaymarchUtils.SetDataVolume(res, oracle.synth_volume(cfg["volume"], (n, n, n)))
    URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(res, FWindowingParameters(*cfg["window"]))
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)

    class Gpu:  # the volume interface mk.run_cfg3 drives
        def __init__(self, *a):
            pass

        @property
        def light(self):
            return URaymarchUtils.ReadLightVolume(res)

        def add_dir_light(self, light, added, world):
            assert URaymarchUtils.AddDirLightToSingleVolume(res, light, added, world, bGPUSync=True)

        def change_dir_light(self, old, new, world):
            assert URaymarchUtils.ChangeDirLightInSingleVolume(res, old, new, world, bGPUSync=True)

    seen = []
    real_inputs = mk.inputs
    mk.inputs = lambda cfg: (None, None, None)  # the data volume already sits on the GPU
    try:
        mk.run_cfg3(Gpu, lambda k, light: (seen.append(k), assert_digests(light, WANT["cfg3"]["light"][str(k)], f"cfg3 after update {k}")))
    finally:
        mk.inputs = real_inputs
    assert tuple(seen) == tuple(cfg["checkpoints"])
    res.release()
