"""The CUDA path (through the C ABI) against outputs of the REFERENCE'S OWN SHADERS: tests/golden/ref_shaders.npz was produced by
AddDirLightShader.usf / ChangeDirLightShader.usf / WindowedRaymarchMaterials.usf compiled for the CPU from /root/reference
(oracle/ref.mk, tests/golden/make_golden_ref.py; tests/test_ref_shaders_cpu.py holds the oracle to the same vectors). Bit-exact."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

from tbraymarcherplugin_b200 import FMT_G8, synth
from tbraymarcherplugin_b200.raymarch_utils import URaymarchUtils

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
_rspec = importlib.util.spec_from_file_location("make_golden_ref", GOLDEN / "make_golden_ref.py")
make_golden_ref = importlib.util.module_from_spec(_rspec)
_rspec.loader.exec_module(make_golden_ref)


class _GpuVolume:
    """Adapter with the method names make_golden_ref.shader_sequence drives (the oracle / reference volume interface)."""

    def __init__(self, light32, gpu_sync):
        data, _, win = make_golden_ref.shader_inputs()
        Z, Y, X = data.shape
        self.res = URaymarchUtils.InitializeRaymarchResources((X, Y, Z), FMT_G8, bLightVolume32Bit=light32)
        URaymarchUtils.SetDataVolume(self.res, data)
        URaymarchUtils.ColorCurveToTexture(self.res, synth.soft_ct_curve())
        URaymarchUtils.SetWindowingParameters(self.res, win)
        URaymarchUtils.ClearResourceLightVolumes(self.res, 0.0)
        self.sync = gpu_sync

    @property
    def light(self):
        return URaymarchUtils.ReadLightVolume(self.res)

    def add_dir_light(self, light, added, world):
        assert URaymarchUtils.AddDirLightToSingleVolume(self.res, light, added, world, bGPUSync=self.sync)

    def change_dir_light(self, old, new, world):
        assert URaymarchUtils.ChangeDirLightInSingleVolume(self.res, old, new, world, bGPUSync=self.sync)


@pytest.mark.parametrize("gpu_sync", [False, True])
@pytest.mark.parametrize("light32", [True, False])
@pytest.mark.parametrize("world_name", list(make_golden_ref.SHADER_WORLDS))
def test_cuda_path_equals_the_reference_shaders_golden_outputs(world_name, light32, gpu_sync):
    """Full reset, removal, in-place change, fallback change and a lit frame: the CUDA path against what the reference's own
    AddDirLightShader.usf / ChangeDirLightShader.usf / WindowedRaymarchMaterials.usf produced for the same inputs."""
    want = np.load(GOLDEN / "ref_shaders.npz")
    vol = _GpuVolume(light32, gpu_sync)
    world = make_golden_ref.SHADER_WORLDS[world_name]()
    tag = f"{world_name}_{'r32f' if light32 else 'g8'}"
    got = {}
    make_golden_ref.shader_sequence(vol, world, got, tag)
    for k, v in got.items():
        assert v.dtype == want[k].dtype and np.array_equal(v, want[k]), f"{k}: CUDA path differs from the reference shader"
    if light32:
        cam = synth.benchmark_camera(40, 24, jitter=True, frame=3)
        rgba, _ = URaymarchUtils.PerformWindowedLitRaymarch(vol.res, cam, world, 48.0)
        assert np.array_equal(rgba, want[f"{tag}_lit"])
        assert np.array_equal(URaymarchUtils.PerformRaymarchCubeSetup(vol.res, cam, world), want[f"{tag}_setup"])
