"""What compute-sanitizer runs (mandatory for kernels with cross-CTA dependencies): a 64^3 sweep with the three schedules
(per-slice, generic fused, TMA-staged fused: both kernel generations, clip plane on), a ChangeDirLight, and one small lit frame.
    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_target.py"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
from tbraymarcherplugin_b200 import FMT_G8, synth
from tbraymarcherplugin_b200.raymarch_utils import FSweepStats, FWindowingParameters, URaymarchUtils

n = 64
data = synth.perlin_ct_volume((n, n, n))
win = FWindowingParameters(0.45, 0.5, True, False)
results = {}
for world_name, world in (("identity", synth.identity_world()), ("clipped", synth.clipped_world())):
    # reserved[0] bit 6: second kernel generation; bits 8-9 = 2 with bits 4-5 = 2: tiles of 7 rows, two pixels per thread
    for impl, sync, bits in ((1, False, 0), (3, True, 0), (2, True, 0), (2, True, 64), (2, True, 512 + 32)):
        res = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=True)
        URaymarchUtils.SetDataVolume(res, data)
        URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
        URaymarchUtils.SetWindowingParameters(res, win)
        URaymarchUtils.SetOptions(res, sweep_impl=impl, debug_flags=(bits, 0, 0))
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
        for l in synth.LIGHTS[:3]:
            st = FSweepStats()
            assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=sync, stats=st)
        assert URaymarchUtils.ChangeDirLightInSingleVolume(res, synth.LIGHTS[0], synth.rotate_about_z(synth.LIGHTS[0], 5.0), world, bGPUSync=sync)
        results[(world_name, impl, bits)] = URaymarchUtils.ReadLightVolume(res)
        if impl == 2 and bits == 0:
            frame, steps = URaymarchUtils.PerformWindowedLitRaymarch(res, synth.benchmark_camera(96, 64), world, 64.0)
        res.release()
    ref = results[(world_name, 1, 0)]
    for k, v in results.items():
        if k[0] == world_name:
            assert np.array_equal(v, ref), k
# the other light-volume settings through the TMA-staged kernel: G8 (byte bricks, ChangeDirLight with the R32F scratch brick), half resolution
for light32, half in ((False, False), (True, True), (False, True)):
    out = []
    for impl in (3, 2):
        res = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=light32, LightVolumeHalfResolution=half)
        URaymarchUtils.SetDataVolume(res, data)
        URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
        URaymarchUtils.SetWindowingParameters(res, win)
        URaymarchUtils.SetOptions(res, sweep_impl=impl)
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
        world = synth.identity_world()
        for l in synth.LIGHTS[:3]:
            st = FSweepStats()
            assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True, stats=st)
            assert impl == 3 or 3 in st.impl, st.impl
        assert URaymarchUtils.ChangeDirLightInSingleVolume(res, synth.LIGHTS[0], synth.rotate_about_z(synth.LIGHTS[0], 5.0), world, bGPUSync=True)
        out.append(URaymarchUtils.ReadLightVolume(res))
        res.release()
    assert np.array_equal(out[0], out[1]), (light32, half)
print("sanitize target ok: schedules agree bit for bit", {k: float(v.max()) for k, v in results.items()})
