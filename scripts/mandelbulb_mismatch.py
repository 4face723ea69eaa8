"""Measured mismatch of the Mandelbulb march against the CPU oracle (SDFMarcher.usf:24-112 restated, libm transcendentals):
fraction of pixels whose (distance, hit) differ by more than 1e-4, for the default Power == 8 kernel (transcendental-free iteration) and for
the trigonometric path (TBRM_MANDELBULB_TRIG=1 in a second process), plus timing. Writes gpurun_out/mandelbulb_mismatch.json lines."""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import oracle
from tbraymarcherplugin_b200 import synth
from tbraymarcherplugin_b200.raymarch_utils import FMandelbulbParameters, URaymarchUtils

world = synth.identity_world()
cases = {"test_small": ((96, 64), 256.0, 12.0), "test_twin": ((240, 135), 256.0, 16.0), "cfg5": ((1920, 1080), 1024.0, 16.0)}
out = {"path": "trig" if os.environ.get("TBRM_MANDELBULB_TRIG") == "1" else "p8"}
for name, (view, steps, iters) in cases.items():
    cam = synth.benchmark_camera(*view, jitter=False)
    mb = FMandelbulbParameters(MaxSteps=steps, MaxIterations=iters)
    got, n_it = URaymarchUtils.PerformMandelbulbRaymarchReturnDistance(mb, cam, world)
    t0 = time.perf_counter()
    got, n_it = URaymarchUtils.PerformMandelbulbRaymarchReturnDistance(mb, cam, world)
    gpu_ms = 1e3 * (time.perf_counter() - t0)
    ref, ref_it = oracle.mandelbulb(mb, cam, world)
    L = oracle.lib()
    try:
        L.tbo_set_mandelbulb_variant(1)
        twin, twin_it = oracle.mandelbulb(mb, cam, world)
    finally:
        L.tbo_set_mandelbulb_variant(0)
    out[name] = {"pixels": int(got.shape[0] * got.shape[1]), "vs_reference_formulation": float((np.abs(got - ref).max(-1) > 1e-4).mean()),
                 "vs_p8_twin": float((np.abs(got - twin).max(-1) > 1e-4).mean()), "iterations": int(n_it), "ref_iterations": int(ref_it),
                 "host_call_ms_incl_download": gpu_ms}
print(json.dumps(out))
