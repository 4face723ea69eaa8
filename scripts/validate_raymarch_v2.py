"""One-shot validation of raymarch_fast2_kernel (second-generation lit ray march) and of the default kernel's 32-bit tap addressing (ADDR32)
against the first-generation fast kernel with 64-bit addressing and the generic kernel: frames and executed-step counts must be
bit-identical; prints timings. Exit code 0 = identical."""
import ctypes as C
import sys

sys.path.insert(0, '.')
import numpy as np
import torch

from tbraymarcherplugin_b200 import FMT_G8, _capi, synth
from tbraymarcherplugin_b200.raymarch_utils import FWindowingParameters, URaymarchUtils

lib = _capi.load()
ok = True
LOG = []


def frame(res, cam, world, steps, kernel):
    URaymarchUtils.SetOptions(res, sweep_impl=2, debug_flags=(0, kernel))
    best, out = 1e9, None
    for _ in range(3):
        ms = C.c_float()
        lib.tbrm_timer_begin(res.handle)
        out = URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, steps)
        lib.tbrm_timer_end(res.handle, C.byref(ms))
        best = min(best, ms.value)
    return out[0], out[1], best


for n, view, steps, world_name, half in [(512, (1920, 1080), 512.0, "identity", False), (256, (640, 400), 256.0, "clipped", False),
                                         (256, (640, 400), 256.0, "scaled_rotated", False), (96, (320, 200), 128.0, "identity", True)]:
    world = {"identity": synth.identity_world, "clipped": synth.clipped_world, "scaled_rotated": synth.scaled_rotated_world}[world_name]()
    res = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=True, LightVolumeHalfResolution=half)
    d = torch.empty(n * n * n, dtype=torch.uint8, device='cuda')
    _capi.check(lib.tbrm_synth_volume_u8(0, 1, (C.c_int32 * 3)(n, n, n), synth.PERLIN_SEED, C.c_void_p(d.data_ptr()), 1))
    URaymarchUtils.SetDataVolumeDevice(res, d.data_ptr())
    URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(res, FWindowingParameters(0.45, 0.5, True, False))
    URaymarchUtils.SetOptions(res, sweep_impl=0)
    for l in synth.LIGHTS[:2]:
        URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=not half)
    cam = synth.benchmark_camera(*view)
    f1, s1, t1 = frame(res, cam, world, steps, 2)   # first-generation fast kernel, 64-bit tap addressing (what round 1 measured)
    f2, s2, t2 = frame(res, cam, world, steps, 3)
    f0, s0, t0 = frame(res, cam, world, steps, 0)   # the default: the same kernel with 32-bit / IMAD tap addressing (ADDR32)
    same = np.array_equal(f1, f2) and s1 == s2 and np.array_equal(f1, f0) and s1 == s0
    msg = f"n={n} {view} {world_name} half={half}: v1 {t1:.3f} ms, v2 {t2:.3f} ms, default (ADDR32) {t0:.3f} ms, steps {s1} / {s2} / {s0}, identical={same}"
    if n <= 256:
        fg, sg, tg = frame(res, cam, world, steps, 1)
        same_g = np.array_equal(fg, f2) and sg == s2
        msg += f", generic {tg:.3f} ms identical={same_g}"
        same = same and same_g
    if not same:
        dd = np.abs(f1 - f2)
        msg += f" MAXDIFF {dd.max():.3e} at {np.argwhere(dd > 0)[:3].tolist()}"
    print(msg, flush=True)
    LOG.append(msg)
    ok = ok and same
    res.release()
print("V2 OK" if ok else "V2 MISMATCH", flush=True)
try:  # keep the timings where a gpurun call brings them back
    import os
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/raymarch_ab.txt", "w") as fh:
            fh.write("\n".join(LOG) + "\n")
except OSError:
    pass
sys.exit(0 if ok else 1)
