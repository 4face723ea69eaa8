"""Times the fused illumination sweep and the lit march on one GPU for the environment it runs in — run it once per setting to A/B the
switches: `TBRM_SWEEP_PX=1|2|auto`, `TBRM_RAYMARCH_ADDR64=1`, `TBRM_RAYMARCH_V2=1`.

    python scripts/time_sweep_ab.py [N=256] [view width=1920] [view height=1080] [steps=512]

Torch-free (libcudart through ctypes, CUDA events on the resource set's stream via tbrm_timer_begin / tbrm_timer_end). Prints one JSON line:
the minimum and the median over the repetitions of one full reset of two lights (4 axis passes) and of one frame."""
import ctypes as C
import json
import os
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tbraymarcherplugin_b200 import FMT_G8, _capi, synth  # noqa: E402
from tbraymarcherplugin_b200.raymarch_utils import FSweepStats, FWindowingParameters, URaymarchUtils  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
steps = float(sys.argv[4]) if len(sys.argv) > 4 else 512.0
lib = _capi.load()
rt = None
for name in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        rt = C.CDLL(name)
        break
    except OSError:
        pass
assert rt is not None, "libcudart not found"


def dmalloc(nbytes):
    p = C.c_void_p()
    assert rt.cudaMalloc(C.byref(p), C.c_size_t(nbytes)) == 0
    return p


d_vol = dmalloc(N * N * N)
_capi.check(lib.tbrm_synth_volume_u8(0, 1, (C.c_int32 * 3)(N, N, N), synth.PERLIN_SEED & 0xFFFFFFFF, d_vol, 1))
res = URaymarchUtils.InitializeRaymarchResources((N, N, N), FMT_G8, bLightVolume32Bit=True)
URaymarchUtils.SetDataVolumeDevice(res, d_vol.value)
URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
URaymarchUtils.SetWindowingParameters(res, FWindowingParameters(0.45, 0.5, True, False))
world = synth.identity_world()
cam = synth.benchmark_camera(W, H)
d_frame = dmalloc(W * H * 16)


def timed(fn, reps):
    fn()
    ms = []
    for _ in range(reps):
        t = C.c_float()
        _capi.check(lib.tbrm_timer_begin(res.handle))
        fn()
        _capi.check(lib.tbrm_timer_end(res.handle, C.byref(t)))
        ms.append(t.value)
    return {"ms_min": min(ms), "ms_median": statistics.median(ms), "reps": reps}


impl = []


def reset():
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    for l in synth.LIGHTS[:2]:
        st = FSweepStats()
        URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True, stats=st)
        impl.append(list(st.impl))


def frame():
    c, w = cam.to_c(), world.to_c()
    _capi.check(lib.tbrm_raymarch_lit(res.handle, C.byref(c), C.byref(w), steps, 0, H, d_frame, 1, None))


out = {"volume": N, "view": [W, H], "steps": steps,
       "env": {k: os.environ.get(k) for k in ("TBRM_SWEEP_GEN", "TBRM_SWEEP_PX", "TBRM_RAYMARCH_ADDR64", "TBRM_RAYMARCH_V2")},
       "reset_2_lights": timed(reset, 10), "frame": timed(frame, 10), "sweep_impl": impl[-2:]}
print(json.dumps(out), flush=True)
