"""Device-side timing of the two fast raymarch kernels (device output, no step counter): reserved[1] = 2 (first generation)
vs 3 (second generation)."""
import ctypes as C
import sys

sys.path.insert(0, '.')
import torch

from tbraymarcherplugin_b200 import FMT_G8, _capi, synth
from tbraymarcherplugin_b200.raymarch_utils import FWindowingParameters, URaymarchUtils

lib = _capi.load()
for n, view, steps in [(512, (1920, 1080), 512.0), (256, (512, 512), 256.0)]:
    world = synth.identity_world()
    res = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=True)
    d = torch.empty(n * n * n, dtype=torch.uint8, device='cuda')
    _capi.check(lib.tbrm_synth_volume_u8(0, 1, (C.c_int32 * 3)(n, n, n), synth.PERLIN_SEED, C.c_void_p(d.data_ptr()), 1))
    URaymarchUtils.SetDataVolumeDevice(res, d.data_ptr())
    URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(res, FWindowingParameters(0.45, 0.5, True, False))
    for l in synth.LIGHTS[:2]:
        URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True)
    cam = synth.benchmark_camera(*view)
    out = torch.empty(view[0] * view[1] * 4, dtype=torch.float32, device='cuda')
    for kernel in (2, 3, 2, 3):
        URaymarchUtils.SetOptions(res, sweep_impl=2, debug_flags=(0, kernel))
        URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, steps, device_out_ptr=out.data_ptr(), count_steps=False)
        ms = C.c_float()
        lib.tbrm_timer_begin(res.handle)
        for _ in range(5):
            URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, steps, device_out_ptr=out.data_ptr(), count_steps=False)
        lib.tbrm_timer_end(res.handle, C.byref(ms))
        print(f"n={n} {view}: kernel generation {kernel - 1}: {ms.value / 5:.3f} ms", flush=True)
    res.release()
