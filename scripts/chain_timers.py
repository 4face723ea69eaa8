"""Diagnostic (TBRM_CHAIN_TIMERS builds): per-section cycle sums of sweep_chain_kernel for one light at N^3.
    TBRM_EXTRA_NVCC_FLAGS=-DTBRM_CHAIN_TIMERS python -m tbraymarcherplugin_b200.build --force; python scripts/chain_timers.py 512 3"""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from tbraymarcherplugin_b200 import _capi, synth, FMT_G8
from tbraymarcherplugin_b200.raymarch_utils import *
lib = _capi.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
li = int(sys.argv[2]) if len(sys.argv) > 2 else 3
res = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=True)
d = torch.empty(n * n * n, dtype=torch.uint8, device='cuda')
_capi.check(lib.tbrm_synth_volume_u8(0, 1, (C.c_int32 * 3)(n, n, n), synth.PERLIN_SEED, C.c_void_p(d.data_ptr()), 1))
URaymarchUtils.SetDataVolumeDevice(res, d.data_ptr())
URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
URaymarchUtils.SetWindowingParameters(res, FWindowingParameters(0.45, 0.5, True, False))
w = synth.identity_world()
for it in range(3):
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    URaymarchUtils.AddDirLightToSingleVolume(res, synth.LIGHTS[li], True, w, bGPUSync=True)
    URaymarchUtils.FlushRenderingCommands(res)
buf = np.zeros(4096 * 4 * 8, dtype=np.int64)
_capi.check(lib.tbrm_debug_download_derived(res.handle, 3, buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.nbytes)))
t = buf.reshape(4096, 4, 8)
used = t[(t.sum(axis=(1, 2)) > 0)]
names = ["a: halo request, T load", "c: halo check / park, probe", "barrier", "thread-0 work (flag, TMA issue)", "d: taps, lerp, forward, export", "block end: light brick update", "block head: T wait", "-"]
print("tiles with data:", used.shape[0], " slices:", n)
for wi in range(4):
    print(f"warp {wi}: " + ", ".join(f"{names[i].split(':')[0]}={used[:, wi, i].mean() / n:7.1f}" for i in range(7)) + f"  | total/slice {used[:, wi, :7].sum(axis=1).mean() / n:7.1f} cycles")
# the pacing tiles: the ones that wait least for their halo (everybody else shows slack as halo wait)
c0 = used[:, :3, 1].max(axis=1) / n
order = np.argsort(c0)
idx = np.nonzero(t.sum(axis=(1, 2)) > 0)[0]
print("tiles that wait least in (c) [tile: per-warp a c bar svc d end head]:")
for o in order[:6]:
    print(f"  tile {idx[o]:4d} (row {idx[o] // (n // 64)}, col {idx[o] % (n // 64)}):", " | ".join(" ".join(f"{used[o, wi, i] / n:6.0f}" for i in range(7)) for wi in range(4)))
print("c (max over warps 0-2) percentiles:", np.percentile(c0, [0, 5, 25, 50, 75, 95, 100]).round(0))
g = np.zeros((n // 8, n // 64))
tot = np.zeros((n // 8, n // 64))
for o in range(used.shape[0]):
    r, c = idx[o] // (n // 64), idx[o] % (n // 64)
    g[r, c] = used[o, :3, 1].max() / n
    tot[r, c] = used[o, 0, :7].sum() / n
print("halo wait (c, max over warps 0-2) per tile, cycles per slice; rows = tile rows (every 4th), cols = tile columns")
for r in range(0, n // 8, 4):
    print(f"  row {r:2d}: " + " ".join(f"{g[r, c]:5.0f}" for c in range(n // 64)) + "   | lifetime/slice: " + " ".join(f"{tot[r, c]:5.0f}" for c in range(n // 64)))
