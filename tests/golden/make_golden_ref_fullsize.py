"""Writes tests/golden/ref_fullsize_hashes.json: SHA-256 checksums of what the REFERENCE'S OWN shaders (oracle/_ref/libtbrm_ref.so, see
make_golden_ref.py) produce at BASELINE.json's FULL sizes — light volume and frame of configs[0] (256^3 sphere, 1 light, 512x512, 256
steps) and configs[1] (512^3 CT-like Perlin, 2 lights, 1080p, 512 steps, windowing on). The arrays are far too large to commit
(0.5 GiB light volume); their checksums are not: one digest per block of 64 Z-slices / 64 image rows (to localise a mismatch) and the
digest of the digests. tests/test_ref_fullsize.py holds the oracle (cfg1, CPU) and the CUDA path (cfg1 + cfg2, GPU) to them.

    python tests/golden/make_golden_ref_fullsize.py        (a few minutes of CPU; needs oracle/_ref)
"""
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE.parent))

import oracle  # noqa: E402
from tbraymarcherplugin_b200 import synth  # noqa: E402
from tbraymarcherplugin_b200.raymarch_utils import FWindowingParameters  # noqa: E402

BLOCK = 64

# BASELINE.json configs[3] on ONE GPU's worth of work: 1024^3, 3 lights, 3840 x 2160, 768 steps (about 20 minutes of CPU for the reference build:
# generated only with --cfg4). The 8-GPU slab-sharded run must give the same bits (tests/test_gpu_multi.py checks N GPUs against 1).
CFG4 = dict(volume="perlin", n=1024, lights=[0, 1, 2], view=(3840, 2160), steps=768.0, window=(0.45, 0.5, True, False))

CONFIGS = {
    # BASELINE.json configs[0]; TF soft_ct so that rays do not saturate in one step, default windowing
    "cfg1": dict(volume="sphere", n=256, lights=[0], view=(512, 512), steps=256.0, window=(0.5, 1.0, True, True)),
    # BASELINE.json configs[1] = the bench.py workload
    "cfg2": dict(volume="perlin", n=512, lights=[0, 1], view=(1920, 1080), steps=512.0, window=(0.45, 0.5, True, False)),
}


def digests(a: np.ndarray) -> dict:
    """SHA-256 of every block of BLOCK leading-axis slices of a C-contiguous array, and of the concatenated block digests."""
    a = np.ascontiguousarray(a)
    blocks = [hashlib.sha256(a[i:i + BLOCK].tobytes()).hexdigest() for i in range(0, a.shape[0], BLOCK)]
    return {"shape": list(a.shape), "dtype": str(a.dtype), "blocks": blocks, "all": hashlib.sha256("".join(blocks).encode()).hexdigest()}


def inputs(cfg):
    n = cfg["n"]
    data = oracle.synth_volume(cfg["volume"], (n, n, n))
    return data, oracle.prepare_tf(synth.soft_ct_curve()), FWindowingParameters(*cfg["window"])


def run(cfg, Volume, march):
    """Full reset with cfg's lights, then the frame. Volume: the oracle's or the reference's volume class; march(vol, cam, world, steps)."""
    data, tf, win = inputs(cfg)
    vol = Volume(data, tf, win)
    world = synth.identity_world()
    for i in cfg["lights"]:
        vol.add_dir_light(synth.LIGHTS[i], True, world)
    cam = synth.benchmark_camera(*cfg["view"])
    return vol.light, march(vol, cam, world, cfg["steps"])


# BASELINE.json configs[2]: 512^3, 4 lights, incremental ChangeDirLight updates. Update k turns light k % 4 by a further 5 degrees about +Z
#  through ChangeDirLightInSingleVolume; checkpoints after the reset and after updates 4, 8 and 16.
CFG3 = dict(volume="perlin", n=512, window=(0.45, 0.5, True, False), updates=16, checkpoints=(0, 4, 8, 16))


def run_cfg3(Volume, on_checkpoint):
    data, tf, win = inputs(CFG3)
    vol = Volume(data, tf, win)
    world = synth.identity_world()
    lights = list(synth.LIGHTS)
    for l in lights:
        vol.add_dir_light(l, True, world)
    on_checkpoint(0, vol.light)
    for k in range(1, CFG3["updates"] + 1):
        i = (k - 1) % 4
        new = synth.rotate_about_z(lights[i], 5.0)
        vol.change_dir_light(lights[i], new, world)
        lights[i] = new
        if k in CFG3["checkpoints"]:
            on_checkpoint(k, vol.light)


if __name__ == "__main__":
    import refpin

    path = HERE / "ref_fullsize_hashes.json"
    if "--cfg3" in sys.argv:  # adds / refreshes the cfg3 entry only (about 7 minutes on 8 cores)
        out = json.loads(path.read_text())
        entry = {"config": {k: (list(v) if isinstance(v, tuple) else v) for k, v in CFG3.items()}, "light": {}}
        t0 = time.time()

        def keep(k, light):
            entry["light"][str(k)] = digests(light)
            print("cfg3 checkpoint", k, "%.0f s" % (time.time() - t0), entry["light"][str(k)]["all"][:16], float(light.max()), flush=True)

        run_cfg3(refpin.RefVolume, keep)
        out["cfg3"] = entry
        path.write_text(json.dumps(out, indent=1) + "\n")
        sys.exit(0)
    out = json.loads(path.read_text()) if path.exists() else {}
    todo = {"cfg4": CFG4} if "--cfg4" in sys.argv else CONFIGS
    for name, cfg in todo.items():
        t0 = time.time()
        light, frame = run(cfg, refpin.RefVolume, lambda v, cam, w, s: v.raymarch(0, cam, w, s))
        out[name] = {"config": {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}, "light": digests(light), "frame": digests(frame),
                     "light_max": float(light.max()), "frame_alpha_mean": float(frame[..., 3].mean())}
        print(name, "reference shaders: %.0f s" % (time.time() - t0), out[name]["light"]["all"][:16], out[name]["frame"]["all"][:16], flush=True)
        if "--check-oracle" in sys.argv:
            t0 = time.time()
            light_o, frame_o = run(cfg, oracle.OracleVolume, lambda v, cam, w, s: v.raymarch_lit(cam, w, s)[0])
            print(name, "oracle: %.0f s" % (time.time() - t0), "light equal", digests(light_o) == out[name]["light"], "frame equal",
                  digests(frame_o) == out[name]["frame"], flush=True)
    path.write_text(json.dumps(out, indent=1) + "\n")
