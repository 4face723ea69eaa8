"""Writes the committed vectors of tests/golden/ from the CPU oracle (oracle/tbrm_oracle.cpp).

    python tests/golden/make_golden.py

The reference (UE 5.4 / HLSL) cannot run here and ships no golden vectors, so these are NOT reference
output: they freeze the oracle (any drift of the restatement, of the compiler flags or of the synthetic inputs shows up as
a diff in test_golden_cpu.py) and give the CUDA path a committed target next to the live oracle comparison
(test_gpu_golden.py). Inputs are the deterministic generators of tbraymarcherplugin_b200/synth.py."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE.parent))

import oracle  # noqa: E402
from tbraymarcherplugin_b200 import synth  # noqa: E402
from tbraymarcherplugin_b200.raymarch_utils import FMandelbulbParameters, FWindowingParameters  # noqa: E402

CT_WINDOW = FWindowingParameters(0.45, 0.5, True, False)
SWEEP_DIMS = (32, 32, 32)
PLAN_DIMS = (64, 48, 80)
WORLDS = {"identity": synth.identity_world, "scaled_rotated": synth.scaled_rotated_world, "clipped": synth.clipped_world}


def sweep_case():
    """Light volume after a full reset with the four benchmark lights, after removing L2, and after one ChangeDirLight."""
    data = synth.perlin_ct_volume(SWEEP_DIMS)
    out = {}
    for name, mk in WORLDS.items():
        ora = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), CT_WINDOW)
        world = mk()
        for l in synth.LIGHTS:
            ora.add_dir_light(l, True, world)
        out[f"{name}_reset"] = ora.light.copy()
        ora.add_dir_light(synth.LIGHTS[1], False, world)
        out[f"{name}_removed"] = ora.light.copy()
        ora.change_dir_light(synth.LIGHTS[0], synth.rotate_about_z(synth.LIGHTS[0], 5.0), world)
        out[f"{name}_changed"] = ora.light.copy()
    return out


def raymarch_case():
    data = synth.perlin_ct_volume(SWEEP_DIMS)
    ora = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), CT_WINDOW)
    world = synth.identity_world()
    for l in synth.LIGHTS[:2]:
        ora.add_dir_light(l, True, world)
    out = {}
    for jitter in (0, 1):
        cam = synth.benchmark_camera(48, 32, jitter=bool(jitter), frame=3)
        rgba, steps = ora.raymarch_lit(cam, world, 64.0)
        out[f"rgba_jitter{jitter}"] = rgba
        out[f"steps_jitter{jitter}"] = np.array([steps], np.int64)
        out[f"setup_jitter{jitter}"] = oracle.cube_setup(cam, world)
    return out


def plan_case():
    out = {}
    for name, mk in WORLDS.items():
        for i, l in enumerate(synth.LIGHTS):
            p = oracle.plan_dir_light(PLAN_DIMS, CT_WINDOW, l, mk())
            out[f"{name}_L{i}"] = np.frombuffer(bytes(p), np.uint8).copy()
    return out


def pow_case():
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(0.0, 1.0, 384), np.array([0.0, 1.0, 1e-30, 0.5, 0.999999, 1e-6])]).astype(np.float32)
    y = np.concatenate([rng.uniform(0.01, 4.0, 384), np.array([0.2, 0.2, 0.3, 1.0, 100.0, 0.05])]).astype(np.float32)
    r = np.array([oracle.det_pow(float(a), float(b)) for a, b in zip(x, y)], np.float32)
    return {"x": x, "y": y, "pow": r}


def mandelbulb_case():
    cam = synth.benchmark_camera(32, 24, jitter=False)
    params = FMandelbulbParameters(MaxSteps=64.0, MaxIterations=8.0)
    out, iters = oracle.mandelbulb(params, cam, synth.identity_world())
    return {"out": out, "iterations": np.array([iters], np.int64)}


CASES = {"sweep_32": sweep_case, "raymarch_32": raymarch_case, "plans": plan_case, "det_pow": pow_case, "mandelbulb_32x24": mandelbulb_case}

if __name__ == "__main__":
    for name, fn in CASES.items():
        arrays = fn()
        np.savez_compressed(HERE / f"{name}.npz", **arrays)
        print(name, {k: v.shape for k, v in arrays.items()}, f"{(HERE / (name + '.npz')).stat().st_size / 1024:.0f} KiB")
