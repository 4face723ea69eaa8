#!/usr/bin/env python
"""bench.py — the hot path of TBRaymarcherPlugin on B200, measured on BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): 512^3 CT-like Perlin R8 volume, R32F light volume, windowing {C .45, W .5, low cut-off},
'soft_ct' transfer function, 2 directional lights, 1920x1080 view, 512 steps. One STEP is one pass of the hot path over
that input = what ARaymarchVolume::Tick does when the lights changed (RaymarchVolume.cpp:374-378, 418-451) plus the frame
the renderer then draws: ClearLightVolume + AddDirLightToSingleVolume x 2 (fused sweep) + PerformWindowedLitRaymarch.

Metric: Mray-steps/s = executed march-loop iterations of the frame (incl. clipped ones and the final partial step,
SURVEY.md §8d) / device time of the whole step (sweep + raymarch). The per-stage figures (raymarch Mray-steps/s on its own,
sweep Mvoxels/s and GB/s) ride along in "stages".

N > 1 (torchrun, one process per GPU): ONE volume is sharded over the GPUs as Z-slabs — in-kernel NVLink halo exchange for the sweep,
NCCL all-gathers of the data / light slabs, interleaved row blocks of the frame gathered on rank 0; strong scaling.
--sharding volumes gives every rank its own independent volume instead (no data-path collective, weak scaling).

--impl reference: the reference has no CPU implementation and cannot be built here (Unreal Engine 5.4 + HLSL), so this arm
times the reference's own shaders compiled for the CPU (oracle/_ref, kind "reference"; the CPU oracle, kind "port", where that build is
absent) with all host threads on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = "cfg2"
N_VOL = 512
VIEW = (1920, 1080)
STEPS = 512.0
LIGHT_IDS = (0, 1)
SAMPLE_ROWS = list(range(5, 1080, 10))  # 108 evenly spaced rows (10 % of the frame) for the CPU baseline's raymarch sample
SWEEP_SAMPLE_N = 512  # the CPU sweep sample runs ONE WHOLE axis pass over the workload's own volume (no voxel extrapolation)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


GOLDEN_HASHES = ROOT / "tests" / "golden" / "ref_fullsize_hashes.json"  # SHA-256 digests of what the REFERENCE'S OWN shaders produce at full size
HASH_BLOCK = 64  # leading-axis slices per digest block (tests/golden/make_golden_ref_fullsize.py)


def digests(a) -> dict:
    """SHA-256 of every block of 64 leading-axis slices of a C-contiguous array and of the concatenated block digests — the format of
    tests/golden/ref_fullsize_hashes.json."""
    import hashlib

    import numpy as np

    a = np.ascontiguousarray(a)
    blocks = [hashlib.sha256(a[i:i + HASH_BLOCK].tobytes()).hexdigest() for i in range(0, a.shape[0], HASH_BLOCK)]
    return {"shape": list(a.shape), "dtype": str(a.dtype), "blocks": blocks, "all": hashlib.sha256("".join(blocks).encode()).hexdigest()}


def parity_against_reference(cfg: str, light=None, frame=None) -> dict:
    """Bit-parity of this run's light volume / frame with the reference shaders' output for the same configuration (committed digests).
    Runs outside every timed region; what the driver sees of the multi-GPU exchange path's correctness."""
    out = {"against": f"tests/golden/ref_fullsize_hashes.json[{cfg}] (reference shaders compiled for the CPU)"}
    if not GOLDEN_HASHES.exists():
        return {**out, "available": False}
    want = json.loads(GOLDEN_HASHES.read_text()).get(cfg)
    if want is None:
        return {**out, "available": False}
    if light is not None:
        out["light"] = digests(light)["all"] == want["light"]["all"]
    if frame is not None and "frame" in want:
        out["frame"] = digests(frame)["all"] == want["frame"]["all"]
    return out


def workload_config(n_gpus: int, slabs: bool = True) -> dict:
    if n_gpus == 1:
        shard = "1 volume, 1 GPU"
    elif slabs:
        shard = (f"ONE volume Z-slab sharded over {n_gpus} GPUs: data volume replicated (NCCL all-gather of the uploaded slabs), light sweep per "
                 "slab with in-kernel NVLink peer-store halo exchange, light slabs gathered by NCCL all-gather (2 GPUs: pushed into the peer's "
                 "volume by TMA stores from the sweep's last pass instead, --push-gather), frame rendered in interleaved 8-row blocks and "
                 "gathered on rank 0")
    else:
        shard = f"{n_gpus} independent volumes, one per GPU, no collective"
    return {
        "workload": f"{'cfg2' if N_VOL == 512 else 'cfg4'}: {N_VOL}^3 CT-like Perlin R8 volume, R32F light volume, {len(LIGHT_IDS)} dir lights full reset (fused sweep, bGPUSync) + lit raymarch "
                    f"{VIEW[0]}x{VIEW[1]} @ {int(STEPS)} steps, windowing C=.45 W=.5 low cut-off, TF soft_ct, jitter on",
        "volume": [N_VOL] * 3, "view": list(VIEW), "steps": STEPS, "lights": len(LIGHT_IDS),
        "sharding": shard,
        "cache": "inputs larger than L2 (128 MiB data + 512 MiB light volume vs 126 MB L2): no flush needed between steps",
        # kernel selection switches that were set for this run (A/B timing; none changes a result — INTEGRATION.md)
        "switches": {k: os.environ[k] for k in ("TBRM_SWEEP_GEN", "TBRM_SWEEP_PX", "TBRM_RAYMARCH_ADDR64", "TBRM_RAYMARCH_V2") if k in os.environ},
    }


# ------------------------------------------------------------------------------------------------------------------
# CPU oracle legs (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------------------------
def reference_build_available() -> bool:
    """oracle/_ref/libtbrm_ref.so: the reference's own shaders and host math compiled for the CPU (oracle/ref.mk). It is built where
    /root/reference exists and travels to the GPU box as a file. True only if it is there AND loads (a file built for another machine
    must not take the bench line down: the oracle port stands in)."""
    if not (ROOT / "oracle" / "_ref" / "libtbrm_ref.so").exists():
        return False
    try:
        sys.path.insert(0, str(ROOT / "tests"))
        import refpin

        refpin.lib()
        return True
    except Exception as e:  # noqa: BLE001
        print(f"bench.py: oracle/_ref/libtbrm_ref.so does not load ({e!r}); falling back to the oracle port", file=sys.stderr)
        return False


def oracle_sample(data, data_small, light_after_reset, threads: int, kind: str = "port"):
    """One bounded sample of the workload on the host: one sweep axis pass over a 256^3 rendition of the volume (the per-voxel
    work does not depend on the resolution) and the lit raymarch of 20 evenly spaced rows of the real 512^3 / 1080p frame.
    kind "reference": the reference's own AddDirLightShader.usf / WindowedRaymarchMaterials.usf + LightingShaderUtils.cpp compiled for the
    CPU (oracle/_ref, OpenMP over the pixels of a slice / the pixels of a row block); kind "port": the oracle's restatement.
    Returns (estimated seconds for the whole step, estimated ray-steps of the whole frame, detail)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import numpy as np

    import oracle
    from tbraymarcherplugin_b200 import synth
    from tbraymarcherplugin_b200.raymarch_utils import FWindowingParameters

    oracle.lib().tbo_set_threads(threads)
    if kind == "reference":
        import refpin
        Volume = refpin.RefVolume
    else:
        Volume = oracle.OracleVolume
    win = FWindowingParameters(0.45, 0.5, True, False)
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    world = synth.identity_world()
    warm = Volume(np.ascontiguousarray(data_small[:16, :16, :16]), tf, win)  # untimed: spins up the OpenMP team, pages the library in
    warm.add_dir_light(synth.LIGHTS[3], True, world)
    small = Volume(data_small, tf, win)
    t0 = time.perf_counter()
    small.add_dir_light(synth.LIGHTS[3], True, world)  # axis-aligned light: exactly one axis pass over all voxels
    t_pass = (time.perf_counter() - t0) * (N_VOL / SWEEP_SAMPLE_N) ** 3
    total_passes = 2 * len(LIGHT_IDS)  # L1, L2 (and L3) take two axis passes each
    vol = Volume(data, tf, win)
    if light_after_reset is not None:
        vol.light = light_after_reset
    else:
        vol.light[...] = 1.0  # the cost of a march step does not depend on the light values
    cam = synth.benchmark_camera(*VIEW)
    t_rows, steps_rows = 0.0, 0
    for r in SAMPLE_ROWS:
        t0 = time.perf_counter()
        if kind == "reference":
            vol.raymarch(0, cam, world, STEPS, rows=(r, r + 1))
        else:
            _, st = vol.raymarch_lit(cam, world, STEPS, rows=(r, r + 1))
        t_rows += time.perf_counter() - t0
        if kind == "reference":  # the shader does not count its steps: the (bit-identical) oracle does, untimed
            _, st = oracle.OracleVolume.raymarch_lit(vol, cam, world, STEPS, rows=(r, r + 1))
        steps_rows += st
    scale = VIEW[1] / len(SAMPLE_ROWS)
    est_seconds = total_passes * t_pass + t_rows * scale
    est_steps = steps_rows * scale
    detail = {"estimated": True, "sweep_pass_s": t_pass, "sweep_voxel_scale": (N_VOL / SWEEP_SAMPLE_N) ** 3, "sweep_pass_scale": total_passes,
              "raymarch_rows_s": t_rows, "rows": len(SAMPLE_ROWS), "row_scale": scale, "row_steps": steps_rows}
    return est_seconds, est_steps, detail


def sample_text() -> str:
    vox = "the whole volume" if SWEEP_SAMPLE_N == N_VOL else f"a {SWEEP_SAMPLE_N}^3 rendition of the volume (x{(N_VOL // SWEEP_SAMPLE_N) ** 3} voxels)"
    return (f"ESTIMATE from a bounded sample, per step: 1 of the {2 * len(LIGHT_IDS)} sweep axis passes over {vox} (x{2 * len(LIGHT_IDS)} passes) "
            f"+ lit raymarch of {len(SAMPLE_ROWS)} evenly spaced rows of the {N_VOL}^3/{VIEW[1]}p frame "
            f"(x{VIEW[1] / len(SAMPLE_ROWS):g}), extrapolated linearly to the whole step")


REFERENCE_NOTE = {
    "reference": "the reference is a UE 5.4 / HLSL plugin with no CPU path; this runs ITS OWN shader sources (AddDirLightShader.usf, "
                 "WindowedRaymarchMaterials.usf) and host math (LightingShaderUtils.cpp) compiled for the CPU against an HLSL / engine-type shim "
                 "(oracle/ref.mk -> oracle/_ref/libtbrm_ref.so), OpenMP over all host threads",
    "port": "the reference (UE 5.4 / HLSL) has no CPU path; oracle/_ref is not present on this machine, so this is the CPU oracle port "
            "(bit-identical to the reference build) with all host threads",
}


def run_reference(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle

    threads = os.cpu_count() or 1
    kind = "reference" if reference_build_available() else "port"
    # every step is one bounded sample; the row sample shrinks (never below 20 rows) when many steps are asked for, so that the whole
    # run stays within a few minutes: 3 samples -> 10 % of the rows, 23 samples -> 20 rows
    global SAMPLE_ROWS
    n_rows = max(20, min(len(SAMPLE_ROWS), (3 * len(SAMPLE_ROWS)) // max(1, args.warmup + args.steps)))
    SAMPLE_ROWS = [int((i + 0.5) * VIEW[1] / n_rows) for i in range(n_rows)]
    data = oracle.synth_volume("perlin", (N_VOL,) * 3)  # untimed set-up; bit-identical to the device generator
    data_small = data if SWEEP_SAMPLE_N == N_VOL else oracle.synth_volume("perlin", (SWEEP_SAMPLE_N,) * 3)
    times, steps, detail = [], 0.0, {}
    for i in range(args.warmup + args.steps):
        est_s, est_steps, detail = oracle_sample(data, data_small, None, threads, kind)
        if i >= args.warmup:
            times.append(est_s)
            steps = est_steps
    ms = 1e3 * sum(times) / len(times)
    value = steps / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": "Mray-steps/s", "value": value, "unit": "Mray-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "Mray-steps/s", "cores": threads, "kind": kind, "sample": sample_text(), **detail},
        "estimated": True,
        "e2e": {"value": value, "unit": "Mray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": REFERENCE_NOTE[kind],
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args) -> int:
    import numpy as np
    import torch
    import torch.distributed as dist

    from tbraymarcherplugin_b200 import FMT_G8, _capi, sharding, synth
    from tbraymarcherplugin_b200.raymarch_utils import FSweepStats, FWindowingParameters, URaymarchUtils

    rank, world_size = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    lib = _capi.load()
    if lib.tbrm_device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    slabs = world_size > 1 and args.sharding == "slabs"  # ONE volume, Z-slab sharded (strong scaling); else one volume per rank
    t_start = time.perf_counter()

    def trace(what):
        if args.trace:
            print(f"[bench rank {rank} +{time.perf_counter() - t_start:7.2f}s] {what}", file=sys.stderr, flush=True)

    n, (W, H) = N_VOL, VIEW
    win = FWindowingParameters(0.45, 0.5, True, False)
    world = synth.identity_world()
    lights = [synth.LIGHTS[i] for i in LIGHT_IDS]
    cam = synth.benchmark_camera(W, H)

    # inputs: generated on the device (untimed), plus a pinned host copy for the end-to-end leg
    d_vol = torch.empty((n, n, n), dtype=torch.uint8, device="cuda")
    seed = synth.PERLIN_SEED + (0 if slabs else rank)  # independent volumes: every rank owns a different volume of the scene
    _capi.check(lib.tbrm_synth_volume_u8(local, _capi.SYNTH_PERLIN_CT, (C.c_int32 * 3)(n, n, n), seed & 0xFFFFFFFF, C.c_void_p(d_vol.data_ptr()), 1))
    if slabs:
        vol = sharding.FShardedRaymarchVolume((n, n, n), local)
        res = vol.res
        z0, z1 = vol.z0, vol.z1
        vol.SetDataVolumeSlab(d_vol[z0:z1])
        my_rows = vol.local_rows(H)
    else:
        vol = None
        res = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=True, device=local)
        URaymarchUtils.SetDataVolumeDevice(res, d_vol.data_ptr())
        z0, z1 = 0, n
        my_rows = H
    URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(res, win)
    d_img = torch.empty(H * W * 4, dtype=torch.float32, device="cuda")
    h_vol = torch.empty((z1 - z0, n, n), dtype=torch.uint8).pin_memory()  # this rank's share of the step's input
    h_vol.copy_(d_vol[z0:z1])
    h_img = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    if slabs:
        del d_vol  # the sharded volume holds the replicated copy
    torch.cuda.synchronize()

    def timer_begin():
        _capi.check(lib.tbrm_timer_begin(res.handle))

    def timer_end() -> float:
        ms = C.c_float()
        _capi.check(lib.tbrm_timer_end(res.handle, C.byref(ms)))
        return ms.value

    sweep_stats = []

    def sweep():
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)  # a sharded volume clears the slab it owns
        sweep_stats.clear()
        for i, l in enumerate(lights):
            st = FSweepStats()
            if slabs:
                # the last light of the reset pushes its finished bricks into every rank's light volume from inside the sweep kernel
                # (push-gather): gather_light() then only synchronises instead of all-gathering the slabs
                assert vol.AddDirLight(l, True, world, stats=st, push=(args.push_gather and i == len(lights) - 1))
            else:
                assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True, stats=st)
            sweep_stats.append(st)

    def gather_light():
        if slabs:
            vol.GatherLightVolume()  # NCCL all-gather of the light slabs, in place, ordered on the library's stream

    def raymarch(count=True, gather=True):
        """Returns (executed steps of this rank, frame tensor on rank 0 when sharded)."""
        if slabs:
            frame, steps = vol.Render(cam, world, STEPS, gather=gather, count_steps=count)
            return steps, frame
        _, steps = URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, STEPS, device_out_ptr=d_img.data_ptr(), count_steps=count)
        return steps, None

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(x, op):
        if world_size == 1:
            return float(x)
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t.item())

    MAX, SUM = (dist.ReduceOp.MAX, dist.ReduceOp.SUM) if world_size > 1 else (None, None)

    # ---- warm-up (also builds the lazily created replica / brick grid / scratch) ----
    trace("inputs ready")
    ray_steps = 0
    for i in range(max(args.warmup, 3)):
        sweep()
        if args.trace:
            URaymarchUtils.FlushRenderingCommands(res)
            trace(f"warm-up {i}: sweep done")
        gather_light()
        ray_steps, _ = raymarch()
        if slabs:
            vol.Check()  # fail fast if a slab exchange timed out
        trace(f"warm-up {i}: frame done")
    URaymarchUtils.FlushRenderingCommands(res)
    if slabs:
        vol.Check()
    trace("warm-up checked")
    # ---- parity of what the warm-up produced with the reference shaders' output (untimed; rank 0) ----
    parity = None
    if not args.no_parity and (slabs or world_size == 1):
        if slabs:
            _, frame_t = raymarch(count=False)
            vol.Flush()
            if rank == 0:
                parity = parity_against_reference(WORKLOAD, vol.light.cpu().numpy(), frame_t.cpu().numpy())
        else:
            fr, _ = URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, STEPS, count_steps=False)
            parity = parity_against_reference(WORKLOAD, URaymarchUtils.ReadLightVolume(res), fr)
        if rank == 0 and parity and (parity.get("light") is False or parity.get("frame") is False):
            print(f"bench.py: PARITY FAILURE against the reference shaders' digests: {parity}", file=sys.stderr, flush=True)
        trace(f"parity: {parity}")

    # ---- timed region: K steps, device time on the library's stream, max over ranks ----
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    launches0 = lib.tbrm_kernel_launch_count()
    timer_begin()
    for _ in range(args.steps):
        sweep()
        gather_light()
        raymarch(count=False)  # the step counter read-back would synchronise; steps are identical every frame
    total_ms = timer_end()
    launches = lib.tbrm_kernel_launch_count() - launches0
    barrier()
    clock_info = clocks.stop()
    trace("timed region done")
    total_ms = allreduce(total_ms, MAX)
    all_steps = allreduce(ray_steps, SUM)  # slabs: the ranks' shares of ONE frame; volumes: one frame per rank
    launches = int(allreduce(launches, SUM))
    ms_per_step = total_ms / args.steps
    value = all_steps / (ms_per_step * 1e-3) / 1e6

    # ---- per-stage timings (same stream, CUDA events; max over ranks), for the roofline objects ----
    reps = max(3, min(args.steps, 10))

    def stage(fn):
        barrier()
        timer_begin()
        for _ in range(reps):
            fn()
        return allreduce(timer_end() / reps, MAX)

    sweep_ms = stage(sweep)
    def gather_stage():  # what the step does after the sweep: a one-word all-reduce after a pushing sweep, else the all-gather of the slabs
        vol._pushed = bool(args.push_gather)
        vol.GatherLightVolume()

    gather_ms = stage(gather_stage) if slabs else 0.0
    ray_ms = stage(lambda: raymarch(count=False, gather=False))
    frame_gather_ms = (stage(lambda: raymarch(count=False, gather=True)) - ray_ms) if slabs else 0.0
    # one axis pass of the fused sweep on its own (L4 = a single Z pass), the sweep's dominant kernel. On a sharded volume the
    # slabs of a sweep along Z form a chain, so the roofline pass is timed with an X-major single pass... L4 is what N=1 times.
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    st4 = FSweepStats()
    URaymarchUtils.AddDirLightToSingleVolume(res, synth.LIGHTS[3], True, world, bGPUSync=True, stats=st4)
    pass_ms = stage(lambda: URaymarchUtils.AddDirLightToSingleVolume(res, synth.LIGHTS[3], True, world, bGPUSync=True))

    trace("stage timings done")
    hbm_peak, peak_src = peaks()
    vox = float(n) ** 3
    passes = sum(s.passes for s in sweep_stats)
    sweep_bytes = vox * (4.0 + passes * 9.0)  # clear (4 B) + per axis pass 1 B data + 4 B read + 4 B write
    ray_bytes = vox * 1.0 + vox * 4.0 + 256 * 16 * 8 + W * H * 16.0  # compulsory bytes of the raymarch
    roofline = {  # dominant kernel by time: the raymarch (instruction-issue bound, not HBM bound — reported honestly)
        "kernel": "raymarch_fast2_kernel (lit march, second generation without leaps, 128-thread blocks of 2 x 2 warps of 4 x 8 pixels)", "bound": "hbm", "achieved": ray_bytes / (ray_ms * 1e-3) / 1e9, "peak": hbm_peak * world_size, "unit": "GB/s",
        "frac": ray_bytes / (ray_ms * 1e-3) / 1e9 / (hbm_peak * world_size), "traffic": None, "peak_source": peak_src,
        "note": "compulsory bytes / kernel time; ~350 instr per non-empty sample make this kernel issue-bound (DESIGN.md §5.3, §6)",
    }
    # the march is bounded by instruction issue, not by HBM: the same launch against the FP32 pipes, with SURVEY's count of
    # 110 flop per executed step (an upper bound on useful work: skipped empty samples are counted) and peak = SMs x 128 lanes x 2 x SM clock
    sm_clock_ghz = (clock_info.get("sm_mhz") or 1965.0) / 1e3
    fp32_peak = 148 * 128 * 2 * sm_clock_ghz / 1e3  # TFLOP/s
    roofline["fp32"] = {"flop_per_step": 110, "achieved_TFLOPs": all_steps * 110.0 / (ray_ms * 1e-3) / 1e12, "peak_TFLOPs": fp32_peak * world_size,
                        "frac": all_steps * 110.0 / (ray_ms * 1e-3) / 1e12 / (fp32_peak * world_size),
                        "ncu": "profiles/r2_raymarch_fast2_kernel_ncu.txt: 129 thread-instructions per executed step, issue slots 64 % busy, ALU pipe 36 %, "
                               "FMA pipes 27 %, DRAM 0.7 %, top stall long scoreboard (L1 hits, 96 %)"}
    roofline_sweep = {
        "kernel": "sweep_tma_kernel (one axis pass along Z, 64 x 7 tiles where they fill the SMs; on a sharded volume its slabs run as a chain)", "bound": "hbm",
        "achieved": vox * 9.0 / (pass_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": vox * 9.0 / (pass_ms * 1e-3) / 1e9 / hbm_peak,
        "traffic": None, "peak_source": peak_src, "impl": list(st4.impl), "ms": pass_ms,
    }
    ncu = ROOT / "profiles" / "traffic.json"
    if ncu.exists() and world_size == 1 and N_VOL == 512:
        # DRAM bytes per launch measured by ncu on this workload (scripts/ncu_traffic.py writes the file together with the commit and the
        # library checksum it measured; a number taken under the profiler is never a time, only a byte count)
        t = json.loads(ncu.read_text())
        import hashlib

        meta = dict(t.get("_meta") or {})
        lib_path = ROOT / "tbraymarcherplugin_b200" / "libtbrm.so"
        meta["same_library"] = bool(meta.get("libtbrm_sha256_16")) and lib_path.exists() and \
            hashlib.sha256(lib_path.read_bytes()).hexdigest()[:16] == meta.get("libtbrm_sha256_16")
        roofline["traffic"] = t.get("raymarch_fast2_kernel", t.get("raymarch_fast_kernel"))
        roofline_sweep["traffic"] = t.get("sweep_tma_kernel")
        roofline["traffic_source"] = roofline_sweep["traffic_source"] = meta

    # ---- end to end: the reference-facing API with HOST buffers, copies inside the timed region ----
    def e2e_step():
        if slabs:
            vol.SetDataVolumeSlab(h_vol)  # H2D of this rank's slab (pinned) + all-gather over NVLink
        else:
            URaymarchUtils.SetDataVolume(res, h_vol.numpy())  # H2D of the step's input (pinned)
        sweep()
        gather_light()
        if slabs:
            _, frame = raymarch(count=False)
            if rank == 0:
                with torch.cuda.stream(vol.stream):
                    h_img.copy_(frame, non_blocking=True)  # D2H of the assembled frame (pinned)
            vol.Flush()
        else:
            URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, STEPS, out=h_img.numpy(), count_steps=False)  # D2H of the frame (pinned)

    # the box's own pinned-copy rates (one plain copy each way, untimed region): round 1 saw the serial figure differ 3x between two boxes
    # while the streamed figure agreed — the line now says what the PCIe path of THIS box delivers
    def copy_rate(dst, src):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        return src.numel() * src.element_size() / (time.perf_counter() - t0) / 1e9

    d_probe = torch.empty_like(h_vol, device="cuda")
    copy_rate(d_probe, h_vol)
    h2d_gbps = copy_rate(d_probe, h_vol)
    d_frame_probe = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    copy_rate(h_img, d_frame_probe)
    d2h_gbps = copy_rate(h_img, d_frame_probe)
    del d_probe, d_frame_probe

    e2e_step()  # also switches an unsharded resource set to an owned device copy before timing
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    serial_ms = allreduce(1e3 * (time.perf_counter() - t0) / args.steps, MAX)
    e2e_ms, pipelined, streamed_ms = serial_ms, False, None
    if slabs:
        # the sharded volume's streaming entry points: every rank uploads its slab of step i+1 and the ranks replicate it (all-gather on
        # the upload stream, its own communicator) while step i computes; rank 0 downloads frame i while step i+1 computes
        h_imgs = [h_img, torch.empty((H, W, 4), dtype=torch.float32).pin_memory()]

        def e2e_pipelined_slabs(k):
            vol.SetDataVolumeSlabAsync(h_vol)
            for i in range(k):
                vol.PresentDataVolume()
                if i + 1 < k:
                    vol.SetDataVolumeSlabAsync(h_vol)
                sweep()
                gather_light()
                vol.RenderToHostAsync(cam, world, STEPS, h_imgs[i % 2])
            vol.WaitForDownloads()
            vol.Flush()

        e2e_pipelined_slabs(2)
        barrier()
        t0 = time.perf_counter()
        e2e_pipelined_slabs(args.steps)
        torch.cuda.synchronize()
        streamed_ms = allreduce(1e3 * (time.perf_counter() - t0) / args.steps, MAX)
        vol.Check()
        # both modes are the public API; the line reports the faster one as the end-to-end figure and keeps the other next to it
        if streamed_ms < serial_ms:
            e2e_ms, pipelined = streamed_ms, True
    if not slabs:
        # The same K steps through the streaming entry points: step i+1's volume is copied H2D on the upload stream while step i
        # computes, frame i is copied D2H on the download stream while step i+1 computes. Every step still uploads its own input
        # and downloads its own frame inside the timed region (K uploads, K frames); what changes is that the copies overlap.
        h_imgs = [h_img, torch.empty((H, W, 4), dtype=torch.float32).pin_memory()]

        def e2e_pipelined(k):
            URaymarchUtils.SetDataVolumeAsync(res, h_vol.numpy())
            for i in range(k):
                URaymarchUtils.PresentDataVolume(res)
                if i + 1 < k:
                    URaymarchUtils.SetDataVolumeAsync(res, h_vol.numpy())
                sweep()
                URaymarchUtils.PerformWindowedLitRaymarchAsync(res, cam, world, STEPS, out=h_imgs[i % 2].numpy())
            URaymarchUtils.WaitForDownloads(res)
            URaymarchUtils.FlushRenderingCommands(res)

        e2e_pipelined(2)
        barrier()
        t0 = time.perf_counter()
        e2e_pipelined(args.steps)
        torch.cuda.synchronize()
        e2e_ms = allreduce(1e3 * (time.perf_counter() - t0) / args.steps, MAX)
        pipelined = True
    trace("end-to-end done")
    frames = 1 if slabs else world_size
    e2e = {"value": all_steps / (e2e_ms * 1e-3) / 1e6, "unit": "Mray-steps/s", "h2d_bytes_per_step": int(n) ** 3 * frames,
           "d2h_bytes_per_step": H * W * 16 * frames, "ms_per_step": e2e_ms, "serial_ms_per_step": serial_ms,
           "streamed_ms_per_step": streamed_ms if slabs else e2e_ms,
           "pinned_copy_GBps": {"h2d": h2d_gbps, "d2h": d2h_gbps},
           "mode": (("streaming API: the H2D copy of step i+1 and the D2H copy of frame i overlap the compute of their neighbours "
                     "(tbrm_upload_volume_async / tbrm_present_volume / tbrm_raymarch_lit_to_host_async)") if not slabs else
                    ("streaming API of the sharded volume: every rank's slab upload + the data all-gather of step i+1 (upload stream, own "
                     "communicator) and rank 0's frame download of step i overlap the compute of their neighbours "
                     "(SetDataVolumeSlabAsync / PresentDataVolume / RenderToHostAsync)")) if pipelined else
                   "serial: H2D copy, sweep, raymarch, D2H copy one after the other"}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on a bounded sample ----
    cpu_baseline = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        sweep()
        light = URaymarchUtils.ReadLightVolume(res)
        threads = os.cpu_count() or 1
        sys.path.insert(0, str(ROOT / "tests"))
        import oracle

        kind = "reference" if reference_build_available() else "port"
        small = h_vol.numpy() if SWEEP_SAMPLE_N == n else oracle.synth_volume("perlin", (SWEEP_SAMPLE_N,) * 3)
        est_s, est_steps, detail = oracle_sample(h_vol.numpy(), small, light, threads, kind)
        cpu_baseline = {"value": est_steps / est_s / 1e6, "unit": "Mray-steps/s", "cores": threads, "kind": kind, "sample": sample_text(),
                        "est_ms_per_step": est_s * 1e3, "note": REFERENCE_NOTE[kind], **detail}

    # ---- the reference's other light-volume formats (N = 1): G8 is its DEFAULT (RaymarchVolume.h:198-199), half resolution its "massive
    # speedup" option (Readme.md:214). Same reset + frame as the step, timed once each; these run the generic fused sweep today.
    formats = None
    if world_size == 1 and WORKLOAD == "cfg2" and not args.no_formats:
        formats = {}
        for name, l32, half in (("g8", False, False), ("r32f_half", True, True), ("g8_half", False, True)):
            rf = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=l32, LightVolumeHalfResolution=half, device=local)
            URaymarchUtils.SetDataVolumeDevice(rf, d_vol.data_ptr())
            URaymarchUtils.ColorCurveToTexture(rf, synth.soft_ct_curve())
            URaymarchUtils.SetWindowingParameters(rf, win)
            stf = []

            def sweep_f():
                URaymarchUtils.ClearResourceLightVolumes(rf, 0.0)
                stf.clear()
                for l in lights:
                    st = FSweepStats()
                    assert URaymarchUtils.AddDirLightToSingleVolume(rf, l, True, world, bGPUSync=True, stats=st)
                    stf.append(st)

            def frame_f():
                URaymarchUtils.PerformWindowedLitRaymarch(rf, cam, world, STEPS, device_out_ptr=d_img.data_ptr(), count_steps=False)

            def timed(fn, reps=3):
                fn()
                ms = C.c_float()
                _capi.check(lib.tbrm_timer_begin(rf.handle))
                for _ in range(reps):
                    fn()
                _capi.check(lib.tbrm_timer_end(rf.handle, C.byref(ms)))
                return ms.value / reps

            formats[name] = {"sweep_ms": timed(sweep_f), "raymarch_ms": timed(frame_f), "sweep_impl": [list(s.impl) for s in stf],
                             "light_dims": list(rf.LightDims)}
            rf.release()
        trace(f"formats: {formats}")

    # ---- BASELINE.json configs[3] (north_star's scaling figure): the illumination sweep of a 1024^3 volume, 3 lights, full reset, on the
    # same N GPUs — sweep only, Mvoxels/s = light voxels x axis passes / device time (max over ranks), light volume checked bit-for-bit
    # against the reference shaders' digest. Rides along in every line so that the driver's 1/2/4/8 runs carry it.
    scale_cfg4 = None
    if not args.no_cfg4 and WORKLOAD == "cfg2" and (slabs or world_size == 1):
        trace("scale_cfg4: set-up")
        n4, lights4 = 1024, [synth.LIGHTS[i] for i in (0, 1, 2)]
        if slabs:
            vol.release()
            vol = None
            vol4 = sharding.FShardedRaymarchVolume((n4, n4, n4), local)
            res4 = vol4.res
            za, zb = vol4.z0, vol4.z1
            _capi.check(lib.tbrm_synth_volume_u8(local, _capi.SYNTH_PERLIN_CT, (C.c_int32 * 3)(n4, n4, n4), synth.PERLIN_SEED & 0xFFFFFFFF,
                                                 C.c_void_p(vol4.data.data_ptr()), 1))
            torch.cuda.synchronize()
            _capi.check(lib.tbrm_bind_volume_device(res4.handle, C.c_void_p(vol4.data.data_ptr())))
        else:
            res.release()
            d4 = torch.empty((n4, n4, n4), dtype=torch.uint8, device="cuda")
            _capi.check(lib.tbrm_synth_volume_u8(local, _capi.SYNTH_PERLIN_CT, (C.c_int32 * 3)(n4, n4, n4), synth.PERLIN_SEED & 0xFFFFFFFF, C.c_void_p(d4.data_ptr()), 1))
            res4 = URaymarchUtils.InitializeRaymarchResources((n4, n4, n4), FMT_G8, bLightVolume32Bit=True, device=local)
            URaymarchUtils.SetDataVolumeDevice(res4, d4.data_ptr())
            vol4 = None
        URaymarchUtils.ColorCurveToTexture(res4, synth.soft_ct_curve())
        URaymarchUtils.SetWindowingParameters(res4, win)
        stats4 = []

        def sweep4():
            URaymarchUtils.ClearResourceLightVolumes(res4, 0.0)
            stats4.clear()
            for l in lights4:
                st = FSweepStats()
                assert URaymarchUtils.AddDirLightToSingleVolume(res4, l, True, world, bGPUSync=True, stats=st)
                stats4.append(st)

        sweep4()
        URaymarchUtils.FlushRenderingCommands(res4)
        if vol4 is not None:
            vol4.Check()
        barrier()
        _capi.check(lib.tbrm_timer_begin(res4.handle))
        reps4 = 3
        for _ in range(reps4):
            sweep4()
        ms4 = C.c_float()
        _capi.check(lib.tbrm_timer_end(res4.handle, C.byref(ms4)))
        sweep4_ms = allreduce(ms4.value / reps4, MAX)
        passes4 = sum(s.passes for s in stats4)
        par4 = None
        if not args.no_parity:
            if vol4 is not None:
                vol4.GatherLightVolume()
                vol4.Flush()
                vol4.Check()
                if rank == 0:
                    par4 = parity_against_reference("cfg4", light=vol4.light.cpu().numpy())
            else:
                par4 = parity_against_reference("cfg4", light=URaymarchUtils.ReadLightVolume(res4))
        scale_cfg4 = {"workload": "cfg4 sweep: 1024^3 Perlin R8, R32F light volume, 3 dir lights full reset (clear + 6 axis passes), sweep only",
                      "n_gpus": world_size, "sweep_ms": sweep4_ms, "axis_passes": passes4,
                      "Mvoxels_per_s": float(n4) ** 3 * passes4 / (sweep4_ms * 1e-3) / 1e6,
                      "GB_per_s": float(n4) ** 3 * (4.0 + 9.0 * passes4) / (sweep4_ms * 1e-3) / 1e9,
                      "impl": [list(s.impl) for s in stats4], "parity": par4}
        trace(f"scale_cfg4: {scale_cfg4}")
        if vol4 is not None:
            vol4.release()
        else:
            res4.release()
    if slabs and vol is not None:
        vol.Check()
    if rank == 0:
        frame_steps = all_steps if slabs else ray_steps
        line = {
            "metric": "Mray-steps/s", "value": value, "unit": "Mray-steps/s", "n_gpus": world_size, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if (world_size > 1 and not slabs) else "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world_size, slabs), "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "roofline_sweep": roofline_sweep, "cpu_baseline": cpu_baseline, "parity": parity, "scale_cfg4": scale_cfg4, "formats": formats,
            "stages": {
                "ray_steps_per_frame": frame_steps,
                "raymarch": {"ms": ray_ms, "Mray_steps_per_s": all_steps / (ray_ms * 1e-3) / 1e6},
                "sweep": {"ms": sweep_ms, "axis_passes": passes, "Mvoxels_per_s": vox * passes * frames / (sweep_ms * 1e-3) / 1e6,
                          "GB_per_s": sweep_bytes * frames / (sweep_ms * 1e-3) / 1e9,
                          "frac_of_hbm_peak": sweep_bytes * frames / (sweep_ms * 1e-3) / 1e9 / (hbm_peak * world_size),
                          "impl": [list(s.impl) for s in sweep_stats]},
                "light_all_gather_ms": gather_ms, "frame_gather_ms": frame_gather_ms,
            },
        }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()
        if slabs and vol is not None:
            vol.release()
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the (untimed) digest comparison with the reference shaders' output")
    ap.add_argument("--push-gather", default="auto", choices=["auto", "on", "off"],
                    help="N > 1: push the finished light bricks into every rank's volume from the sweep's last pass (TMA stores over NVLink) instead "
                         "of an NCCL all-gather of the slabs. auto = on for 2 GPUs only: measured on B200s, the push wins at N = 2 (step 7.71 -> "
                         "7.39 ms) and loses at N = 8 (5.31 -> 8.27 ms: seven unicast copies per brick against NCCL's switch multicast)")
    ap.add_argument("--no-formats", action="store_true", help="skip the extra timings of the G8 / half-resolution light-volume formats")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the extra 1024^3 sweep measurement (scale_cfg4)")
    ap.add_argument("--trace", action="store_true", help="print progress lines to stderr (debugging multi-GPU runs)")
    ap.add_argument("--sharding", default="slabs", choices=["slabs", "volumes"],
                    help="N > 1: 'slabs' = ONE volume Z-slab sharded over the GPUs (strong scaling, default); 'volumes' = one volume per GPU (weak)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"],
                    help="cfg2 = 512^3 / 1080p / 512 steps / 2 lights (BASELINE.json configs[1], default); cfg4 = 1024^3 / 2160p / 768 steps / 3 lights")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    if args.workload == "cfg4":
        global N_VOL, VIEW, STEPS, LIGHT_IDS, SAMPLE_ROWS
        N_VOL, VIEW, STEPS, LIGHT_IDS = 1024, (3840, 2160), 768.0, (0, 1, 2)
        SAMPLE_ROWS = list(range(54, 2160, 108))  # 20 rows: a 1024^3 CPU sample is bounded harder (the sweep pass runs at 512^3, x8)
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 2
        args.warmup = args.warmup if args.warmup is not None else 1
        return run_reference(args)
    args.steps = args.steps if args.steps is not None else 20
    args.warmup = args.warmup if args.warmup is not None else 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.push_gather = args.push_gather == "on" or (args.push_gather == "auto" and world == 2)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
