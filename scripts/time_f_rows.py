"""Times the kernels of the rows next to the hot path on one GPU and writes gpurun_out/f_rows_timing.json.

    python scripts/time_f_rows.py [N]          (N = volume side, default 512)

Torch-free (starts in a second): device buffers and events come from libcudart through ctypes; ops on a resource set are timed with
tbrm_timer_begin / tbrm_timer_end (CUDA events on the resource set's stream), the ingest / Mandelbulb ops — which run on the per-thread
stream — with events recorded on that stream around the call (device-to-device calls allocate nothing). After a warm-up call, the
minimum and the median of the repetitions are reported together with the algorithmic bytes of DESIGN.md §5.5."""
import ctypes as C
import json
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from tbraymarcherplugin_b200 import FMT_G8, _capi, synth  # noqa: E402
from tbraymarcherplugin_b200.raymarch_utils import FMandelbulbParameters, FWindowingParameters, URaymarchUtils  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = _capi.load()
rt = None
for name in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        rt = C.CDLL(name)
        break
    except OSError:
        pass
assert rt is not None, "libcudart not found"
PER_THREAD = C.c_void_p(2)  # cudaStreamPerThread
results = {"volume": [N, N, N], "peak_hbm_gbs": None}
try:
    results["peak_hbm_gbs"] = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    pass


def cuda(err):
    assert err == 0, f"CUDA error {err}"


def dmalloc(nbytes):
    p = C.c_void_p()
    cuda(rt.cudaMalloc(C.byref(p), C.c_size_t(nbytes)))
    return p


def timed_per_thread(fn, reps):
    e0, e1 = C.c_void_p(), C.c_void_p()
    cuda(rt.cudaEventCreate(C.byref(e0))), cuda(rt.cudaEventCreate(C.byref(e1)))
    fn()  # warm-up
    ms = []
    for _ in range(reps):
        cuda(rt.cudaEventRecord(e0, PER_THREAD))
        fn()
        cuda(rt.cudaEventRecord(e1, PER_THREAD))
        cuda(rt.cudaEventSynchronize(e1))
        t = C.c_float()
        cuda(rt.cudaEventElapsedTime(C.byref(t), e0, e1))
        ms.append(t.value)
    return ms


def timed_resource(res, fn, reps):
    fn()
    ms = []
    for _ in range(reps):
        t = C.c_float()
        _capi.check(lib.tbrm_timer_begin(res.handle))
        fn()
        _capi.check(lib.tbrm_timer_end(res.handle, C.byref(t)))
        ms.append(t.value)
    return ms


def report(name, ms, bytes_=None, **extra):
    r = {"ms_min": min(ms), "ms_median": statistics.median(ms), "reps": len(ms), **extra}
    if bytes_:
        r["algorithmic_bytes"] = bytes_
        r["GBps_at_min"] = bytes_ / min(ms) / 1e6
        if results["peak_hbm_gbs"]:
            r["frac_of_hbm_peak"] = r["GBps_at_min"] / results["peak_hbm_gbs"]
    results[name] = r
    print(name, json.dumps(r), flush=True)


def section(fn):
    try:
        fn()
    except Exception as e:  # keep going: every section is independent
        results.setdefault("errors", []).append(f"{fn.__name__}: {e!r}")
        print("ERROR", fn.__name__, repr(e), flush=True)


t_start = time.time()
vox = N * N * N
d_vol = dmalloc(vox)
_capi.check(lib.tbrm_synth_volume_u8(0, _capi.SYNTH_PERLIN_CT if hasattr(_capi, "SYNTH_PERLIN_CT") else 1, (C.c_int32 * 3)(N, N, N),
                                     synth.PERLIN_SEED & 0xFFFFFFFF, d_vol, 1))
res = URaymarchUtils.InitializeRaymarchResources((N, N, N), FMT_G8, bLightVolume32Bit=True)
URaymarchUtils.SetDataVolumeDevice(res, d_vol.value)
URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
URaymarchUtils.SetWindowingParameters(res, FWindowingParameters(0.45, 0.5, True, False))
world = synth.identity_world().to_c()
W, H = (1920, 1080) if N >= 512 else (512, 512)
steps = 512.0 if N >= 512 else 256.0
cam = synth.benchmark_camera(W, H).to_c()
d_frame = dmalloc(W * H * 16)


def octree():
    ms = timed_resource(res, lambda: _capi.check(lib.tbrm_generate_octree(res.handle)), 20)
    ovox = 1
    for m in range(1):
        d = (C.c_int32 * 3)()
        lib.tbrm_octree_mip_dims(res.handle, 0, d)
        ovox = d[0] * d[1] * d[2]
    report("octree_build", ms, vox * 1 + int(2 * ovox * (1 + 1 / 8 + 1 / 64 + 1 / 512)), octree_voxels=ovox)


def marches():
    n = C.c_uint64(0)
    _capi.check(lib.tbrm_raymarch_intensity(res.handle, C.byref(cam), C.byref(world), steps, 0, H, d_frame, 1, C.byref(n)))
    ms = timed_resource(res, lambda: _capi.check(lib.tbrm_raymarch_intensity(res.handle, C.byref(cam), C.byref(world), steps, 0, H, d_frame, 1, None)), 10)
    report("raymarch_intensity", ms, None, view=[W, H], ray_steps=int(n.value), Mray_steps_per_s=n.value / min(ms) / 1e3)
    for mip in (0, 2):
        _capi.check(lib.tbrm_raymarch_octree(res.handle, C.byref(cam), C.byref(world), steps, mip, 0, H, d_frame, 1, C.byref(n)))
        ms = timed_resource(res, lambda: _capi.check(lib.tbrm_raymarch_octree(res.handle, C.byref(cam), C.byref(world), steps, mip, 0, H, d_frame, 1, None)), 10)
        report(f"raymarch_octree_mip{mip}", ms, None, view=[W, H], step_count=steps, ray_steps=int(n.value), Mray_steps_per_s=n.value / min(ms) / 1e3)


def ingest():
    d_in = dmalloc(vox * 4)
    d_out = dmalloc(vox * 4)
    for k in range(4):  # varied bytes everywhere
        cuda(rt.cudaMemcpy(C.c_void_p(d_in.value + k * vox), d_vol, C.c_size_t(vox), 3))
    lo, hi = C.c_float(), C.c_float()
    for fmt, name, ib, ob in ((0, "u8", 1, 1), (3, "i16", 2, 2), (5, "i32", 4, 2)):
        ms = timed_per_thread(lambda: _capi.check(lib.tbrm_normalize_volume(0, fmt, d_in, 1, vox, d_out, 1, C.byref(lo), C.byref(hi))), 10)
        report(f"normalize_{name}", ms, vox * (2 * ib + ob), min=lo.value, max=hi.value)
    ms = timed_per_thread(lambda: _capi.check(lib.tbrm_convert_volume_to_float(0, 3, d_in, 1, vox, d_out, 1)), 10)
    report("to_float_i16", ms, vox * (2 + 4))
    rt.cudaFree(d_in), rt.cudaFree(d_out)


def mandelbulb():
    M = 256
    d_sdf = dmalloc(M * M * M * 2)
    it = C.c_uint64(0)
    dims, c = (C.c_int32 * 3)(M, M, M), (C.c_float * 3)(0, 0, 0)
    _capi.check(lib.tbrm_mandelbulb_sdf(0, dims, c, 2.0, 8.0, 1, d_sdf, 1, C.byref(it)))
    ms = timed_per_thread(lambda: _capi.check(lib.tbrm_mandelbulb_sdf(0, dims, c, 2.0, 8.0, 1, d_sdf, 1, None)), 5)
    report("mandelbulb_sdf_bake_256_g16", ms, None, sdf_iterations=int(it.value), Giter_per_s=it.value / min(ms) / 1e6)
    mb = FMandelbulbParameters().to_c()
    camm = synth.benchmark_camera(1920, 1080, jitter=False).to_c()
    d_n = dmalloc(1920 * 1080 * 16)
    _capi.check(lib.tbrm_mandelbulb_march_normal(0, C.byref(mb), 0.01, C.byref(camm), C.byref(world), 0, 1080, d_n, 1, C.byref(it)))
    ms = timed_per_thread(lambda: _capi.check(lib.tbrm_mandelbulb_march_normal(0, C.byref(mb), 0.01, C.byref(camm), C.byref(world), 0, 1080, d_n, 1, None)), 3)
    report("mandelbulb_normal_1080p", ms, None, sdf_iterations=int(it.value), Giter_per_s=it.value / min(ms) / 1e6)
    _capi.check(lib.tbrm_mandelbulb_march(0, C.byref(mb), C.byref(camm), C.byref(world), 0, 1080, d_n, 1, C.byref(it)))
    ms = timed_per_thread(lambda: _capi.check(lib.tbrm_mandelbulb_march(0, C.byref(mb), C.byref(camm), C.byref(world), 0, 1080, d_n, 1, None)), 3)
    report("mandelbulb_distance_1080p", ms, None, sdf_iterations=int(it.value), Giter_per_s=it.value / min(ms) / 1e6)


def joined_lights():
    """Same-face passes of 4 lights joined into one per-slice sweep vs the same lights added one after the other (per-slice and fused)."""
    from tbraymarcherplugin_b200.raymarch_utils import FSweepStats

    w = synth.identity_world()
    lights = [synth.rotate_about_z(synth.LIGHTS[0], 3.0 * i) for i in range(4)]  # the same two faces for all four
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    st = FSweepStats()
    URaymarchUtils.AddDirLightsToSingleVolumeJoined(res, lights, True, w, stats=st)
    ms = timed_resource(res, lambda: URaymarchUtils.AddDirLightsToSingleVolumeJoined(res, lights, True, w), 3)
    report("add_4_lights_joined_per_slice", ms, None, sweeps=st.passes, launches=st.kernel_launches)
    for sync, name in ((False, "per_slice"), (True, "fused")):
        def seq():
            for l in lights:
                URaymarchUtils.AddDirLightToSingleVolume(res, l, True, w, bGPUSync=sync)
        ms = timed_resource(res, seq, 3)
        report(f"add_4_lights_one_by_one_{name}", ms, None)


for s in (octree, marches, ingest, mandelbulb, joined_lights):
    section(s)
results["wall_s"] = time.time() - t_start
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
(out / "f_rows_timing.json").write_text(json.dumps(results, indent=1))
print("done in %.1f s" % results["wall_s"])
