"""Full-size GPU checks at BASELINE.json's sizes, through size-independent properties (the oracle is too slow there):
the three sweep schedules agree bit-for-bit, the closed form of a homogeneous volume, Add-then-Remove, the fast raymarch
kernel vs the generic one, row-sharded rendering, and a sampled comparison against the oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle
from tbraymarcherplugin_b200 import FMT_G8, _capi, synth
from tbraymarcherplugin_b200.raymarch_utils import FDirLightParameters, FSweepStats, FWindowingParameters, URaymarchUtils

pytestmark = pytest.mark.gpu
CT_WINDOW = FWindowingParameters(0.45, 0.5, True, False)


def device_volume(n, kind=_capi.SYNTH_PERLIN_CT):
    import torch

    d = torch.empty(n * n * n, dtype=torch.uint8, device="cuda")
    _capi.check(_capi.load().tbrm_synth_volume_u8(0, kind, (C.c_int32 * 3)(n, n, n), synth.PERLIN_SEED, C.c_void_p(d.data_ptr()), 1))
    return d


def make_res(n, d, impl, curve=synth.soft_ct_curve(), win=CT_WINDOW, debug1=0):
    res = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=True)
    URaymarchUtils.SetDataVolumeDevice(res, d.data_ptr())
    URaymarchUtils.ColorCurveToTexture(res, curve)
    URaymarchUtils.SetWindowingParameters(res, win)
    URaymarchUtils.SetOptions(res, sweep_impl=impl)
    return res


def test_sweep_schedules_agree_bitwise_at_256():
    n = 256
    d = device_volume(n)
    out = {}
    for impl in (1, 3, 2):  # per-slice launches, generic fused, TMA-staged fused
        res = make_res(n, d, impl)
        for l in synth.LIGHTS[:3]:
            st = FSweepStats()
            URaymarchUtils.AddDirLightToSingleVolume(res, l, True, synth.identity_world(), bGPUSync=(impl != 1), stats=st)
            # option values: 1 per-slice, 2 fused (TMA-staged when eligible), 3 generic fused; stats: 1 per-slice, 2 generic fused, 3 TMA
            assert set(st.impl) == {{1: 1, 3: 2, 2: 3}[impl]}
        out[impl] = URaymarchUtils.ReadLightVolume(res)
        res.release()
    assert np.array_equal(out[1], out[3]) and np.array_equal(out[1], out[2])
    assert out[1].max() > 1.5  # three lights add up


def test_homogeneous_volume_closed_form_at_512_slices():
    import torch

    n, I = 512, 0.9
    d = torch.full((n * n * n,), 128, dtype=torch.uint8, device="cuda")
    res = make_res(n, d, 2, win=FWindowingParameters())
    URaymarchUtils.SetOptions(res, border_exact=True, sweep_impl=2)
    st = FSweepStats()
    URaymarchUtils.AddDirLightToSingleVolume(res, FDirLightParameters((0, 0, -1), I), True, synth.identity_world(), bGPUSync=True, stats=st)
    assert st.passes == 1 and st.impl == (3,)
    L = URaymarchUtils.ReadLightVolume(res)
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    a_tf = float(oracle.sample_windowed_tf(128 / 255.0, 1.0, tf, FWindowingParameters())[3])
    alpha = 1.0 - (1.0 - a_tf) ** (100.0 / n)
    k = np.arange(n)
    exp = I * (1.0 - alpha) ** k
    exp = np.where(exp > 1e-3, exp, 0.0)[::-1]  # slice n-1 is the first one
    got = L[:, n // 2, n // 2]
    assert np.allclose(got, exp, atol=5e-5)
    assert np.array_equal(L[:, 0, 0], got) and np.array_equal(L[:, n - 1, 17], got)  # every column of a homogeneous volume is alike


def test_add_then_remove_restores_light_volume_at_512():
    n = 512
    d = device_volume(n)
    res = make_res(n, d, 2)
    world = synth.identity_world()
    URaymarchUtils.AddDirLightToSingleVolume(res, synth.LIGHTS[1], True, world, bGPUSync=True)
    base = URaymarchUtils.ReadLightVolume(res)
    URaymarchUtils.AddDirLightToSingleVolume(res, synth.LIGHTS[0], True, world, bGPUSync=True)
    URaymarchUtils.AddDirLightToSingleVolume(res, synth.LIGHTS[0], False, world, bGPUSync=True)
    back = URaymarchUtils.ReadLightVolume(res)
    assert np.abs(back - base).max() < 1e-6
    # ClearResourceLightVolumes + the same light again reproduces the first state bit-for-bit (idempotence of a reset)
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    URaymarchUtils.AddDirLightToSingleVolume(res, synth.LIGHTS[1], True, world, bGPUSync=True)
    assert np.array_equal(URaymarchUtils.ReadLightVolume(res), base)


def test_fast_raymarch_equals_generic_kernel_and_oracle_rows_at_cfg2():
    n = 512
    d = device_volume(n)
    res = make_res(n, d, 2)
    world = synth.identity_world()
    for l in synth.LIGHTS[:2]:
        URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True)
    cam = synth.benchmark_camera(1920, 1080)
    fast, steps = URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, 512.0)
    URaymarchUtils.SetOptions(res, sweep_impl=2, debug_flags=(0, 1))  # reserved[1] = 1: force the generic raymarch kernel
    slow, steps2 = URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, 512.0)
    assert steps == steps2 and steps > 5e8
    assert np.array_equal(fast, slow)
    assert 0.2 < (fast[..., 3] > 0).mean() < 0.9
    # rows rendered separately (the unit of image-tile sharding) are the rows of the full frame
    URaymarchUtils.SetOptions(res, sweep_impl=2)
    part, _ = URaymarchUtils.PerformWindowedLitRaymarch(res, cam, world, 512.0, rows=(500, 540))
    assert np.array_equal(part, fast[500:540])
    # three rows against the CPU oracle on the downloaded 512^3 volumes
    import torch

    data = d.cpu().numpy().reshape(n, n, n)
    ora = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), CT_WINDOW)
    ora.light = URaymarchUtils.ReadLightVolume(res)
    for r in (270, 540, 811):
        ref, _ = ora.raymarch_lit(cam, world, 512.0, rows=(r, r + 1))
        assert np.array_equal(ref[0], fast[r]), f"row {r}: max diff {np.abs(ref[0] - fast[r]).max()}"
