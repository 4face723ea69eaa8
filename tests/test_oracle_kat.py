"""Known-answer tests that pin the CPU oracle (the reference ships no golden vectors: parity is unpinned).

Closed forms follow SURVEY.md §4: homogeneous volume + axis-aligned light -> L_s = I(1-a)^s; empty volume -> light = sum I*w;
Add then Remove ~ 0; Change == Remove + Add when the major axes match; single ray through a homogeneous slab."""
import math

import numpy as np
import pytest

import oracle
from tbraymarcherplugin_b200 import synth
from tbraymarcherplugin_b200.raymarch_utils import (FCamera, FClippingPlaneParameters, FDirLightParameters, FRaymarchWorldParameters,
                                                    FTransform, FWindowingParameters)


def test_det_pow_matches_libm():
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.random(4000), 1 - 10 ** rng.uniform(-7, 0, 2000), np.arange(256) / 255.0]).astype(np.float32)
    worst = 0.0
    for y in (0.05, 0.1953125, 0.39, 1.0, 7.5, 100.0):
        for x in xs[(xs > 0) & (xs <= 1)][::7]:
            worst = max(worst, abs(oracle.det_pow(float(x), y) - math.pow(float(x), float(np.float32(y)))))
    assert worst < 1.5e-7
    assert oracle.det_pow(1.0, 0.3) == 1.0 and oracle.det_pow(0.0, 0.3) == 0.0


def test_round_to_half_is_rne_fp16():
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.standard_normal(2000) * 10 ** rng.uniform(-8, 4, 2000), [0.0, 1.0, 65504.0, 6e-8, 0.333, 1 / 255]]).astype(np.float32)
    for v in vals:
        assert oracle.lib().tbo_round_to_half(float(v)) == float(np.float32(np.float16(v))), v


def test_default_tf_is_white_ramp_full_opacity():
    tf = oracle.default_tf()
    ramp = (np.arange(256, dtype=np.float32) / np.float32(255)).astype(np.float16).astype(np.float32)
    assert np.array_equal(tf[:, 0], ramp) and np.array_equal(tf[:, 3], np.ones(256, np.float32))


def test_trilinear_sampling_of_a_linear_ramp():
    X, Y, Z = 8, 6, 5
    z, y, x = np.meshgrid(np.arange(Z), np.arange(Y), np.arange(X), indexing="ij")
    vol = (0.1 * x + 0.01 * y + 0.5 * z).astype(np.float32)
    # inside the volume the sample of a linear field is the field at the texel-space position u*N-0.5
    for (u, v, w) in [(0.5, 0.5, 0.5), (0.3, 0.71, 0.42), (0.9, 0.2, 0.65)]:
        exp = 0.1 * (u * X - 0.5) + 0.01 * (v * Y - 0.5) + 0.5 * (w * Z - 0.5)
        assert abs(oracle.sample_data(vol, u, v, w) - exp) < 1e-5
    # address modes at the edge
    assert oracle.sample_data(vol, 0.0, 0.5, 0.5, oracle.ADDR_CLAMP) == pytest.approx(oracle.sample_data(vol, 0.5 / X, 0.5, 0.5), abs=1e-6)
    b = oracle.sample_data(vol, 0.0, 0.5 / Y, 0.5 / Z, oracle.ADDR_BORDER, 7.0)  # half border, half texel 0
    assert b == pytest.approx(0.5 * 7.0 + 0.5 * vol[0, 0, 0], abs=1e-6)
    wv = oracle.sample_data(vol, 0.0, 0.5 / Y, 0.5 / Z, oracle.ADDR_WRAP)
    assert wv == pytest.approx(0.5 * vol[0, 0, X - 1] + 0.5 * vol[0, 0, 0], abs=1e-6)


def test_windowed_tf_cutoffs_and_step_correction():
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    w = FWindowingParameters(0.45, 0.5, True, False)
    assert np.all(oracle.sample_windowed_tf(0.1, 0.2, tf, w) == 0)  # below the window, low cut-off on
    hi = oracle.sample_windowed_tf(0.95, 0.2, tf, w)  # above the window, high cut-off off -> clamps to the last texel
    assert hi[0] == pytest.approx(1.0, abs=1e-3) and hi[3] == pytest.approx(1 - (1 - float(tf[255, 3])) ** 0.2, abs=2e-7)
    w2 = FWindowingParameters(0.45, 0.5, True, True)
    assert np.all(oracle.sample_windowed_tf(0.95, 0.2, tf, w2) == 0)
    mid = oracle.sample_windowed_tf(0.45, 1.0, tf, w)  # TFPos 0.5 -> between texels 127 and 128, step 1 -> alpha unchanged
    assert mid[3] == pytest.approx(0.5 * (tf[127, 3] + tf[128, 3]), abs=1e-6)


def _homogeneous(n, value):
    return np.full((n, n, n), value, np.uint8)


@pytest.mark.parametrize("direction,face", [((0, 0, -1), 4), ((0, 0, 1), 5), ((-1, 0, 0), 0), ((0, 1, 0), 3)])
def test_homogeneous_volume_axis_aligned_light_is_geometric(direction, face):
    n, I = 24, 0.9
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    w = FWindowingParameters()
    vol = oracle.OracleVolume(_homogeneous(n, 128), tf, w, border_exact=True)
    light = FDirLightParameters(direction, I)
    world = synth.identity_world()
    plan = oracle.plan_dir_light(vol.ldims, w, light, world, True)
    assert plan.add_passes == 1 and plan.passes[0].face == face and plan.passes[0].weight == 1.0
    assert vol.add_dir_light(light, True, world) == 1
    a_tf = float(oracle.sample_windowed_tf(128 / 255.0, 1.0, tf, w)[3])
    alpha = 1.0 - (1.0 - a_tf) ** (100.0 / n)  # StepSize = 1/n, VOLUME_DENSITY = 100
    axis, dirn = face // 2, (1 if face % 2 else -1)
    L = np.moveaxis(vol.light, 2 - axis, 0)  # slices along the sweep axis first
    for k in range(n):
        j = (n - 1 - k) if dirn < 0 else k
        # the first slice samples outside the volume (gated off), slice k has crossed k samples
        exp = I * (1.0 - alpha) ** k
        exp = exp if exp > 1e-3 else 0.0
        assert np.allclose(L[j], exp, atol=2e-6), (k, float(L[j].mean()), exp)


def test_empty_volume_light_is_sum_of_axis_weights():
    n = 20
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    w = FWindowingParameters(0.45, 0.5, True, False)  # zeros are below the window -> cut off -> no extinction
    vol = oracle.OracleVolume(np.zeros((n, n, n), np.uint8), tf, w, border_exact=True)
    for l in synth.LIGHTS:
        vol.add_dir_light(l, True, synth.identity_world())
    assert np.allclose(vol.light, sum(l.LightIntensity for l in synth.LIGHTS), atol=1e-5)


def test_add_then_remove_restores_the_light_volume():
    n = 24
    vol = oracle.OracleVolume(synth.perlin_ct_volume((n, n, n)), oracle.prepare_tf(synth.soft_ct_curve()), FWindowingParameters(0.45, 0.5, True, False))
    world = synth.scaled_rotated_world()
    vol.add_dir_light(synth.LIGHTS[1], True, world)
    base = vol.light.copy()
    vol.add_dir_light(synth.LIGHTS[0], True, world)
    assert np.abs(vol.light - base).max() > 0.1
    vol.add_dir_light(synth.LIGHTS[0], False, world)
    assert np.abs(vol.light - base).max() < 1e-6


def test_change_equals_remove_plus_add_when_axes_match():
    n = 24
    data = synth.perlin_ct_volume((n, n, n))
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    w = FWindowingParameters(0.45, 0.5, True, False)
    world = synth.identity_world()
    old, new = synth.LIGHTS[0], synth.rotate_about_z(synth.LIGHTS[0], 5.0)
    a = oracle.OracleVolume(data, tf, w)
    b = oracle.OracleVolume(data, tf, w)
    a.add_dir_light(old, True, world)
    b.add_dir_light(old, True, world)
    assert a.change_dir_light(old, new, world) == 2  # fused path, not the fallback
    b.add_dir_light(old, False, world)
    b.add_dir_light(new, True, world)
    # differences: per axis pass the |delta| > 1e-3 write gate (a skips small deltas, b gates each light separately)
    # plus Change's missing saturate gate at the volume faces -> at most ~2e-3 per pass, two passes
    assert np.abs(a.light - b.light).max() < 4e-3
    assert np.abs(a.light - b.light).mean() < 2e-4


def test_change_falls_back_when_axes_differ():
    n = 16
    vol = oracle.OracleVolume(synth.sphere_volume((n, n, n)), oracle.default_tf())
    code = vol.change_dir_light(synth.LIGHTS[0], synth.LIGHTS[1], synth.identity_world())
    assert code >= 100  # Remove + Add


def test_zero_direction_light_is_a_no_op():
    n = 8
    vol = oracle.OracleVolume(synth.sphere_volume((n, n, n)), oracle.default_tf())
    assert vol.add_dir_light(FDirLightParameters((0, 0, 0), 1.0), True, synth.identity_world()) == 0
    assert not vol.light.any()


def test_g8_light_volume_quantises_every_store():
    n = 16
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    v32 = oracle.OracleVolume(_homogeneous(n, 100), tf, FWindowingParameters(), border_exact=True)
    v8 = oracle.OracleVolume(_homogeneous(n, 100), tf, FWindowingParameters(), light32=False, border_exact=True)
    light = FDirLightParameters((0, 0, -1), 0.8)
    v32.add_dir_light(light, True, synth.identity_world())
    v8.add_dir_light(light, True, synth.identity_world())
    assert v8.light.dtype == np.uint8
    # 8-bit propagation buffers accumulate at most half an LSB of error per slice
    assert np.abs(v8.light.astype(np.float32) / 255.0 - v32.light).max() < (n * 0.5 + 1) / 255.0


def test_single_ray_through_homogeneous_slab():
    n, steps = 16, 64.0
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    w = FWindowingParameters()
    vol = oracle.OracleVolume(_homogeneous(n, 128), tf, w)
    vol.light[:] = 0.5
    # camera on the -X axis looking along +X: the centre pixel's ray crosses the unit cube with thickness 1
    cam = FCamera((-2.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 1.0, 1, 1, 0.0, 0, False)
    rgba, nsteps = vol.raymarch_lit(cam, synth.identity_world(), steps)
    s = oracle.sample_windowed_tf(128 / 255.0, 100.0 / steps, tf, w)
    A, rgb, n_exec = 0.0, np.zeros(3), 0
    for i in range(int(steps)):
        n_exec += 1
        rgb += s[:3] * 0.5 * s[3] * (1 - A)
        A += s[3] * (1 - A)
        if A > 0.95:
            A = 1.0
            break
    assert nsteps in (n_exec, n_exec + 1)  # +1 when thickness*steps leaves a fractional final step
    assert rgba[0, 0, 3] == pytest.approx(A, abs=1e-5)
    assert np.allclose(rgba[0, 0, :3], rgb, atol=1e-5)


def test_ray_that_misses_the_cube_is_transparent():
    vol = oracle.OracleVolume(_homogeneous(8, 255), oracle.default_tf())
    cam = FCamera((-2.0, 0.0, 0.0), (-2.0, 5.0, 0.0), (0.0, 0.0, 1.0), 10.0, 4, 4, 0.0, 0, True)
    rgba, nsteps = vol.raymarch_lit(cam, synth.identity_world(), 32.0)
    assert not rgba.any() and nsteps == 0


def test_clip_plane_removes_half_of_the_opacity():
    n = 16
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    vol = oracle.OracleVolume(_homogeneous(n, 128), tf, FWindowingParameters())
    vol.light[:] = 1.0
    cam = FCamera((-2.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 1.0, 1, 1, 0.0, 0, False)
    full, _ = vol.raymarch_lit(cam, synth.identity_world(), 8.0)
    # keep only x > 0 (local): the direction is the side that is NOT clipped away
    clipped = FRaymarchWorldParameters(FTransform(), FClippingPlaneParameters((0.0, 0.0, 0.0), (1.0, 0.0, 0.0)))
    half, nsteps = vol.raymarch_lit(cam, clipped, 8.0)
    a1 = float(oracle.sample_windowed_tf(128 / 255.0, 100.0 / 8.0, tf, FWindowingParameters())[3])
    assert full[0, 0, 3] == pytest.approx(1 - (1 - a1) ** 8, abs=1e-5) or full[0, 0, 3] == 1.0
    assert half[0, 0, 3] == pytest.approx(1 - (1 - a1) ** 4, abs=1e-5) or half[0, 0, 3] == 1.0
    assert nsteps >= 8  # clipped steps are still executed


def test_cube_setup_centre_ray():
    cam = FCamera((-2.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 1.0, 1, 1, 0.0, 0, False)
    et = oracle.cube_setup(cam, synth.identity_world())[0, 0]
    assert np.allclose(et, [0.0, 0.5, 0.5, 1.0], atol=1e-5)
    # scene geometry in front of the exit point truncates the thickness (local depth 1.5 + 0.3)
    cam.SceneDepth = 1.8
    et = oracle.cube_setup(cam, synth.identity_world())[0, 0]
    assert et[3] == pytest.approx(0.3, abs=1e-5)


def test_mandelbulb_far_ray_misses_and_centre_ray_hits():
    from tbraymarcherplugin_b200.raymarch_utils import FMandelbulbParameters

    cam = FCamera((-2.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 30.0, 9, 9, 0.0, 0, False)
    out, iters = oracle.mandelbulb(FMandelbulbParameters(MaxSteps=256.0), cam, synth.identity_world())
    assert out[4, 4, 1] == 1.0 and 0.0 < out[4, 4, 0] <= 1.0  # centre pixel hits the bulb
    assert out[0, 0, 1] == 0.0  # corner pixel leaves the cube
    assert iters > 0


def test_oracle_synth_generator_equals_numpy_twin():
    for kind, fn in (("sphere", synth.sphere_volume), ("perlin", synth.perlin_ct_volume)):
        dims = (24, 20, 28)
        assert np.array_equal(oracle.synth_volume(kind, dims), fn(dims)), kind


def test_mandelbulb_power8_twin_stays_close_to_the_reference_formulation():
    """The kernels iterate Power == 8 without transcendentals (angle doubling, r^8 by squaring: csrc/mandelbulb.cu); the oracle carries the
    same arithmetic as variant 1. Against the reference's formulation (variant 0, bit-identical to SDFMarcher.usf compiled for the CPU) it
    differs by rounding only, which the iteration amplifies next to the surface: well under 1 % of the pixels / voxels move by more than
    1e-4 — the same order as libm-vs-CUDA transcendentals, and far inside the 2 % budget of the GPU tests."""
    from tbraymarcherplugin_b200.raymarch_utils import FMandelbulbParameters

    L = oracle.lib()
    world, cam = synth.identity_world(), synth.benchmark_camera(240, 135, jitter=False)
    mb = FMandelbulbParameters(MaxSteps=256.0, MaxIterations=16.0)
    try:
        L.tbo_set_mandelbulb_variant(0)
        d0, i0 = oracle.mandelbulb(mb, cam, world)
        n0, _ = oracle.mandelbulb_normal(mb, 0.01, cam, world)
        s0, _ = oracle.mandelbulb_sdf((40, 36, 32), (0.1, 0.0, -0.05), 2.4, 8.0, False)
        L.tbo_set_mandelbulb_variant(1)
        d1, i1 = oracle.mandelbulb(mb, cam, world)
        n1, _ = oracle.mandelbulb_normal(mb, 0.01, cam, world)
        s1, _ = oracle.mandelbulb_sdf((40, 36, 32), (0.1, 0.0, -0.05), 2.4, 8.0, False)
        other, _ = oracle.mandelbulb(FMandelbulbParameters(MaxSteps=64.0, MaxIterations=8.0, Power=6.0), cam, world)
        L.tbo_set_mandelbulb_variant(0)
        other0, _ = oracle.mandelbulb(FMandelbulbParameters(MaxSteps=64.0, MaxIterations=8.0, Power=6.0), cam, world)
    finally:
        L.tbo_set_mandelbulb_variant(0)
    assert (d0[..., 1] == 1).mean() > 0.3 and not np.array_equal(d0, d1)
    assert (np.abs(d0 - d1).max(-1) > 1e-4).mean() < 0.005 and abs(i0 - i1) / i0 < 1e-3
    assert (n0[..., 3] != n1[..., 3]).mean() < 0.005
    assert (np.abs(s0 - s1) > 1e-4).mean() < 0.01
    assert np.array_equal(other, other0)  # any other power takes the reference's formulation in both variants


def test_joined_same_axis_sweeps_equal_consecutive_adds_up_to_summation_order():
    """SURVEY.md §8(f) row 1 (not in the reference): tbo_add_dir_lights_joined sweeps the passes of all lights that propagate from the same
    cube face once. One light: bit-identical to AddDirLight. Several: fewer sweeps, and a light volume that differs from consecutive
    AddDirLight calls only by the order in which a voxel's contributions are added."""
    from tbraymarcherplugin_b200.raymarch_utils import FDirLightParameters

    data = synth.perlin_ct_volume((40, 32, 24))
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    win = FWindowingParameters(0.45, 0.5, True, False)
    lights = synth.LIGHTS + [synth.rotate_about_z(synth.LIGHTS[0], 7.0), synth.rotate_about_z(synth.LIGHTS[2], -9.0), FDirLightParameters((0, 0, 0), 1.0)]
    for world in (synth.identity_world(), synth.clipped_world()):
        for light32 in (True, False):
            seq = oracle.OracleVolume(data, tf, win, light32=light32)
            joined = oracle.OracleVolume(data, tf, win, light32=light32)
            n_seq = sum(seq.add_dir_light(l, True, world) for l in lights)
            n_joined = oracle.add_dir_lights_joined(joined, lights, True, world)
            assert n_seq == 11 and n_joined < n_seq  # a zero direction adds nothing; same-face passes share a sweep
            d = np.abs(seq.light.astype(np.float64) - joined.light.astype(np.float64))
            assert d.max() <= (4e-6 if light32 else 1.0) and seq.light.max() > (3.0 if light32 else 200)
            one_a, one_b = oracle.OracleVolume(data, tf, win, light32=light32), oracle.OracleVolume(data, tf, win, light32=light32)
            one_a.add_dir_light(lights[0], True, world)
            assert oracle.add_dir_lights_joined(one_b, [lights[0]], True, world) == 2 and np.array_equal(one_a.light, one_b.light)
            # removing what was added restores the volume (the signs mirror AddDirLight's)
            oracle.add_dir_lights_joined(joined, lights, False, world)
            assert np.abs(joined.light.astype(np.float64)).max() <= (1e-5 if light32 else 2.0)
    # more than 8 passes on one face split into two sweeps
    many = [synth.rotate_about_z(synth.LIGHTS[2], 2.0 * i) for i in range(10)]  # L3 is dominated by +Z: all first passes share a face
    v = oracle.OracleVolume(data, tf, win)
    assert oracle.add_dir_lights_joined(v, many, True, synth.identity_world()) >= 3
