"""SURVEY.md §8(f) rows 2-4 — the oracle against the REFERENCE'S OWN CODE for the other materials (intensity march, octree march), the
octree generation, the Mandelbulb variants (normal-returning march, SDF bake) and volume ingest (normalisation, float conversion,
FVolumeInfo mappings). tests/golden/ref_materials.npz / ref_ingest.npz were produced by the reference's shaders and C++ templates compiled
for the CPU (oracle/ref.mk, tests/golden/make_golden_ref.py); with oracle/_ref/libtbrm_ref.so present the vectors are also re-derived
live. Bit-exact, Mandelbulb included: oracle and reference build share libm."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

import oracle
import refpin
from tbraymarcherplugin_b200 import synth
from tbraymarcherplugin_b200.raymarch_utils import FWindowingParameters

GOLDEN = Path(__file__).resolve().parent / "golden"
_spec = importlib.util.spec_from_file_location("make_golden_ref", GOLDEN / "make_golden_ref.py")
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)

needs_ref = pytest.mark.skipif(not refpin.available(), reason="oracle/_ref/libtbrm_ref.so not built and /root/reference absent")


@pytest.mark.parametrize("dims", mk.MATERIAL_DIMS)
def test_octree_and_the_two_other_materials_equal_the_reference(dims):
    want = np.load(GOLDEN / "ref_materials.npz")
    tag = "x".join(map(str, dims))
    data = synth.perlin_ct_volume(dims)
    mips = oracle.generate_octree(data)
    for m, a in enumerate(mips):
        assert a.shape == want[f"octree_{tag}_mip{m}"].shape and np.array_equal(a, want[f"octree_{tag}_mip{m}"]), f"mip {m}"
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    for wname, wv in mk.MATERIAL_WINDOWS.items():
        for world_name, mkw in mk.MATERIAL_WORLDS.items():
            vol = oracle.OracleVolume(data, tf, FWindowingParameters(*wv))
            rgba, steps = oracle.raymarch_intensity(vol, mk.material_camera(), mkw(), 40.0)
            assert np.array_equal(rgba, want[f"intensity_{tag}_{wname}_{world_name}"]) and steps > 0
            for mip in (0, 2):
                rgba, _ = oracle.raymarch_octree(vol, mk.material_camera(), mkw(), 40.0, mips, mip)
                assert np.array_equal(rgba, want[f"octree_march_{tag}_{wname}_{world_name}_mip{mip}"])


def test_octree_semantics():
    data = synth.perlin_ct_volume((20, 9, 5))
    mips = oracle.generate_octree(data)
    assert [m.shape for m in mips] == [(8, 16, 32), (4, 8, 16), (2, 4, 8), (1, 2, 4)]  # pow-2 sides, 4 mips (RaymarchVolume.cpp:873-877)
    assert np.array_equal(mips[0][:5, :9, :20], data.astype(np.uint16) * 257)          # UNORM8 -> UNORM16 of the same value
    assert not mips[0][5:].any() and not mips[0][:, 9:].any() and not mips[0][:, :, 20:].any()  # outside the data volume: Load returns 0
    for m in range(1, 4):  # a max pyramid
        Z, Y, X = mips[m].shape
        assert np.array_equal(mips[m], mips[m - 1].reshape(Z, 2, Y, 2, X, 2).max(axis=(1, 3, 5)))


def test_mandelbulb_variants_equal_the_reference():
    want = np.load(GOLDEN / "ref_materials.npz")
    cam = synth.benchmark_camera(*mk.MANDELBULB_VIEW, jitter=False)
    dist, _ = oracle.mandelbulb(mk.mandelbulb_params(), cam, synth.identity_world())
    assert np.array_equal(dist, want["mandelbulb_distance"])
    nrm, iters = oracle.mandelbulb_normal(mk.mandelbulb_params(), 0.01, cam, synth.identity_world())
    assert np.array_equal(nrm, want["mandelbulb_normal"]) and iters > 0
    hit = (nrm[..., 3] == 1) & (np.abs(nrm[..., :3]).sum(axis=-1) > 0)  # low-precision hits return a black normal
    assert hit.sum() > 100 and np.allclose(np.linalg.norm(nrm[hit][:, :3], axis=-1), 1.0, atol=1e-5)
    sdf16, _ = oracle.mandelbulb_sdf(g16=True, **mk.SDF_CASE)
    sdf32, _ = oracle.mandelbulb_sdf(g16=False, **mk.SDF_CASE)
    assert np.array_equal(sdf16, want["mandelbulb_sdf_g16"]) and np.array_equal(sdf32, want["mandelbulb_sdf_r32f"])
    assert (sdf32 < 0).any() and sdf16.max() > 1000  # inside / outside the bulb


def test_ingest_equals_the_reference():
    want = np.load(GOLDEN / "ref_ingest.npz")
    for key in (0, 1, 2, 3, 4, 5, 6, 16):
        fmt = key % 10
        a = want[f"in_{key}"]
        n, lo, hi = oracle.normalize_array(fmt, a)
        assert n.dtype == want[f"normalized_{key}"].dtype and np.array_equal(n, want[f"normalized_{key}"]), key
        assert np.array_equal(np.array([lo, hi], np.float32), want[f"minmax_{key}"]), key
        if fmt != 6:
            assert np.array_equal(oracle.convert_to_float(fmt, a), want[f"float_{key}"])
    # all-negative float data: the reference's InMax starts at FLT_MIN, so the reported maximum is FLT_MIN, not the true maximum
    assert want["minmax_16"][1] == np.finfo(np.float32).tiny and want["minmax_16"][0] < -1
    # FVolumeInfo::{Normalize,Denormalize}{Value,Range} (VolumeInfo.cpp:18-55) on the default range [-1000, 3000]
    v = want["info_values"]
    lo, hi = np.float32(-1000.0), np.float32(3000.0)
    expect = np.stack([(v - lo) / (hi - lo), (v * (hi - lo)) + lo, v / (hi - lo), v * (hi - lo)]).astype(np.float32)
    assert np.array_equal(expect, want["info_maps"]) and np.array_equal(np.stack([v] * 4), want["info_maps_raw"])


@needs_ref
@pytest.mark.parametrize("case", ["ref_materials", "ref_ingest"])
def test_golden_vectors_are_what_the_reference_build_produces_today(case):
    want = np.load(GOLDEN / f"{case}.npz")
    got = mk.CASES[case]()
    assert set(want.files) == set(got)
    for k in want.files:
        assert np.array_equal(want[k], got[k], equal_nan=True), k


@needs_ref
def test_live_reference_on_other_formats_and_sizes():
    rng = np.random.default_rng(3)
    for dims, dtype in (((33, 17, 70), np.uint8), ((16, 24, 8), np.uint16), ((12, 12, 12), np.float32)):
        base = synth.perlin_ct_volume(dims)
        data = base if dtype == np.uint8 else (base.astype(np.uint16) * 257 if dtype == np.uint16 else (base / np.float32(200)).astype(np.float32))
        a, b = oracle.generate_octree(data), refpin.generate_octree(data)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        tf = oracle.prepare_tf(synth.soft_ct_curve())
        win = FWindowingParameters(float(rng.uniform(0.3, 0.6)), float(rng.uniform(0.3, 1.0)), True, bool(rng.integers(2)))
        va, vb = oracle.OracleVolume(data, tf, win), refpin.RefVolume(data, tf, win)
        cam = synth.benchmark_camera(44, 30, jitter=True, frame=6)
        for world in (synth.scaled_rotated_world(), synth.clipped_world()):
            assert np.array_equal(oracle.raymarch_intensity(va, cam, world, 33.0)[0], vb.raymarch(1, cam, world, 33.0))
            for mip in range(4):
                assert np.array_equal(oracle.raymarch_octree(va, cam, world, 33.0, a, mip)[0], vb.raymarch(2, cam, world, 33.0, octree=b, octree_mip=mip))
