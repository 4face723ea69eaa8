"""Pins the oracle — and the product's host math — to the REFERENCE'S OWN CODE for the rows of SURVEY.md §8 whose reference
implementation is plain C++ (a15-a21: LightingShaderUtils.cpp; (f)3: ConvertArrayToNormalizedArray, FVolumeInfo):

  * tests/golden/ref_*.npz hold outputs of those reference functions, compiled unmodified from /root/reference against the
    engine-type shim of oracle/ue_shim (oracle/ref.mk, tests/golden/make_golden_ref.py);
  * where oracle/_ref/libtbrm_ref.so is available (development container, or prebuilt and shipped) the same comparison also
    runs live on fresh random inputs.

The shaders (HLSL) have no such anchor: for them parity stays unpinned.

One documented difference: for an exactly axis-aligned light the reference divides by zero in GetUVOffset /
GetStepSizeAndUVWOffset for the weight-0 second pass (inf / NaN uniforms); policy Q11 replaces them by zeros."""
import ctypes as C
import importlib.util
from pathlib import Path

import numpy as np
import pytest

import oracle
import refpin
from tbraymarcherplugin_b200 import _capi
from tbraymarcherplugin_b200.raymarch_utils import FDirLightParameters, FWindowingParameters, plan_dir_light

GOLDEN = Path(__file__).resolve().parent / "golden"
_spec = importlib.util.spec_from_file_location("make_golden_ref", GOLDEN / "make_golden_ref.py")
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)

needs_ref = pytest.mark.skipif(not refpin.available(), reason="oracle/_ref/libtbrm_ref.so not built and /root/reference absent")


def _lib_plan_vector(dims, win, light, world) -> np.ndarray:
    """tbrm_plan_dir_light of the PRODUCT library in the layout of refpin.plan_to_vector."""
    p = plan_dir_light(dims, win, light, world)
    q = oracle.LightPlan.from_buffer_copy(bytes(p))  # identical layouts (test_host_cpu.py checks the sizes)
    return refpin.plan_to_vector(q)


def _assert_plan_matches_reference(got: np.ndarray, ref: np.ndarray, what: str):
    finite = np.isfinite(ref)
    assert np.array_equal(got[finite], ref[finite]), f"{what}: differs from the reference at {np.nonzero(finite & (got != ref))[0]}"
    # Q11: where the reference produced inf / NaN (weight-0 pass of an axis-aligned light) the restatement carries zeros
    assert np.all(got[~finite] == 0.0), what


def test_oracle_and_product_host_math_equal_the_reference_golden_vectors():
    g = np.load(GOLDEN / "ref_hostmath.npz")
    win = FWindowingParameters()
    n_nonfinite = 0
    for i, (d, l, inten, w, ref) in enumerate(zip(g["dims"], g["dirs"], g["intensity"], g["worlds"], g["plans"])):
        dims, light, world = tuple(int(x) for x in d), FDirLightParameters(tuple(l), float(inten)), mk.world_from_row(w)
        _assert_plan_matches_reference(refpin.plan_to_vector(oracle.plan_dir_light(dims, win, light, world)), ref, f"oracle, case {i}")
        _assert_plan_matches_reference(_lib_plan_vector(dims, win, light, world), ref, f"libtbrm.so, case {i}")
        n_nonfinite += int(not np.all(np.isfinite(ref)))
    assert len(g["plans"]) > 300 and 0 < n_nonfinite < 20


def test_reference_golden_vectors_cover_the_interesting_cases():
    g = np.load(GOLDEN / "ref_hostmath.npz")
    plans = g["plans"]
    assert (plans[:, 0] == 1).sum() >= 1                      # a zero light direction: nothing happens
    assert {0, 1, 2} <= set(plans[:, 1].astype(int).tolist())  # Add runs 0, 1 or 2 axis passes
    faces = plans[:, 11].astype(int)                           # first pass face
    assert set(faces.tolist()) == {0, 1, 2, 3, 4, 5}
    # GetPermutationMatrix: pos = px*row0 + py*row1 + Loop*row2 -> X sweeps (y,z,x), Y sweeps (x,z,y), Z sweeps identity
    perm = g["permutation_rows"]
    assert perm[0].tolist() == [[0, 1, 0], [0, 0, 1], [1, 0, 0]] and perm[2].tolist() == [[1, 0, 0], [0, 0, 1], [0, 1, 0]]
    assert np.array_equal(perm[4], np.eye(3)) and np.array_equal(perm[0], perm[1]) and np.array_equal(perm[2], perm[3])


@needs_ref
def test_golden_vectors_are_what_the_reference_build_produces_today():
    want = np.load(GOLDEN / "ref_hostmath.npz")
    got = mk.hostmath_case()
    for k in want.files:
        assert np.array_equal(want[k], got[k], equal_nan=True), k


@needs_ref
def test_oracle_equals_live_reference_on_fresh_random_inputs():
    rng = np.random.default_rng()  # fresh inputs every run, on purpose
    seed = int(rng.integers(1 << 31))
    rng = np.random.default_rng(seed)
    from tbraymarcherplugin_b200.raymarch_utils import FClippingPlaneParameters, FRaymarchWorldParameters, FTransform
    for i in range(500):
        t = FTransform.from_axis_angle(tuple(rng.standard_normal(3)), float(rng.uniform(-180, 180)), tuple(rng.uniform(-50, 50, 3)),
                                       tuple(rng.uniform(0.2, 4.0, 3)))
        world = FRaymarchWorldParameters(t, FClippingPlaneParameters(tuple(rng.uniform(-40, 40, 3)), tuple(rng.standard_normal(3))))
        dims = tuple(int(x) for x in rng.integers(3, 2048, 3))
        light = FDirLightParameters(tuple(rng.standard_normal(3) * 2), float(rng.uniform(0.01, 2.0)))
        ref = refpin.plan_to_vector(refpin.plan_dir_light(dims, light, world))
        _assert_plan_matches_reference(refpin.plan_to_vector(oracle.plan_dir_light(dims, FWindowingParameters(), light, world)), ref,
                                       f"seed {seed}, case {i}")


@pytest.mark.skipif(not refpin.REFERENCE.exists(), reason="/root/reference absent")
def test_reference_sources_are_the_ones_the_golden_vectors_were_made_from():
    """tests/golden/ref_sources.json records the SHA-256 of every reference file oracle/ref.mk compiles; the checkout here must still match
    (otherwise the committed ref_*.npz / ref_fullsize_hashes.json describe another revision and have to be regenerated)."""
    import json

    want = json.loads((GOLDEN / "ref_sources.json").read_text())["sha256"]
    assert want == mk.reference_source_digests(), "reference sources changed: re-run tests/golden/make_golden_ref.py and make_golden_ref_fullsize.py"
