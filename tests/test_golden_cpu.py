"""The committed vectors of tests/golden/ (written by tests/golden/make_golden.py from the CPU oracle) against the oracle as
built today — a drift check of the restatement, its compiler flags and the synthetic input generators — and against the
product's host parameter math (libtbrm.so, no GPU needed). The reference ships no golden vectors: these pin
the ORACLE, not the reference."""
import ctypes as C
import importlib.util
from pathlib import Path

import numpy as np
import pytest

import oracle
from tbraymarcherplugin_b200 import _capi, synth
from tbraymarcherplugin_b200.raymarch_utils import plan_dir_light

GOLDEN = Path(__file__).resolve().parent / "golden"
_spec = importlib.util.spec_from_file_location("make_golden", GOLDEN / "make_golden.py")
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


@pytest.mark.parametrize("case", list(make_golden.CASES))
def test_oracle_reproduces_golden_vectors_bit_for_bit(case):
    want = np.load(GOLDEN / f"{case}.npz")
    got = make_golden.CASES[case]()
    assert set(want.files) == set(got)
    for k in want.files:
        assert want[k].dtype == got[k].dtype and np.array_equal(want[k], got[k]), f"{case}/{k} drifted"


def test_host_plan_of_the_library_equals_golden_plans():
    want = np.load(GOLDEN / "plans.npz")
    assert C.sizeof(_capi.LightPlan) == want["identity_L0"].size
    for name, mk in make_golden.WORLDS.items():
        for i, l in enumerate(synth.LIGHTS):
            p = plan_dir_light(make_golden.PLAN_DIMS, make_golden.CT_WINDOW, l, mk())
            assert bytes(p) == want[f"{name}_L{i}"].tobytes(), f"{name} L{i}"


def test_golden_vectors_are_not_trivial():
    s = np.load(GOLDEN / "sweep_32.npz")
    assert s["identity_reset"].max() > 1.5 and not np.array_equal(s["identity_reset"], s["identity_removed"])
    assert not np.array_equal(s["identity_removed"], s["identity_changed"])
    r = np.load(GOLDEN / "raymarch_32.npz")
    assert r["rgba_jitter1"][..., 3].max() > 0.3 and int(r["steps_jitter0"][0]) > 10000
    assert (np.load(GOLDEN / "mandelbulb_32x24.npz")["out"][..., 1] == 1).sum() > 20
