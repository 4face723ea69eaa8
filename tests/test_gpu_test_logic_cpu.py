"""The Python logic of GPU tests that were written without a GPU at hand, exercised on the CPU against an oracle-backed stand-in for the
operator surface (tests/fake_ops.py): plumbing, shapes, tolerances and file handling of the test code itself. Nothing here says anything
about the kernels — the same functions run against the CUDA path under `-m gpu`."""
import pytest

import fake_ops
import test_gpu_zz_materials as M0
import test_zzz_gpu_more as M
import test_ref_fullsize as F
from tbraymarcherplugin_b200 import raymarch_utils as RU
from tbraymarcherplugin_b200 import raymarch_volume as RV


@pytest.fixture
def fake_surface(monkeypatch):
    for mod in (M0, M, F):
        monkeypatch.setattr(mod, "URaymarchUtils", fake_ops.FakeRaymarchUtils)
    monkeypatch.setattr(RU, "UMHDLoader", fake_ops.FakeMHDLoader)
    monkeypatch.setattr(RU, "UVolumeTextureToolkit", fake_ops.FakeVolumeTextureToolkit)
    init = RV.ARaymarchVolume.__init__
    monkeypatch.setattr(init, "__defaults__", tuple(fake_ops.FakeRaymarchUtils if d is RU.URaymarchUtils else d for d in init.__defaults__))


@pytest.mark.parametrize("dims", [(1, 1, 1), (7, 3, 1)])
def test_logic_of_the_degenerate_size_test(fake_surface, dims):
    M.test_degenerate_and_ragged_sizes_match_oracle(dims, True)


def test_logic_of_the_mandelbulb_twin_test(fake_surface):
    M.test_mandelbulb_power8_kernels_match_their_cpu_twin()


def test_logic_of_the_raw_loader_and_actor_tests(fake_surface, tmp_path):
    (tmp_path / "a").mkdir(), (tmp_path / "b").mkdir()
    M.test_headerless_raw_file_loads_like_the_mhd_path(tmp_path / "a")
    M.test_raymarch_volume_actor_from_mhd_file_ticks_and_renders_every_material(tmp_path / "b")


def test_logic_of_the_small_volume_raymarch_v2_test(fake_surface):
    M.test_second_generation_raymarch_on_small_and_degenerate_volumes((1, 7, 1))


def test_logic_of_the_joined_lights_test(fake_surface):
    M.test_joined_same_axis_sweeps_match_their_cpu_twin((33, 17, 9), True)
    M.test_joined_same_axis_sweeps_match_their_cpu_twin((33, 17, 9), False)


def test_logic_of_the_full_size_digest_test(fake_surface):
    F.test_cuda_path_equals_the_reference_shaders_at_full_size("cfg1")
