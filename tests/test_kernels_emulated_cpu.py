"""The kernels' own source on the CPU: tbraymarcherplugin_b200/csrc compiled by g++ against the SIMT emulator of tests/emu (fibers for the
threads of a block, real barriers and shuffles, co-resident blocks for the cooperative sweeps, emulated tensor maps / mbarriers / TMA
copies, a device heap with guard pages), driven through the same C ABI and the same test functions as `-m gpu`.

What this covers that the oracle-vs-golden tests cannot: the indexing, bounds, tiling, halo exchange and launch geometry of the CUDA code
itself — on a machine without a GPU. What it cannot: timing, the PTX memory model (emulated accesses are stronger than relaxed), occupancy.
The selection below is sized for the CPU suite; `pytest tests -m gpu --emulate-kernels` runs any GPU test this way."""
import gc

import numpy as np
import pytest

import emu_lib
import oracle
import test_gpu_golden as G
import test_gpu_zz_materials as M0
import test_zzz_gpu_more as M
from tbraymarcherplugin_b200 import synth
from tbraymarcherplugin_b200.raymarch_utils import FDirLightParameters, FSweepStats, FWindowingParameters, URaymarchUtils


@pytest.fixture
def emulated(monkeypatch):
    yield emu_lib.use(monkeypatch)
    gc.collect()  # resource sets of the emulated library are destroyed by it, before the operator surface points at libtbrm.so again


@pytest.mark.parametrize("impl", [1, 2, 3])  # per-slice, TMA-staged fused (asserted to have run), generic fused
def test_emulated_sweep_kernels_equal_golden(emulated, impl):
    G.test_sweep_equals_golden("identity", impl)
    G.test_sweep_equals_golden("clipped", impl)


def test_emulated_lit_raymarch_and_cube_setup_equal_golden(emulated):
    G.test_lit_raymarch_and_cube_setup_equal_golden(1)


def test_emulated_tma_and_fused_sweeps_on_thin_volumes(emulated):
    """X % 16 == 0 volumes one or a few voxels thick: the TMA-staged and the generic fused sweep (which one takes a pass depends on the
    light) against the oracle, AddDirLight with axis-aligned lights and a ChangeDirLight, with and without a clip plane."""
    win = FWindowingParameters(0.45, 0.5, True, False)
    impls = set()
    for dims in [(16, 1, 1), (16, 16, 1), (32, 3, 5), (64, 8, 4)]:
        data = np.random.default_rng(sum(dims)).integers(0, 256, dims[::-1]).astype(np.uint8)
        for world in (synth.identity_world(), synth.clipped_world()):
            res = M0.make_res(data, win)
            vol = oracle.OracleVolume(data, oracle.prepare_tf(synth.soft_ct_curve()), win)
            URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
            for l in synth.LIGHTS + [FDirLightParameters((1, 0, 0), 0.7), FDirLightParameters((0, 1, 0), 0.3), FDirLightParameters((0, 0, -1), 0.3)]:
                st = FSweepStats()
                assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True, stats=st)
                vol.add_dir_light(l, True, world)
                impls |= set(st.impl)
            assert np.array_equal(URaymarchUtils.ReadLightVolume(res), vol.light), dims
            n = synth.rotate_about_z(synth.LIGHTS[0], 20.0)
            assert URaymarchUtils.ChangeDirLightInSingleVolume(res, synth.LIGHTS[0], n, world, bGPUSync=True)
            vol.change_dir_light(synth.LIGHTS[0], n, world)
            assert np.array_equal(URaymarchUtils.ReadLightVolume(res), vol.light), dims
            res.release()
    assert impls == {2, 3}  # both fused schedules took passes


@pytest.mark.parametrize("dims", [(1, 1, 1), (7, 3, 1)])
def test_emulated_kernels_on_degenerate_sizes(emulated, dims):
    M.test_degenerate_and_ragged_sizes_match_oracle(dims, True)


def test_emulated_joined_sweep_kernel(emulated):
    M.test_joined_same_axis_sweeps_match_their_cpu_twin((33, 17, 9), True)
    M.test_joined_same_axis_sweeps_match_their_cpu_twin((33, 17, 9), False)


def test_emulated_second_generation_raymarch(emulated):
    M.test_second_generation_raymarch_on_small_and_degenerate_volumes((1, 7, 1))
    M.test_second_generation_raymarch_on_small_and_degenerate_volumes((5, 4, 6))


def test_emulated_loader_and_ingest_kernels(emulated, tmp_path):
    M.test_headerless_raw_file_loads_like_the_mhd_path(tmp_path)
