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


# a protocol bug would show as a spin that never ends inside the emulated library (C code): bound it from outside
pytestmark = pytest.mark.timeout(900, method="thread")


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
    impls = set()
    # the second window rejects bytes up to 178: the other form of the exact empty-space test on tap bytes (threshold >= 128)
    for dims, win in [((16, 1, 1), FWindowingParameters(0.45, 0.5, True, False)), ((16, 16, 1), FWindowingParameters(0.8, 0.2, True, True)),
                      ((32, 3, 5), FWindowingParameters(0.45, 0.5, True, False)), ((64, 8, 4), FWindowingParameters(0.8, 0.2, True, False))]:
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


@pytest.mark.parametrize("dims,nranks,px_flag", [((64, 48, 64), 4, 0), ((64, 48, 64), 4, 48), ((64, 64, 96), 3, 0), ((64, 64, 64), 8, 48)])
def test_emulated_concurrent_slabs_equal_the_unsharded_sweep(emulated, dims, nranks, px_flag):
    """Z-slab sharding with the ranks running CONCURRENTLY, one thread per virtual rank — what N GPUs do, and what one GPU
    cannot (co-resident cooperative kernels of several ranks would starve each other): every rank issues its passes on its own and waits for its neighbours in the kernels, middle ranks have two neighbours, a
    rank may run a pass ahead of its neighbour (the ack protocol keeps it from overwriting exchange cells), slabs are swept in both orders;
    AddDirLight and its removal; with two pixels per thread and with the automatic choice (one pixel per thread for slabs this small)."""
    import ctypes as C
    import threading

    import test_gpu_slab as S
    from tbraymarcherplugin_b200 import _capi

    data = synth.perlin_ct_volume(dims)
    world = synth.identity_world()
    # both sweep orders over the slabs (higher first for -z lights, lower first for +z) and a light in the slab plane
    lights = synth.LIGHTS[:2] + [FDirLightParameters((0.3, -0.2, 0.93), 0.6), FDirLightParameters((0.9, 0.35, 0.0), 0.8)]
    ref_res = S.make_res(data)
    URaymarchUtils.ClearResourceLightVolumes(ref_res, 0.0)
    for l in lights:
        st = FSweepStats()
        assert URaymarchUtils.AddDirLightToSingleVolume(ref_res, l, True, world, bGPUSync=True, stats=st)
        assert set(st.impl) == {3}
    ref = URaymarchUtils.ReadLightVolume(ref_res)
    ranks = S.virtual_ranks(data, nranks)
    for res, _, _ in ranks:  # px_flag 48: the automatic choice of pixels per thread (one, at these sizes); 0: two
        URaymarchUtils.SetOptions(res, sweep_impl=2, debug_flags=(px_flag, 0, 0))
    orders = set()
    for l in lights:
        for p in (0, 1):
            o = C.c_int(0)
            _capi.check(emulated.tbrm_slab_pass_order(ranks[0][0].handle, C.byref(l.to_c()), C.byref(world.to_c()), p, C.byref(o)))
            orders.add(o.value)
    assert {-1, 1} <= orders, orders
    errors = []

    def run(rank, added, which):
        try:
            res = ranks[rank][0]
            _capi.check(emulated.tbrm_slab_set_timeout_ms(res.handle, 120000))
            for l in which:
                st = FSweepStats()
                assert URaymarchUtils.AddDirLightToSingleVolume(res, l, added, world, bGPUSync=True, stats=st)
                assert set(st.impl) == {3}
            _capi.check(emulated.tbrm_slab_check(res.handle))
        except BaseException as e:  # noqa: BLE001 - reported by the main thread
            errors.append((rank, repr(e)))

    def all_ranks(added, which):
        threads = [threading.Thread(target=run, args=(r, added, which)) for r in range(nranks)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors

    for res, _, _ in ranks:
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    all_ranks(True, lights)
    assert ref.max() > 1.0 and np.array_equal(S.merged(ranks), ref)
    all_ranks(False, lights[:1])
    URaymarchUtils.AddDirLightToSingleVolume(ref_res, lights[0], False, world, bGPUSync=True)
    assert np.array_equal(S.merged(ranks), URaymarchUtils.ReadLightVolume(ref_res))


def test_emulated_slabs_of_g8_and_half_resolution_light_volumes(emulated):
    import test_gpu_slab as S

    S.test_sharded_sweep_of_a_g8_light_volume_is_bit_identical((64, 64, 64), 2)
    S.test_sharded_sweep_of_a_half_resolution_light_volume_is_bit_identical((64, 64, 64), 2, False)
    S.test_sharded_sweep_of_a_half_resolution_light_volume_is_bit_identical((128, 64, 64), 4, True)


@pytest.mark.parametrize("dims", [(16, 1, 1), (48, 20, 9), (144, 16, 24), (40, 12, 8)])
def test_emulated_octree_build_kernels(emulated, dims):
    """octree_build_u8x16_kernel (R8 data, X % 16 == 0: 16-byte loads, byte replication instead of the float round trip) and the generic
    kernel ((40, 12, 8)) against the oracle, all four mips."""
    data = np.random.default_rng(sum(dims)).integers(0, 256, dims[::-1]).astype(np.uint8)
    res = M0.make_res(data, FWindowingParameters(0.45, 0.5, True, False))
    URaymarchUtils.GenerateOctree(res)
    for m, want in enumerate(oracle.generate_octree(data)):
        assert np.array_equal(URaymarchUtils.ReadOctreeMip(res, m), want), m
    res.release()


@pytest.mark.parametrize("dims", [(16, 16, 8), (144, 80, 40), (48, 33, 17), (40, 24, 16)])
def test_emulated_brick_grid_and_axis_replica(emulated, dims):
    """the structures derived from every uploaded volume (two-launch brick grid, 16-byte (y,z,x) transpose; byte-wise kernels on other
    sizes) against their definitions in numpy"""
    M.derived_structures_match_numpy(dims)


def test_emulated_lit_march_in_both_addressing_forms(emulated):
    M.test_lit_march_with_64_bit_tap_addressing_still_matches_oracle((40, 24, 56))
    M.test_lit_march_with_64_bit_tap_addressing_still_matches_oracle((1, 7, 1))


def test_emulated_interleaved_rows_of_a_frame(emulated):
    M.test_interleaved_rows_of_a_frame_equal_the_whole_frame(70)


def test_emulated_choice_of_seven_row_tiles_on_a_512_squared_plane(emulated):
    M.test_a_512_squared_plane_takes_seven_row_tiles_and_stays_bit_exact()


@pytest.mark.parametrize("light32", [True, False])
def test_emulated_half_resolution_light_volume_through_the_tma_staged_sweep(emulated, light32):
    M.test_half_resolution_light_volume_through_the_tma_staged_sweep((64, 64, 40), light32)


@pytest.mark.parametrize("px_flag", [16, 32, 48, 256 + 32, 512 + 16, 512 + 32])
def test_emulated_tma_sweep_with_one_and_two_pixels_per_thread(emulated, px_flag):
    M.test_tma_sweep_with_one_and_two_pixels_per_thread((80, 24, 16), px_flag)
    M.test_tma_sweep_with_one_and_two_pixels_per_thread((64, 48, 40), px_flag)


def test_emulated_cpp_example_end_to_end(emulated, tmp_path):
    """examples/mhd_to_frame.cpp linked against the emulated build of the C ABI: file -> resources -> sweep -> octree -> three materials, from C++"""
    from pathlib import Path

    M.test_cpp_example_runs_end_to_end(tmp_path, Path(emulated._name).parent, "tbrm_emu", view=("96", "54", "40"))


def test_emulated_cpp_actor_mirror_end_to_end(emulated, tmp_path):
    """the C++ host mirror (ARaymarchVolume::Tick over URaymarchUtils over the C ABI) against the emulated build: reset, incremental
    ChangeDirLight and a lit frame from C++, bit-identical to the oracle"""
    from pathlib import Path

    M.test_cpp_actor_mirror_end_to_end(tmp_path, Path(emulated._name).parent, "tbrm_emu")


def test_emulated_streaming_entry_points(emulated):
    """double-buffered upload / present / asynchronous frame download (tests/test_gpu_streaming.py) with plain host buffers: the state
    machine of the streaming API gives, frame by frame, what the synchronous calls give"""
    import test_gpu_streaming as T

    T.run_streaming(lambda v: v, lambda shape: np.empty(shape, np.float32), dims=(32, 32, 32), view=(48, 32), steps=40.0)
