"""Z-slab sharding of ONE volume checked on a single GPU: every slab is its own resource set on the same
device ("virtual ranks"), neighbours are connected with same-process arena pointers, and the slabs run each axis pass one
after the other in dependency order (the exchange cells are full depth, so a finished upstream slab has left everything
its neighbour needs). The merged slabs must equal the unsharded sweep BIT FOR BIT. Banded passes (a buffer plane split
into several co-resident waves, what a 1024^2 plane needs on one GPU) are forced on small planes through a test hook."""
import ctypes as C
import os

import numpy as np
import pytest

from tbraymarcherplugin_b200 import FMT_G8, _capi, synth
from tbraymarcherplugin_b200.raymarch_utils import FSweepStats, FWindowingParameters, URaymarchUtils

pytestmark = pytest.mark.gpu
CT_WINDOW = FWindowingParameters(0.45, 0.5, True, False)
SLAB_TIMEOUT_MS = int(os.environ.get("TBRM_TEST_SLAB_TIMEOUT_MS", "1500"))  # the CPU emulator (--emulate-kernels) needs minutes at 256^3
WORLDS = {"identity": synth.identity_world, "scaled_rotated": synth.scaled_rotated_world, "clipped": synth.clipped_world}


_STREAM_OWNERS = []


def make_res(data, sweep_impl=2, band_rows=0, light32=True, half_res=False, flags=0):
    Z, Y, X = data.shape
    res = URaymarchUtils.InitializeRaymarchResources((X, Y, Z), FMT_G8, bLightVolume32Bit=light32, LightVolumeHalfResolution=half_res)
    URaymarchUtils.SetDataVolume(res, data)
    URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(res, CT_WINDOW)
    URaymarchUtils.SetOptions(res, sweep_impl=sweep_impl, debug_flags=(flags, 0, band_rows))
    return res


def unsharded(data, lights, world, light32=True, half_res=False):
    res = make_res(data, light32=light32, half_res=half_res)
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    for l in lights:
        st = FSweepStats()
        assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True, stats=st)
        if set(st.impl) != {3}:
            # e.g. an exactly axis-aligned light on a dimension that is not a power of two (the tap pairs are not uniform):
            # the unsharded sweep falls back to the generic fused kernel, a sharded one reports TBRM_ERR_UNSUPPORTED
            pytest.skip(f"the TMA-staged sweep does not cover light {l.LightDirection} on this volume")
    return URaymarchUtils.ReadLightVolume(res)


def virtual_ranks(data, nranks, band_rows=0, light32=True, half_res=False, flags=0):
    lib = _capi.load()
    Z = data.shape[0]
    ranks = []
    for r in range(nranks):
        res = make_res(data, band_rows=band_rows, light32=light32, half_res=half_res, flags=flags)
        z0, z1 = C.c_int32(), C.c_int32()
        lib.tbrm_slab_partition(res.LightDims[2], nranks, r, C.byref(z0), C.byref(z1))  # slabs are slices of the LIGHT volume
        slab = _capi.Slab(r, nranks, z0.value, z1.value)
        _capi.check(lib.tbrm_slab_configure(res.handle, C.byref(slab)))
        _capi.check(lib.tbrm_slab_set_timeout_ms(res.handle, SLAB_TIMEOUT_MS))
        ranks.append((res, z0.value, z1.value))
    # All virtual ranks share ONE stream. On separate streams their cooperative launches run concurrently on the one GPU, and a downstream
    # slab's launch (spinning on its inbox) can take the SM slots an upstream launch still needs for its last tiles: neither finishes until
    # the exchange timeout fires (seen as a 1-in-18 flake of the 8-rank case in round 2). Real ranks own a GPU each (tests/test_gpu_multi.py).
    shared_stream = lib.tbrm_stream(ranks[0][0].handle)
    _STREAM_OWNERS.append(ranks[0][0])  # the stream's owner must outlive every resource set that borrows it
    for res, _, _ in ranks[1:]:
        _capi.check(lib.tbrm_set_stream(res.handle, C.c_void_p(shared_stream)))
    arenas = []
    for res, _, _ in ranks:
        p, n = C.c_void_p(), C.c_size_t()
        _capi.check(lib.tbrm_slab_arena(res.handle, C.byref(p), C.byref(n)))
        arenas.append(p)
    for r, (res, _, _) in enumerate(ranks):
        if r > 0:
            _capi.check(lib.tbrm_slab_set_peer(res.handle, -1, arenas[r - 1]))
        if r + 1 < nranks:
            _capi.check(lib.tbrm_slab_set_peer(res.handle, +1, arenas[r + 1]))
    return ranks


def sharded_sweep(ranks, lights, world, added=True):
    lib = _capi.load()
    w = world.to_c()
    for light in lights:
        l = light.to_c()
        for p in (0, 1):
            order = C.c_int(0)
            _capi.check(lib.tbrm_slab_pass_order(ranks[0][0].handle, C.byref(l), C.byref(w), p, C.byref(order)))
            if order.value == 0:
                continue
            if order.value == 2:
                pytest.skip("footprints reach both ways: the slabs of this pass can only run concurrently")
            for res, _, _ in (ranks if order.value > 0 else ranks[::-1]):
                st = _capi.SweepStats()
                _capi.check(lib.tbrm_add_dir_light_pass(res.handle, C.byref(l), int(added), C.byref(w), p, 1, C.byref(st)))
                assert st.passes == 1 and st.impl[0] == 3
    for res, _, _ in ranks:
        _capi.check(lib.tbrm_slab_check(res.handle))


def merged(ranks):
    out = None
    for res, z0, z1 in ranks:
        L = URaymarchUtils.ReadLightVolume(res)
        if out is None:
            out = np.zeros_like(L)
        out[z0:z1] = L[z0:z1]
    return out


@pytest.mark.parametrize("world_name", list(WORLDS))
@pytest.mark.parametrize("dims,nranks", [((64, 48, 64), 2), ((64, 48, 64), 4), ((128, 64, 96), 3), ((64, 64, 32), 2), ((256, 256, 256), 8)])
def test_sharded_sweep_is_bit_identical(dims, nranks, world_name):
    data = synth.perlin_ct_volume(dims)
    world = WORLDS[world_name]()
    ref = unsharded(data, synth.LIGHTS, world)
    ranks = virtual_ranks(data, nranks)
    for res, _, _ in ranks:
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    sharded_sweep(ranks, synth.LIGHTS, world)
    got = merged(ranks)
    assert ref.max() > 1.0
    d = np.abs(got - ref)
    assert np.array_equal(got, ref), f"{np.count_nonzero(d)} voxels differ, max {d.max():.3e}, first at {np.argwhere(d > 0)[:4].tolist()}"
    # removing a light is the same exchange with the opposite sign
    sharded_sweep(ranks, [synth.LIGHTS[0]], world, added=False)
    res1 = make_res(data)
    URaymarchUtils.WriteLightVolume(res1, ref)
    URaymarchUtils.AddDirLightToSingleVolume(res1, synth.LIGHTS[0], False, world, bGPUSync=True)
    assert np.array_equal(merged(ranks), URaymarchUtils.ReadLightVolume(res1))


@pytest.mark.parametrize("dims,nranks", [((64, 64, 64), 2), ((64, 64, 64), 4), ((128, 64, 96), 3)])
def test_sharded_sweep_of_a_g8_light_volume_is_bit_identical(dims, nranks):
    """The reference's default light-volume format on a sharded volume: byte bricks, exchange cells carrying what the G8 propagation buffers
    would hold, sweeps along X on the (y,z,x)-ordered copy of each rank's light volume; banded on one GPU as well."""
    data = synth.perlin_ct_volume(dims)
    world = synth.identity_world()
    ref = unsharded(data, synth.LIGHTS, world, light32=False)
    ranks = virtual_ranks(data, nranks, light32=False)
    for res, _, _ in ranks:
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    sharded_sweep(ranks, synth.LIGHTS, world)
    got = merged(ranks)
    assert ref.max() > 0.5 and np.array_equal(got, ref), f"{np.count_nonzero(got != ref)} voxels differ"
    sharded_sweep(ranks, [synth.LIGHTS[0]], world, added=False)
    res1 = make_res(data, light32=False)
    URaymarchUtils.WriteLightVolume(res1, ref)
    URaymarchUtils.AddDirLightToSingleVolume(res1, synth.LIGHTS[0], False, world, bGPUSync=True)
    assert np.array_equal(merged(ranks), URaymarchUtils.ReadLightVolume(res1))
    banded = make_res(data, band_rows=3, light32=False)
    URaymarchUtils.ClearResourceLightVolumes(banded, 0.0)
    for l in synth.LIGHTS:
        st = FSweepStats()
        assert URaymarchUtils.AddDirLightToSingleVolume(banded, l, True, world, bGPUSync=True, stats=st)
        assert set(st.impl) == {3} and st.kernel_launches > st.passes
    assert np.array_equal(URaymarchUtils.ReadLightVolume(banded), ref)


@pytest.mark.parametrize("light32", [True, False])
@pytest.mark.parametrize("dims,nranks", [((64, 64, 64), 2), ((128, 64, 64), 4)])
def test_sharded_sweep_of_a_half_resolution_light_volume_is_bit_identical(dims, nranks, light32):
    """LightVolumeHalfResolution on a sharded volume: the slabs are slices of the light volume, every rank's data boxes span twice its tiles."""
    data = synth.perlin_ct_volume(dims)
    world = synth.identity_world()
    ref = unsharded(data, synth.LIGHTS[:3], world, light32=light32, half_res=True)
    ranks = virtual_ranks(data, nranks, light32=light32, half_res=True)
    for res, _, _ in ranks:
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    sharded_sweep(ranks, synth.LIGHTS[:3], world)
    got = merged(ranks)
    assert ref.max() > 0.5 and np.array_equal(got, ref), f"{np.count_nonzero(got != ref)} voxels differ"


@pytest.mark.parametrize("band_rows", [1, 3])
@pytest.mark.parametrize("dims", [(64, 64, 64), (128, 64, 96)])
def test_banded_passes_are_bit_identical(dims, band_rows):
    """One GPU, one volume, but every pass is cut into bands of `band_rows` tile rows that run one after the other."""
    data = synth.perlin_ct_volume(dims)
    world = synth.identity_world()
    ref = unsharded(data, synth.LIGHTS, world)
    res = make_res(data, band_rows=band_rows)
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    for l in synth.LIGHTS:
        st = FSweepStats()
        assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True, stats=st)
        assert set(st.impl) == {3} and st.kernel_launches > st.passes
    assert np.array_equal(URaymarchUtils.ReadLightVolume(res), ref)


@pytest.mark.parametrize("flags,dims", [(512 + 32, (128, 64, 64)), (512 + 16, (128, 64, 96))])  # bits 8-9 = 2: tiles of 7 rows; bits 4-5: two / one pixel per thread
def test_seven_row_tiles_in_banded_and_sharded_passes(flags, dims):
    """Tiles of 7 rows (what fills 148 SMs with four blocks each on 512^2 and 1024^2 planes) in the launches that exchange light between
    bands and slabs: banded passes of one GPU, and the sweeps along Z of a sharded volume (its sweeps along X / Y keep 8-row tiles)."""
    data = synth.perlin_ct_volume(dims)
    world = synth.identity_world()
    lights = synth.LIGHTS[1:] if flags & 32 else synth.LIGHTS  # (the two-pixel form needs uniform tap pairs: not the first light on this volume)
    ref = unsharded(data, lights, world)
    for band_rows in (2, 5):  # (3 would leave a last band of one pixel row, thinner than the footprints reach: such a pass takes the generic kernel)
        res = make_res(data, band_rows=band_rows, flags=flags)
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
        for l in lights:
            st = FSweepStats()
            assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True, stats=st)
            assert set(st.impl) == {3} and st.kernel_launches > st.passes
        assert np.array_equal(URaymarchUtils.ReadLightVolume(res), ref), band_rows
    ranks = virtual_ranks(data, 4 if dims[2] == 64 else 3, flags=flags)
    for res, _, _ in ranks:
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    sharded_sweep(ranks, lights, world)
    assert np.array_equal(merged(ranks), ref)
    ranks = virtual_ranks(data, 2, band_rows=2, flags=flags)
    for res, _, _ in ranks:
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    sharded_sweep(ranks, lights, world)
    assert np.array_equal(merged(ranks), ref)


@pytest.mark.parametrize("flags", [512 + 32, 512 + 16])
@pytest.mark.parametrize("dims,nranks", [((128, 64, 64), 2), ((128, 64, 64), 4), ((64, 64, 96), 3), ((256, 256, 256), 2)])
def test_seven_row_tiles_in_the_slabs_of_sweeps_along_x_and_y(dims, nranks, flags):
    """Slabs of 32 / 16 / 128 slices are not multiples of 7 rows: the tile rows of a slab are anchored at its first row, its last tile row ends
    past the slab (those pixels are masked, the light maps end with the slab, the rows come from the neighbour's exchange cells)."""
    data = synth.perlin_ct_volume(dims)
    world = synth.identity_world()
    lights = synth.LIGHTS[1:] if flags & 32 and dims[0] != dims[1] else synth.LIGHTS
    ref = unsharded(data, lights, world)
    for band_rows in ((0, 2) if dims[2] < 256 else (0,)):
        ranks = virtual_ranks(data, nranks, band_rows=band_rows, flags=flags)
        for res, _, _ in ranks:
            URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
        sharded_sweep(ranks, lights, world)
        got = merged(ranks)
        assert np.array_equal(got, ref), (band_rows, np.count_nonzero(got != ref), np.argwhere(got != ref)[:4].tolist())
        sharded_sweep(ranks, [lights[0]], world, added=False)
        res1 = make_res(data)
        URaymarchUtils.WriteLightVolume(res1, ref)
        URaymarchUtils.AddDirLightToSingleVolume(res1, lights[0], False, world, bGPUSync=True)
        assert np.array_equal(merged(ranks), URaymarchUtils.ReadLightVolume(res1))


def test_sharded_and_banded_together():
    data = synth.perlin_ct_volume((64, 64, 64))
    world = synth.identity_world()
    ref = unsharded(data, synth.LIGHTS[:3], world)
    ranks = virtual_ranks(data, 2, band_rows=2)
    for res, _, _ in ranks:
        URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    sharded_sweep(ranks, synth.LIGHTS[:3], world)
    assert np.array_equal(merged(ranks), ref)


def test_missing_neighbour_times_out_instead_of_hanging():
    lib = _capi.load()
    data = synth.perlin_ct_volume((64, 48, 64))
    world = synth.identity_world()
    ranks = virtual_ranks(data, 2)
    l, w = synth.LIGHTS[3].to_c(), world.to_c()  # (0,0,-1): one pass along Z, the slab that is second in sweep order waits for a hand-off
    order = C.c_int(0)
    _capi.check(lib.tbrm_slab_pass_order(ranks[0][0].handle, C.byref(l), C.byref(w), 0, C.byref(order)))
    late = ranks[-1] if order.value > 0 else ranks[0]
    _capi.check(lib.tbrm_slab_set_timeout_ms(late[0].handle, 100))
    _capi.check(lib.tbrm_add_dir_light_pass(late[0].handle, C.byref(l), 1, C.byref(w), 0, 1, None))
    assert lib.tbrm_slab_check(late[0].handle) == _capi.TBRM_ERR_CUDA
    assert b"timed out" in lib.tbrm_last_error()
    assert lib.tbrm_slab_check(late[0].handle) == _capi.TBRM_OK  # the flag is cleared by the check


def test_sharding_rejects_unsupported_configurations():
    lib = _capi.load()
    res = URaymarchUtils.InitializeRaymarchResources((40, 32, 32), FMT_G8, bLightVolume32Bit=True)  # X % 16 != 0
    slab = _capi.Slab(0, 2, 0, 16)
    assert lib.tbrm_slab_configure(res.handle, C.byref(slab)) == _capi.TBRM_ERR_UNSUPPORTED
    res = URaymarchUtils.InitializeRaymarchResources((32, 32, 32), FMT_G8, bLightVolume32Bit=True)
    bad = _capi.Slab(0, 2, 0, 12)  # not the partition rule
    assert lib.tbrm_slab_configure(res.handle, C.byref(bad)) == _capi.TBRM_ERR_INVALID_ARGUMENT
