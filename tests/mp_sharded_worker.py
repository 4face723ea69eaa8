"""Worker of tests/test_gpu_multi.py (one process per GPU, launched by torchrun): ONE volume Z-slab sharded over the ranks.
Checks, bit for bit, against the unsharded sweep / raymarch that rank 0 also runs on its own GPU:
the gathered light volume after a full reset with all four lights and after removing one, and the gathered frame."""
import ctypes as C
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np
import torch
import torch.distributed as dist

from tbraymarcherplugin_b200 import FMT_G8, _capi, sharding, synth
from tbraymarcherplugin_b200.raymarch_utils import FSweepStats, FWindowingParameters, URaymarchUtils


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rank, world_size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _capi.load()
    win = FWindowingParameters(0.45, 0.5, True, False)
    world = synth.identity_world()
    cam = synth.benchmark_camera(320, 200)

    d_vol = torch.empty((n, n, n), dtype=torch.uint8, device="cuda")
    _capi.check(lib.tbrm_synth_volume_u8(local, _capi.SYNTH_PERLIN_CT, (C.c_int32 * 3)(n, n, n), synth.PERLIN_SEED, C.c_void_p(d_vol.data_ptr()), 1))

    vol = sharding.FShardedRaymarchVolume((n, n, n), local)
    URaymarchUtils.ColorCurveToTexture(vol.res, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(vol.res, win)
    vol.SetDataVolumeSlab(d_vol[vol.z0:vol.z1])
    vol.Flush()
    assert torch.equal(vol.data, d_vol), "all-gather of the data slabs"

    for rep in range(2):  # twice: the second round reuses both arena regions
        vol.ClearLightVolume(0.0)
        for l in synth.LIGHTS:
            st = FSweepStats()
            assert vol.AddDirLight(l, True, world, stats=st)
            assert set(st.impl) == {3}, st.impl
    vol.AddDirLight(synth.LIGHTS[1], False, world)
    # incremental updates (cfg 3): every remaining light is rotated 5 degrees about +Z per update, as ONE ChangeDirLight each
    current = [synth.LIGHTS[0], synth.LIGHTS[2], synth.LIGHTS[3]]
    for step in (1, 2):
        for i, base in enumerate((synth.LIGHTS[0], synth.LIGHTS[2], synth.LIGHTS[3])):
            new = synth.rotate_about_z(base, 5.0 * step)
            assert vol.ChangeDirLight(current[i], new, world)
            current[i] = new
    vol.GatherLightVolume()
    frame, _ = vol.Render(cam, world, 200.0)
    vol.Check()

    if rank == 0:
        ref = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=True, device=local)
        URaymarchUtils.SetDataVolumeDevice(ref, d_vol.data_ptr())
        URaymarchUtils.ColorCurveToTexture(ref, synth.soft_ct_curve())
        URaymarchUtils.SetWindowingParameters(ref, win)
        URaymarchUtils.ClearResourceLightVolumes(ref, 0.0)
        for l in synth.LIGHTS:
            URaymarchUtils.AddDirLightToSingleVolume(ref, l, True, world, bGPUSync=True)
        URaymarchUtils.AddDirLightToSingleVolume(ref, synth.LIGHTS[1], False, world, bGPUSync=True)
        cur = [synth.LIGHTS[0], synth.LIGHTS[2], synth.LIGHTS[3]]
        for step in (1, 2):
            for i, base in enumerate((synth.LIGHTS[0], synth.LIGHTS[2], synth.LIGHTS[3])):
                new = synth.rotate_about_z(base, 5.0 * step)
                URaymarchUtils.ChangeDirLightInSingleVolume(ref, cur[i], new, world, bGPUSync=True)
                cur[i] = new
        L_ref = URaymarchUtils.ReadLightVolume(ref)
        L = vol.light.cpu().numpy()
        d = np.abs(L - L_ref)
        assert np.array_equal(L, L_ref), f"light volume: {np.count_nonzero(d)} voxels differ, max {d.max():.3e}, first {np.argwhere(d > 0)[:3].tolist()}"
        img_ref, _ = URaymarchUtils.PerformWindowedLitRaymarch(ref, cam, world, 200.0)
        img = frame.cpu().numpy()
        assert img_ref[..., 3].max() > 0.5
        assert np.array_equal(img, img_ref), f"frame: max diff {np.abs(img - img_ref).max():.3e}"
    else:
        assert frame is None
    dist.barrier()

    # push-gather: a full reset whose LAST light pushes its finished bricks into every rank's volume from inside the sweep kernel; no
    # all-gather follows (GatherLightVolume only synchronises). Every rank must hold the unsharded result bit for bit.
    vol.light.fill_(-7.0)  # whatever the peers do not deliver stays recognisable
    torch.cuda.synchronize()
    dist.barrier()
    vol.ClearLightVolume(0.0)
    for i, l in enumerate(synth.LIGHTS):
        assert vol.AddDirLight(l, True, world, push=(i == len(synth.LIGHTS) - 1))
    vol.GatherLightVolume()
    frame2, _ = vol.Render(cam, world, 200.0)
    vol.Check()
    vol.Flush()
    ref2 = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=True, device=local)
    URaymarchUtils.SetDataVolumeDevice(ref2, d_vol.data_ptr())
    URaymarchUtils.ColorCurveToTexture(ref2, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(ref2, win)
    URaymarchUtils.ClearResourceLightVolumes(ref2, 0.0)
    for l in synth.LIGHTS:
        URaymarchUtils.AddDirLightToSingleVolume(ref2, l, True, world, bGPUSync=True)
    L2_ref = URaymarchUtils.ReadLightVolume(ref2)
    L2 = vol.light.cpu().numpy()
    d2 = np.abs(L2 - L2_ref)
    assert np.array_equal(L2, L2_ref), f"push-gather, rank {rank}: {np.count_nonzero(d2)} voxels differ, max {d2.max():.3e}, first {np.argwhere(d2 > 0)[:3].tolist()}"
    if rank == 0:
        img2_ref, _ = URaymarchUtils.PerformWindowedLitRaymarch(ref2, cam, world, 200.0)
        assert np.array_equal(frame2.cpu().numpy(), img2_ref), "frame after push-gather"
    ref2.release()
    dist.barrier()
    vol.release()
    # the other light-volume settings of the reference on a sharded volume: G8 (its default format), half resolution, both
    # (light-volume slabs must be multiples of 8 slices: half resolution needs n / 2 / world_size % 8 == 0)
    for light32, half in ((False, False), (True, True), (False, True)):
        if half and (n // 2) % (8 * world_size) != 0:
            continue
        other_formats(lib, d_vol, n, local, rank, world, win, cam, light32, half)
    dist.destroy_process_group()
    print(f"sharded ok rank {rank}/{world_size} n={n}", flush=True)


def other_formats(lib, d_vol, n, local, rank, world, win, cam, light32, half):
    vol = sharding.FShardedRaymarchVolume((n, n, n), local, bLightVolume32Bit=light32, LightVolumeHalfResolution=half)
    URaymarchUtils.ColorCurveToTexture(vol.res, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(vol.res, win)
    vol.SetDataVolumeSlab(d_vol[vol.z0:vol.z1])
    vol.Flush()
    vol.ClearLightVolume(0.0)
    for i, l in enumerate(synth.LIGHTS):
        st = FSweepStats()
        assert vol.AddDirLight(l, True, world, stats=st, push=(i == len(synth.LIGHTS) - 1))  # (pushes only an R32F volume)
        assert set(st.impl) == {3}, st.impl
    new = synth.rotate_about_z(synth.LIGHTS[0], 15.0)
    assert vol.ChangeDirLight(synth.LIGHTS[0], new, world)
    vol.GatherLightVolume()
    frame, _ = vol.Render(cam, world, 200.0)
    vol.Check()
    if rank == 0:
        ref = URaymarchUtils.InitializeRaymarchResources((n, n, n), FMT_G8, bLightVolume32Bit=light32, LightVolumeHalfResolution=half, device=local)
        URaymarchUtils.SetDataVolumeDevice(ref, d_vol.data_ptr())
        URaymarchUtils.ColorCurveToTexture(ref, synth.soft_ct_curve())
        URaymarchUtils.SetWindowingParameters(ref, win)
        URaymarchUtils.ClearResourceLightVolumes(ref, 0.0)
        for l in synth.LIGHTS:
            URaymarchUtils.AddDirLightToSingleVolume(ref, l, True, world, bGPUSync=True)
        URaymarchUtils.ChangeDirLightInSingleVolume(ref, synth.LIGHTS[0], new, world, bGPUSync=True)
        L_ref, L = URaymarchUtils.ReadLightVolume(ref), vol.light.cpu().numpy()
        assert L_ref.max() > 0 and np.array_equal(L, L_ref), f"light32={light32} half={half}: {np.count_nonzero(L != L_ref)} light voxels differ"
        img_ref, _ = URaymarchUtils.PerformWindowedLitRaymarch(ref, cam, world, 200.0)
        assert np.array_equal(frame.cpu().numpy(), img_ref), f"light32={light32} half={half}: frame differs"
        ref.release()
    dist.barrier()
    vol.release()


if __name__ == "__main__":
    main()
