"""Streaming entry points (double-buffered upload, asynchronous frame download): a sequence of different volumes pushed through
tbrm_upload_volume_async / tbrm_present_volume / tbrm_raymarch_lit_to_host_async gives, frame by frame, exactly what the
synchronous calls give."""
import numpy as np
import pytest

from tbraymarcherplugin_b200 import FMT_G8, synth
from tbraymarcherplugin_b200.raymarch_utils import FWindowingParameters, URaymarchUtils

pytestmark = pytest.mark.gpu
CT_WINDOW = FWindowingParameters(0.45, 0.5, True, False)


def make_res(dims):
    res = URaymarchUtils.InitializeRaymarchResources(dims, FMT_G8, bLightVolume32Bit=True)
    URaymarchUtils.ColorCurveToTexture(res, synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(res, CT_WINDOW)
    return res


def step(res, world):
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    for l in synth.LIGHTS[:2]:
        assert URaymarchUtils.AddDirLightToSingleVolume(res, l, True, world, bGPUSync=True)


def run_streaming(pin_volume, new_frame, dims=(64, 64, 64), view=(160, 96), steps=128.0):
    """pin_volume(ndarray) -> host buffer the upload reads, new_frame(shape) -> host buffer a frame lands in (pinned memory on a GPU)"""
    base = synth.perlin_ct_volume(dims)
    volumes = [np.ascontiguousarray(np.roll(base, 7 * k, axis=k % 3)) for k in range(5)]
    world = synth.identity_world()
    cam = synth.benchmark_camera(*view)
    # synchronous reference
    ref = make_res(dims)
    want = []
    for v in volumes:
        URaymarchUtils.SetDataVolume(ref, v)
        step(ref, world)
        want.append(URaymarchUtils.PerformWindowedLitRaymarch(ref, cam, world, steps)[0])
    assert not np.array_equal(want[0], want[1])
    # streaming: volume k+1 uploads while frame k is computed, frames download behind the next step
    res = make_res(dims)
    pinned = [pin_volume(v) for v in volumes]
    frames = [new_frame((cam.Height, cam.Width, 4)) for _ in volumes]
    URaymarchUtils.SetDataVolumeAsync(res, pinned[0])
    for k in range(len(volumes)):
        URaymarchUtils.PresentDataVolume(res)
        if k + 1 < len(volumes):
            URaymarchUtils.SetDataVolumeAsync(res, pinned[k + 1])
        step(res, world)
        URaymarchUtils.PerformWindowedLitRaymarchAsync(res, cam, world, steps, out=frames[k])
    URaymarchUtils.WaitForDownloads(res)
    URaymarchUtils.FlushRenderingCommands(res)
    for k in range(len(volumes)):
        assert np.array_equal(frames[k], want[k]), f"frame {k}"
    # presenting without a pending upload is an error, not a silent reuse
    with pytest.raises(Exception):
        URaymarchUtils.PresentDataVolume(res)
    ref.release(), res.release()


def test_streaming_frames_equal_synchronous_frames():
    import torch

    keep = []  # the tensors own the pinned memory their numpy views point into

    def pin(v):
        keep.append(torch.from_numpy(v).pin_memory())
        return keep[-1].numpy()

    def frame(shape):
        keep.append(torch.empty(shape, dtype=torch.float32).pin_memory())
        return keep[-1].numpy()

    run_streaming(pin, frame)
