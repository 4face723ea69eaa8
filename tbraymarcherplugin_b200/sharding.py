"""Host-side sharding logic for multi-GPU runs (one process per GPU, launched by torchrun) — SURVEY.md §8(e), DESIGN.md §8.

Two ways to spread the hot path over the GPUs of a box:

* **independent volumes** (a scene holds several ARaymarchVolumes, each with its own resources): volumes are dealt to
  ranks, no data-path collective (``volumes_of_rank``);
* **one volume, Z-slab sharded** (``FShardedRaymarchVolume``): every rank holds the whole R8 data volume (1 B/voxel,
  replicated by an NCCL all-gather of the uploaded slabs) and a full-size light volume of which it owns the slices
  ``tbrm_slab_partition`` assigns to it. ``AddDirLight`` runs on every rank; the propagated light crosses slab boundaries
  *inside the sweep kernel* through NVLink peer stores into the neighbours' exchange arenas (libtbrm.so, CUDA IPC), so a
  sharded sweep is bit-identical to an unsharded one. ``GatherLightVolume`` is an in-place NCCL all-gather of the slabs;
  ``Render`` deals the frame to the ranks in interleaved 8-row blocks and gathers them on rank 0.

torch / torch.distributed are plumbing here: device memory the collectives can address, the NCCL calls and their stream.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _capi


# --------------------------------------------------------------------------------------------------------------
# pure host logic (no GPU needed; covered by the gloo tests)
# --------------------------------------------------------------------------------------------------------------
def volumes_of_rank(n_volumes: int, rank: int, world_size: int) -> List[int]:
    """Volume indices owned by `rank`: round-robin, every volume owned exactly once."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_volumes, world_size))


def row_blocks_of_rank(height: int, rank: int, world_size: int, block_rows: int = 8) -> List[Tuple[int, int]]:
    """Interleaved [begin, end) row blocks of an image of `height` rows owned by `rank`."""
    if block_rows <= 0:
        raise ValueError("block_rows must be positive")
    blocks = [(b, min(b + block_rows, height)) for b in range(0, height, block_rows)]
    return blocks[rank::world_size]


def rows_of_rank(height: int, rank: int, world_size: int, block_rows: int = 8) -> np.ndarray:
    """Image rows rank `rank` renders, in the order they appear in its compacted output (tbrm_raymarch_lit_interleaved)."""
    blocks = row_blocks_of_rank(height, rank, world_size, block_rows)
    if not blocks:
        return np.zeros(0, np.int64)
    return np.concatenate([np.arange(b, e, dtype=np.int64) for b, e in blocks])


def assemble_rows(height: int, world_size: int, per_rank_rows: Sequence[Sequence], block_rows: int = 8):
    """Inverse of row_blocks_of_rank: per_rank_rows[r] is the list of row-block arrays rank r rendered, in order."""
    out = [None] * len(range(0, height, block_rows))
    for r in range(world_size):
        for i, blk in enumerate(per_rank_rows[r]):
            out[r + i * world_size] = blk
    return np.concatenate(out, axis=0)


def slab_of_rank(z_slices: int, rank: int, world_size: int) -> Tuple[int, int]:
    """The Z-slab [z_begin, z_end) of `rank` — the partition rule of the library (multiples of 8 slices)."""
    z0, z1 = C.c_int32(), C.c_int32()
    _capi.load().tbrm_slab_partition(int(z_slices), int(world_size), int(rank), C.byref(z0), C.byref(z1))
    return z0.value, z1.value


def max_over_ranks(value: float, device=None) -> float:
    """Timing reduction used by bench.py: every multi-GPU number is the max over ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


_PERM_CACHE: dict = {}  # (height, world, block_rows, device) -> source row of every frame row in the all-gathered buffer


def gather_interleaved_rows(local_rows, height: int, block_rows: int, dst: int = 0, group=None):
    """Gather the compacted row blocks every rank rendered into the full frame on rank `dst` (None elsewhere).

    `local_rows`: tensor (rows_of_rank, W, C) on the backend's device (CUDA for nccl, CPU for gloo). Ranks own different
    numbers of rows, so blocks are padded to the largest share for the fixed-size gather and scattered by row index."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [len(rows_of_rank(height, r, world, block_rows)) for r in range(world)]
    if local_rows.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local_rows.shape[0]} rows, expected {counts[rank]}")
    pad = max(counts)
    send = local_rows
    if send.shape[0] < pad:
        send = torch.cat([send, send.new_zeros((pad - send.shape[0],) + tuple(send.shape[1:]))], 0)
    send = send.contiguous()
    if send.is_cuda:
        # one collective + one gather kernel (the per-rank gather / index_copy loop cost 0.43 ms at 8 GPUs for a 33 MB frame: launch latency).
        # Every rank receives every share (31 MB over NVSwitch: ~0.05 ms); only `dst` assembles the frame.
        allbuf = send.new_empty((world * pad,) + tuple(send.shape[1:]))
        dist.all_gather_into_tensor(allbuf, send, group=group)
        if rank != dst:
            return None
        key = (height, world, block_rows, str(send.device))
        perm = _PERM_CACHE.get(key)
        if perm is None:
            src_of_row = [0] * height
            for r in range(world):
                for j, row in enumerate(rows_of_rank(height, r, world, block_rows)):
                    src_of_row[row] = r * pad + j
            perm = torch.as_tensor(src_of_row, device=send.device)
            _PERM_CACHE[key] = perm
        return allbuf.index_select(0, perm)
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = send.new_empty((height,) + tuple(send.shape[1:]))
    for r in range(world):
        idx = torch.as_tensor(rows_of_rank(height, r, world, block_rows), device=send.device)
        out.index_copy_(0, idx, bufs[r][: counts[r]])
    return out


# --------------------------------------------------------------------------------------------------------------
# one volume sharded over the ranks of a process group
# --------------------------------------------------------------------------------------------------------------
class FShardedRaymarchVolume:
    """FBasicRaymarchRenderingResources of ONE ARaymarchVolume, Z-slab sharded over a torch.distributed group.

    Every method is collective: all ranks call it with the same arguments (like the render commands of one volume)."""

    BLOCK_ROWS = 8

    def __init__(self, data_dims: Sequence[int], device: int, group=None, bLightVolume32Bit: bool = True, LightVolumeHalfResolution: bool = False):
        import torch
        import torch.distributed as dist

        from .raymarch_utils import FMT_G8, URaymarchUtils

        self.lib = _capi.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device
        X, Y, Z = (int(d) for d in data_dims)
        self.dims = (X, Y, Z)
        # data slabs (upload + all-gather) and light slabs (sweep + all-gather) follow the same partition rule on their own slice counts
        self.z0, self.z1 = slab_of_rank(Z, self.rank, self.world)
        slabs = [slab_of_rank(Z, r, self.world) for r in range(self.world)]
        if any(b - a != self.z1 - self.z0 or b <= a for a, b in slabs):
            raise ValueError(f"{Z} slices do not split into {self.world} equal slabs of a multiple of 8 slices")
        self.res = URaymarchUtils.InitializeRaymarchResources(self.dims, FMT_G8, bLightVolume32Bit=bLightVolume32Bit,
                                                              LightVolumeHalfResolution=LightVolumeHalfResolution, device=device)
        self.light32 = bool(bLightVolume32Bit)
        LX, LY, LZ = (int(d) for d in self.res.LightDims)
        self.lz0, self.lz1 = slab_of_rank(LZ, self.rank, self.world)
        lslabs = [slab_of_rank(LZ, r, self.world) for r in range(self.world)]
        if any(b - a != self.lz1 - self.lz0 or b <= a for a, b in lslabs):
            raise ValueError(f"{LZ} light-volume slices do not split into {self.world} equal slabs of a multiple of 8 slices")
        dev = torch.device("cuda", device)
        # collectives run on these tensors, the library computes on them: caller-owned, bound into the resource set
        self.data = torch.empty((Z, Y, X), dtype=torch.uint8, device=dev)
        # the light volume stays the library's own allocation (a CUDA IPC handle names a whole allocation: the peers map it for the
        # push-gather); torch sees it through the CUDA array interface, for the NCCL all-gather and for reads
        self.lib.tbrm_light_volume_device_ptr.restype = C.c_void_p
        light_ptr = self.lib.tbrm_light_volume_device_ptr(self.res.handle)

        class _LightView:
            __cuda_array_interface__ = {"shape": (LZ, LY, LX), "typestr": "<f4" if bLightVolume32Bit else "|u1", "data": (int(light_ptr), False),
                                        "version": 2}

        self._light_view = _LightView()
        with torch.cuda.device(dev):
            self.light = torch.as_tensor(self._light_view, device=dev)
            self.light.zero_()
        torch.cuda.synchronize(dev)
        _capi.check(self.lib.tbrm_bind_volume_device(self.res.handle, C.c_void_p(self.data.data_ptr())))
        self.res.bIsInitialized = True
        slab = _capi.Slab(self.rank, self.world, self.lz0, self.lz1)
        _capi.check(self.lib.tbrm_slab_configure(self.res.handle, C.byref(slab)))
        # exchange arenas: CUDA IPC handles travel through the process group, neighbours map each other's arena
        handle = (C.c_ubyte * 64)()
        _capi.check(self.lib.tbrm_slab_ipc_handle(self.res.handle, handle))
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        for side, peer in ((-1, self.rank - 1), (+1, self.rank + 1)):
            if 0 <= peer < self.world:
                buf = (C.c_ubyte * 64).from_buffer_copy(handles[peer])
                _capi.check(self.lib.tbrm_slab_open_peer(self.res.handle, side, buf))
        # push-gather: every rank maps every other rank's light volume
        lhandle = (C.c_ubyte * 64)()
        _capi.check(self.lib.tbrm_slab_light_ipc_handle(self.res.handle, lhandle))
        lhandles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(lhandles, bytes(lhandle), group=group)
        for peer in range(self.world):
            if peer != self.rank:
                buf = (C.c_ubyte * 64).from_buffer_copy(lhandles[peer])
                _capi.check(self.lib.tbrm_slab_open_peer_light(self.res.handle, peer, buf))
        self._pushed = False  # the last sweep call pushed its light bricks to the peers: no all-gather needed
        self._sync_word = torch.zeros(1, dtype=torch.float32, device=dev)
        dist.barrier(group=group)
        # the library enqueues on a torch-owned stream, and the NCCL ops are issued under the same stream, so that
        # sweep -> all-gather -> raymarch -> gather stay stream-ordered without host synchronisation
        self.stream = torch.cuda.Stream(device=dev)
        _capi.check(self.lib.tbrm_set_stream(self.res.handle, C.c_void_p(self.stream.cuda_stream)))
        self._frame = None

    # ---- inputs -------------------------------------------------------------------------------------------
    def SetDataVolumeSlab(self, slab) -> None:
        """Upload this rank's slab of the data volume ([z0:z1] of the (Z,Y,X) array; numpy, or a torch tensor on any device)
        and replicate the volume on every GPU with an all-gather over NVLink."""
        import torch
        import torch.distributed as dist

        t = torch.as_tensor(slab) if not isinstance(slab, torch.Tensor) else slab
        if tuple(t.shape) != (self.z1 - self.z0, self.dims[1], self.dims[0]) or t.dtype != torch.uint8:
            raise ValueError("slab must be uint8 with shape (z1 - z0, Y, X)")
        with torch.cuda.stream(self.stream):
            self.data[self.z0:self.z1].copy_(t, non_blocking=True)
            dist.all_gather_into_tensor(self.data.view(-1), self.data[self.z0:self.z1].reshape(-1), group=self.group)
        # the replica / brick grid derived from the data volume are rebuilt lazily
        _capi.check(self.lib.tbrm_bind_volume_device(self.res.handle, C.c_void_p(self.data.data_ptr())))

    # ---- streaming inputs / outputs: copies and the data all-gather overlap the previous / next step's kernels ------------------
    def _ensure_streaming(self) -> None:
        import torch
        import torch.distributed as dist

        if getattr(self, "_upload_stream", None) is not None:
            return
        dev = self.light.device
        self._upload_stream = torch.cuda.Stream(device=dev)
        self._download_stream = torch.cuda.Stream(device=dev)
        self._data_back = torch.empty_like(self.data)
        self._ev_uploaded = torch.cuda.Event()
        self._ev_back_free = torch.cuda.Event()
        self._ev_back_free.record(self.stream)
        self._upload_pending = False
        # its own communicator: the all-gather of the NEXT step's data runs on the upload stream while the render queue's collectives
        # (light all-gather, frame gather) run on the library's stream
        self._upload_group = dist.new_group(ranks=list(range(self.world))) if self.group is None else self.group
        self._downloads = []

    def SetDataVolumeSlabAsync(self, slab) -> None:
        """Start uploading this rank's slab of the NEXT data volume (pinned host tensor for a true overlap) into the back buffer and
        replicating it with an all-gather, on the upload stream. PresentDataVolume makes it current."""
        import torch
        import torch.distributed as dist

        self._ensure_streaming()
        t = torch.as_tensor(slab) if not isinstance(slab, torch.Tensor) else slab
        if tuple(t.shape) != (self.z1 - self.z0, self.dims[1], self.dims[0]) or t.dtype != torch.uint8:
            raise ValueError("slab must be uint8 with shape (z1 - z0, Y, X)")
        if self._upload_pending:
            raise RuntimeError("an uploaded volume is waiting for PresentDataVolume")
        with torch.cuda.stream(self._upload_stream):
            self._upload_stream.wait_event(self._ev_back_free)  # the render queue no longer reads the back buffer
            self._data_back[self.z0:self.z1].copy_(t, non_blocking=True)
            dist.all_gather_into_tensor(self._data_back.view(-1), self._data_back[self.z0:self.z1].reshape(-1), group=self._upload_group)
            self._ev_uploaded.record(self._upload_stream)
        self._upload_pending = True

    def PresentDataVolume(self) -> None:
        """Make the uploaded volume current: the render queue waits for the upload, the buffers swap."""
        if not getattr(self, "_upload_pending", False):
            raise RuntimeError("no uploaded volume is pending")
        self.stream.wait_event(self._ev_uploaded)
        self.data, self._data_back = self._data_back, self.data
        self._ev_back_free.record(self.stream)  # everything enqueued so far may still read the old buffer; what follows reads the new one
        _capi.check(self.lib.tbrm_bind_volume_device(self.res.handle, C.c_void_p(self.data.data_ptr())))
        self._upload_pending = False

    def RenderToHostAsync(self, cam, world, step_count: float, out) -> None:
        """Render (collective); rank 0 copies the assembled frame into `out` (pinned (H, W, 4) float32 tensor) on the download stream.
        Read `out` after WaitForDownloads."""
        import torch

        self._ensure_streaming()
        frame, _ = self.Render(cam, world, step_count, gather=True, count_steps=False)
        if self.rank == 0:
            ev = torch.cuda.Event()
            ev.record(self.stream)
            with torch.cuda.stream(self._download_stream):
                self._download_stream.wait_event(ev)
                out.copy_(frame, non_blocking=True)
                frame.record_stream(self._download_stream)
                done = torch.cuda.Event()
                done.record(self._download_stream)
            self._downloads.append(done)

    def WaitForDownloads(self) -> None:
        for ev in getattr(self, "_downloads", []):
            ev.synchronize()
        self._downloads = []

    # ---- sweep (collective; same arguments on every rank) ---------------------------------------------------
    def ClearLightVolume(self, value: float = 0.0) -> None:
        from .raymarch_utils import URaymarchUtils

        self._pushed = False
        URaymarchUtils.ClearResourceLightVolumes(self.res, value)

    def AddDirLight(self, light, added: bool, world, stats=None, push: bool = False) -> bool:
        """`push=True` (collective: the same on every rank) for the LAST light of a reset: its last axis pass stores every finished light brick
        into all peers' light volumes from inside the sweep kernel (TMA stores over NVLink), and GatherLightVolume only synchronises."""
        from .raymarch_utils import URaymarchUtils

        push = bool(push and self.light32)  # the pushed bricks are float boxes: a G8 volume is gathered by NCCL
        _capi.check(self.lib.tbrm_slab_push_light(self.res.handle, 1 if push else 0))
        try:
            ok = URaymarchUtils.AddDirLightToSingleVolume(self.res, light, added, world, bGPUSync=True, stats=stats)
        finally:
            _capi.check(self.lib.tbrm_slab_push_light(self.res.handle, 0))
        self._pushed = bool(push and ok)
        return ok

    def ChangeDirLight(self, old_light, new_light, world, stats=None) -> bool:
        from .raymarch_utils import URaymarchUtils

        self._pushed = False  # this rank's slab changes again: the peers' copies of it are stale until the next gather
        return URaymarchUtils.ChangeDirLightInSingleVolume(self.res, old_light, new_light, world, bGPUSync=True, stats=stats)

    def GatherLightVolume(self) -> None:
        """In-place all-gather of the slabs: afterwards every rank holds the whole light volume (the raymarch reads it)."""
        import torch
        import torch.distributed as dist

        with torch.cuda.stream(self.stream):
            if self._pushed:
                # the bricks already sit in every rank's volume; what remains is to know that every rank's sweep (and with it its
                # stores) has completed before anybody reads: a one-word all-reduce, stream-ordered after the sweep on every rank
                dist.all_reduce(self._sync_word, group=self.group)
            else:
                dist.all_gather_into_tensor(self.light.view(-1), self.light[self.lz0:self.lz1].reshape(-1), group=self.group)
        self._pushed = False

    # ---- frame ------------------------------------------------------------------------------------------
    def local_rows(self, height: int) -> int:
        return int(self.lib.tbrm_raymarch_interleaved_rows(height, self.BLOCK_ROWS, self.rank, self.world))

    def Render(self, cam, world, step_count: float, gather: bool = True, count_steps: bool = False):
        """Lit raymarch of this rank's interleaved row blocks; the frame is assembled on rank 0 (returned there as a
        (H, W, 4) CUDA tensor, None elsewhere). Returns (frame, executed march steps of this rank or 0)."""
        import torch

        rows = self.local_rows(cam.Height)
        if self._frame is None or self._frame.shape != (rows, cam.Width, 4):
            self._frame = torch.empty((rows, cam.Width, 4), dtype=torch.float32, device=self.light.device)
        c, w = cam.to_c(), world.to_c()
        steps = C.c_uint64(0)
        _capi.check(self.lib.tbrm_raymarch_lit_interleaved(self.res.handle, C.byref(c), C.byref(w), float(step_count), self.BLOCK_ROWS,
                                                           self.rank, self.world, C.c_void_p(self._frame.data_ptr()), 1,
                                                           C.byref(steps) if count_steps else None))
        frame = None
        if gather:
            with torch.cuda.stream(self.stream):
                frame = gather_interleaved_rows(self._frame, cam.Height, self.BLOCK_ROWS, dst=0, group=self.group)
        return frame, int(steps.value)

    # ---- queue control --------------------------------------------------------------------------------------
    def Flush(self) -> None:
        _capi.check(self.lib.tbrm_flush(self.res.handle))

    def Check(self) -> None:
        """Raises if a slab exchange timed out since the last call (synchronises)."""
        _capi.check(self.lib.tbrm_slab_check(self.res.handle))

    def release(self) -> None:
        self.Flush()
        self.res.release()
