"""Host logic of the N > 1 path on CPU: world-size-2 gloo processes deal volumes / image rows, reduce timings with max,
and reassemble a frame exactly."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np

from tbraymarcherplugin_b200 import sharding

ROOT = Path(__file__).resolve().parents[1]


def test_every_volume_and_row_is_owned_exactly_once():
    for world in (1, 2, 3, 8):
        owned = sorted(v for r in range(world) for v in sharding.volumes_of_rank(11, r, world))
        assert owned == list(range(11))
        rows = sorted(b for r in range(world) for b in sharding.row_blocks_of_rank(1080, r, world, 8))
        assert rows[0][0] == 0 and rows[-1][1] == 1080 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    parts = [[np.full((e - b, 4), b) for b, e in sharding.row_blocks_of_rank(37, r, 2, 8)] for r in range(2)]
    img = sharding.assemble_rows(37, 2, parts, 8)
    assert img.shape == (37, 4) and list(img[:, 0]) == [8 * (i // 8) for i in range(37)]


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    from tbraymarcherplugin_b200 import sharding
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(os.environ["RANK"]), world_size=2)
    rank = dist.get_rank()
    # 1. timing reduction = max over ranks
    assert sharding.max_over_ranks(10.0 + rank) == 11.0
    # 2. each rank "renders" its row blocks of a synthetic frame; gathering + assembling reproduces the frame
    H, W = 45, 6
    frame = np.arange(H * W * 4, dtype=np.float32).reshape(H, W, 4)
    mine = [frame[b:e] for b, e in sharding.row_blocks_of_rank(H, rank, 2, 8)]
    gathered = [None, None]
    dist.all_gather_object(gathered, mine)
    assert np.array_equal(sharding.assemble_rows(H, 2, gathered, 8), frame)
    # 2b. the collective used by FShardedRaymarchVolume.Render: ranks hold their compacted interleaved rows (unequal shares),
    #     rank 0 receives the whole frame
    for H2 in (45, 64, 7):
        f2 = torch.arange(H2 * W * 4, dtype=torch.float32).reshape(H2, W, 4)
        local = f2[torch.as_tensor(sharding.rows_of_rank(H2, rank, 2, 8))]
        got = sharding.gather_interleaved_rows(local, H2, 8, dst=0)
        assert (got is None) == (rank != 0)
        if rank == 0:
            assert torch.equal(got, f2)
    # 2c. the Z-slab partition rule (library code, needs no GPU): slabs tile [0, Z) and are multiples of 8 slices
    for Z, n in ((512, 2), (512, 8), (1024, 8), (96, 3), (64, 4)):
        slabs = [sharding.slab_of_rank(Z, r, n) for r in range(n)]
        assert slabs[0][0] == 0 and slabs[-1][1] == Z and all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
        assert all((b - a) % 8 == 0 and b > a for a, b in slabs)
    # 3. volumes are partitioned, whole-job throughput = sum of per-rank units / max time
    vols = sharding.volumes_of_rank(5, rank, 2)
    n = torch.tensor([len(vols)], dtype=torch.float64)
    dist.all_reduce(n)
    assert int(n.item()) == 5
    dist.destroy_process_group()
    print("ok", rank)
""")


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT), port=29731))
    procs = [subprocess.Popen([sys.executable, str(script)], env={**os.environ, "RANK": str(r)}, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "ok 0" in outs[0] and "ok 1" in outs[1]
