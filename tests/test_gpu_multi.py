"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): one volume Z-slab sharded over 2 (and 4, 8 when present)
processes, one per GPU, launched with torchrun. The checks live in tests/mp_sharded_worker.py."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def gpu_count() -> int:
    from tbraymarcherplugin_b200 import _capi

    return _capi.load().tbrm_device_count()


@pytest.mark.parametrize("nproc,n", [(2, 128), (2, 256), (4, 256), (8, 256)])
def test_sharded_volume_matches_single_gpu(nproc, n):
    if gpu_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + nproc), str(ROOT / "tests" / "mp_sharded_worker.py"), str(n)]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env={**os.environ})
    assert p.returncode == 0, p.stdout[-4000:]
    assert p.stdout.count("sharded ok") == nproc, p.stdout[-4000:]
