import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_addoption(parser):
    parser.addoption("--emulate-kernels", action="store_true", default=False,
                     help="run against tests/emu (the CUDA sources compiled for the CPU, SIMT emulator) instead of libtbrm.so: lets the "
                          "-m gpu tests of the non-cooperative kernels execute on a machine without a GPU")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if config.getoption("--emulate-kernels"):
        import emu_lib
        from tbraymarcherplugin_b200 import _capi

        _capi._lib = emu_lib.load()


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle and the CUDA library exist (both are built in-tree and travel with the snapshot)."""
    if not (ROOT / "oracle" / "libtbrm_oracle.so").exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True)
    from tbraymarcherplugin_b200 import build

    if not build.LIB_PATH.exists():
        build.build()
    yield


def has_gpu() -> bool:
    from tbraymarcherplugin_b200 import _capi

    return _capi.load().tbrm_device_count() > 0
