"""TEST INFRASTRUCTURE: the product's CUDA sources compiled for the CPU against the SIMT emulator of tests/emu (see tests/emu/cuda_on_host.h),
loaded with the prototypes of tbraymarcherplugin_b200/_capi.py. `use()` points the Python operator surface of THIS process at it (tests only;
the product never looks for it)."""
import ctypes as C
import importlib.util
from pathlib import Path

from tbraymarcherplugin_b200 import _capi

_HERE = Path(__file__).resolve().parent
_emu = None


def load() -> C.CDLL:
    global _emu
    if _emu is None:
        spec = importlib.util.spec_from_file_location("build_emu", _HERE / "emu" / "build_emu.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        lib = C.CDLL(str(mod.build()))
        for name, (res, args) in _capi.PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _emu = lib
    return _emu


def use(monkeypatch) -> C.CDLL:
    lib = load()
    monkeypatch.setattr(_capi, "_lib", lib)
    return lib
