"""TEST INFRASTRUCTURE: builds tests/emu/_build/libtbrm_emu.so — the product's CUDA sources (tbraymarcherplugin_b200/csrc) compiled by g++
against the SIMT emulator of tests/emu/cuda_on_host.h, so that the kernels' own source runs on a machine without a GPU.

The only edit the sources get is syntactic: `kernel<<<grid, block, smem, stream>>>(args)` becomes
`tbrm_emu::launch(grid, block, smem, stream, [&] { kernel(args); })`; the acquire / release / relaxed global accesses written as inline PTX become atomic accesses
(tbrm_emu::ld_global / st_global), and the cooperative launch of the generic fused sweep becomes tbrm_emu::launch_cooperative (all blocks
co-resident). The PTX helper block of the TMA-staged sweep (mbarriers, cp.async.bulk.tensor loads / stores, bulk-group waits) is replaced
by tests/emu/emu_tma_helpers.h — synchronous copies through an emulated tensor map with the device's bounds handling and
transaction-byte accounting — and its dynamic shared memory comes from the emulator. TBRM_EMU_TMA=0 / TBRM_EMU_COOPERATIVE=0 make the
emulated machine report no tensor-map driver entry point / no cooperative launch, so the fallbacks of those paths can be exercised too.

    python tests/emu/build_emu.py            (rebuilds only when a source is newer than the library)
"""
import re
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
CSRC = ROOT / "tbraymarcherplugin_b200" / "csrc"
BUILD = HERE / "_build"
LIB = BUILD / "libtbrm_emu.so"
UNITS = ["api.cu", "sweep.cu", "raymarch.cu", "mandelbulb.cu", "synth.cu", "ingest.cu"]
CXX = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
FLAGS = ["-O1", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-mavx2", "-mfma", "-w", "-pthread"]


def _match_back(text: str, end: int) -> int:
    """start of the kernel expression that ends at `end`: identifier (with namespaces) plus an optional template argument list"""
    i = end
    while i > 0 and text[i - 1].isspace():
        i -= 1
    if text[i - 1] == ">":
        depth = 0
        while True:
            i -= 1
            if text[i] == ">":
                depth += 1
            elif text[i] == "<":
                depth -= 1
                if depth == 0:
                    break
    while i > 0 and (text[i - 1].isalnum() or text[i - 1] in "_:"):
        i -= 1
    return i


def rewrite_launches(text: str) -> str:
    out, pos = [], 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            break
        start = _match_back(text, k)
        close = text.index(">>>", k)
        paren = text.index("(", close)
        assert text[close + 3:paren].strip() == "", text[k:paren + 1]
        depth, j = 0, paren
        while True:
            if text[j] == "(":
                depth += 1
            elif text[j] == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        kernel, cfg, args = text[start:k].strip(), text[k + 3:close], text[paren + 1:j]
        ncfg = len(re.findall(r",", re.sub(r"\([^()]*\)", "", cfg))) + 1
        cfg_full = cfg + ", 0" * (4 - ncfg)
        out.append(text[pos:start])
        out.append(f"tbrm_emu::launch({cfg_full}, [&] {{ {kernel}({args}); }})")
        pos = j + 1
    out.append(text[pos:])
    return "".join(out)


_PTX = [
    # relaxed / acquire loads of a flag or ring cell: an atomic load that also lets the other fibers run (every spin loop goes through one)
    (re.compile(r'asm volatile\("ld\.(?:relaxed|acquire)\.(?:gpu|sys)\.global\.u(?:32|64) %0, \[%1\];"\s*:\s*"=[rl]"\((\w+)\)\s*:\s*"l"\((\w+)\)\s*:\s*"memory"\);'),
     r'\1 = tbrm_emu::ld_global(\2);'),
    (re.compile(r'asm volatile\("st\.(?:relaxed|release)\.(?:gpu|sys)\.global\.u(?:32|64) \[%0\], %1;"\s*::\s*"l"\((\w+)\),\s*"[rl]"\((\w+)\)\s*:\s*"memory"\);'),
     r'tbrm_emu::st_global(\1, \2);'),
    (re.compile(r'asm volatile\("mov\.u64 %0, %globaltimer;"\s*:\s*"=l"\((\w+)\)\);'), r'\1 = tbrm_emu::globaltimer_ns();'),
]


def rewrite_ptx(text: str) -> str:
    for rx, repl in _PTX:
        text = rx.sub(repl, text)
    # type-erased cooperative launch -> typed (the kernel variable keeps its function-pointer type)
    return text.replace("cudaLaunchCooperativeKernel((const void*) kernel,", "tbrm_emu::launch_cooperative(kernel,")


def must_replace(text: str, old: str, new: str) -> str:
    assert text.count(old) == 1, f"expected exactly one occurrence of {old!r}"
    return text.replace(old, new)


TMA_KERNEL_TYPE = "void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const PushMaps, const TmaParams, const float4*)"


def splice_tma_helpers(text: str) -> str:
    """sweep_tma.cuh: the PTX helper block (mbarrier / TMA / bulk-group waits) -> tests/emu/emu_tma_helpers.h; the type-erased cooperative
    launch of the sweep_tma_kernel instantiations (one signature) -> typed."""
    a = text.index("// ---- PTX helpers")
    b = text.index('}  // namespace tbrm', a)
    text = text[:a] + '}  // namespace tbrm\n#include "emu_tma_helpers.h"\nnamespace tbrm {\n' + text[b:]
    return must_replace(text, "return cudaLaunchCooperativeKernel(kern,", f"return tbrm_emu::launch_cooperative(({TMA_KERNEL_TYPE}) kern,")


def generate() -> Path:
    gen = BUILD / "gen" / "pkg" / "csrc"
    gen.mkdir(parents=True, exist_ok=True)
    (BUILD / "gen" / "include").mkdir(exist_ok=True)
    shutil.copy(ROOT / "include" / "tbrm.h", BUILD / "gen" / "include" / "tbrm.h")
    for src in sorted(CSRC.iterdir()):
        if src.suffix not in (".cu", ".cuh", ".hpp", ".h"):
            continue
        text = rewrite_ptx(rewrite_launches(src.read_text()))
        if src.name == "sweep_tma.cuh":
            text = splice_tma_helpers(text)
        if src.name in ("sweep_tma_kernel.cuh", "sweep_split.cuh"):
            text = must_replace(text, "extern __shared__ __align__(128) unsigned char smem[];", "unsigned char* smem = tbrm_emu::dynamic_smem();")
            text = must_replace(text, 'asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");', "")
        assert "asm volatile" not in text, f"{src.name}: inline PTX the emulator build does not know:\n" + text[text.index("asm volatile"):][:300]
        (gen / (src.name + ".cpp" if src.suffix == ".cu" else src.name)).write_text(text)
    return gen


def newest_source() -> float:
    files = [p for p in CSRC.iterdir()] + [p for p in HERE.iterdir() if p.is_file()] + [ROOT / "include" / "tbrm.h"]
    return max(p.stat().st_mtime for p in files)


def build(force: bool = False) -> Path:
    if LIB.exists() and not force and LIB.stat().st_mtime > newest_source():
        return LIB
    gen = generate()
    inc = ["-I", str(HERE / "include"), "-I", str(HERE), "-I", str(gen), "-include", str(HERE / "cuda_on_host.h")]
    objs, procs = [], []
    for unit in UNITS + ["simt_engine"]:
        src = HERE / "simt_engine.cpp" if unit == "simt_engine" else gen / (unit + ".cpp")
        obj = BUILD / (unit + ".o")
        objs.append(str(obj))
        procs.append((unit, subprocess.Popen([CXX, *FLAGS, *inc, "-c", str(src), "-o", str(obj)], stderr=subprocess.PIPE, text=True)))
    failed = False
    for unit, p in procs:
        _, err = p.communicate()
        if p.returncode:
            failed = True
            sys.stderr.write(f"---- {unit}\n{err[:6000]}\n")
    if failed:
        raise RuntimeError("emulator build failed")
    subprocess.run([CXX, "-shared", "-pthread", "-o", str(LIB), *objs, "-lz"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
