"""hlsl2cpp.py — streams the reference's shader sources into C++ that g++ accepts (TEST INFRASTRUCTURE, build step of oracle/ref.mk).

    python oracle/hlsl2cpp.py /root/reference oracle/_ref/gen

Reads the .usf files of the hot path FROM WHERE THEY LIE under the reference checkout and writes one .inc file per shader into the
output directory (oracle/_ref/gen: a git-ignored build intermediate that is deleted again after compiling — nothing of the reference is
committed to this repository). oracle/ref_shaders.cpp includes each .inc inside its own namespace, on top of oracle/hlsl_shim.

The rewrites are purely syntactic; no expression, condition, constant or statement order of a shader is changed:
  * comments are dropped; `#pragma once` is dropped; `#include "/Engine/..."` is dropped (the shim supplies the environment);
    relative includes are inlined once per shader, like the HLSL preprocessor does;
  * `[numthreads(...)]` attributes and `: SV_DispatchThreadID` semantics are dropped;
  * unsuffixed floating-point literals get an `f` suffix (HLSL literals are float; C++'s would drag the arithmetic into double);
  * `inout T x` / `out T x` parameters become `T& x`, `in T x` becomes `T x`;
  * multi-component swizzles become member calls: `.xyz` -> `.xyz()`, and the one swizzled store `a.rgb = e;` -> `a.set_rgb(e);`;
  * `float v = ReadBuffer.SampleLevel(...);` (implicit float4 -> float truncation) gets an explicit `.x`;
  * `for (int i = 0; i < MaxSteps; i++)` whose `i` is read after the loop (legacy HLSL scoping, PerformWindowedRaymarchOctree)
    becomes `int i = 0; for (i = 0; ...)`.
"""
from __future__ import annotations

import re
import sys
from pathlib import Path

# (output name, file relative to <reference>/Source)
SHADERS = [
    ("AddDirLightShader", "Raymarcher/Shaders/Private/AddDirLightShader.usf"),
    ("ChangeDirLightShader", "Raymarcher/Shaders/Private/ChangeDirLightShader.usf"),
    ("WindowedRaymarchMaterials", "Raymarcher/Shaders/Private/WindowedRaymarchMaterials.usf"),
    ("GenerateOctreeShader", "Raymarcher/Shaders/Private/GenerateOctreeShader.usf"),
    ("SDFMarcher", "FractalMarcher/Shaders/Private/SDFMarcher.usf"),
    ("CalculateMandelbulbSDF", "FractalMarcher/Shaders/Private/CalculateMandelbulbSDF.usf"),
]

FLOAT_LITERAL = re.compile(r"(?<![\w.])(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")


def strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def inline_includes(path: Path, seen: set) -> str:
    out = []
    for line in strip_comments(path.read_text(encoding="utf-8", errors="replace")).splitlines():
        m = re.match(r'\s*#include\s+"([^"]+)"', line)
        if m:
            inc = m.group(1)
            if inc.startswith("/Engine/"):
                continue
            target = (path.parent / inc).resolve()
            if target not in seen:  # every file of the path carries `#pragma once`
                seen.add(target)
                out.append(inline_includes(target, seen))
            continue
        if re.match(r"\s*#pragma\s+once", line):
            continue
        out.append(line)
    return "\n".join(out)


def rewrite(text: str) -> str:
    text = re.sub(r"\[numthreads\([^\]]*\)\]", "", text)
    text = re.sub(r"\s*:\s*SV_DispatchThreadID", "", text)
    text = FLOAT_LITERAL.sub(r"\1f", text)
    text = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", text)
    text = re.sub(r"\bin\s+(?=(?:float|int|uint)\d?\b)", "", text)
    text = re.sub(r"(\w+)\.rgb\s*=\s*([^;]*);", r"\1.set_rgb(\2);", text)
    text = re.sub(r"\.(xyz|xy|rgb)\b(?!\s*\()", r".\1()", text)
    text = re.sub(r"(float\s+\w+\s*=\s*\w*ReadBuffer\.SampleLevel\([^;]*\));", r"\1.x;", text)
    text = text.replace("for (int i = 0; i < MaxSteps; i++)", "int i = 0; for (i = 0; i < MaxSteps; i++)")
    return re.sub(r"\n\s*\n+", "\n", text)


def main(ref_root: str, out_dir: str) -> None:
    src = Path(ref_root) / "Source"
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    for name, rel in SHADERS:
        path = (src / rel).resolve()
        body = rewrite(inline_includes(path, {path}))
        (out / f"{name}.inc").write_text(f"// generated from {rel} by oracle/hlsl2cpp.py — build intermediate, do not commit\n{body}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
