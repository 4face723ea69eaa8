"""Writes tests/golden/ref_*.npz from the REFERENCE'S OWN host code (oracle/_ref/libtbrm_ref.so = LightingShaderUtils.cpp,
VolumeInfo.cpp and the TextureUtilities.h templates of /root/reference, compiled against the engine-type shim of
oracle/ue_shim by oracle/ref.mk). Unlike make_golden.py these vectors ARE reference output: they pin the oracle (and the
product's host math) to the reference for SURVEY.md §8 rows a15-a21 and the volume normalisation of row (f)3, and they
travel to machines where /root/reference does not exist.

    python tests/golden/make_golden_ref.py        (needs /root/reference or a prebuilt oracle/_ref/libtbrm_ref.so)
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE.parent))

import refpin  # noqa: E402
from tbraymarcherplugin_b200 import synth  # noqa: E402
from tbraymarcherplugin_b200.raymarch_utils import (FClippingPlaneParameters, FDirLightParameters, FRaymarchWorldParameters,  # noqa: E402
                                                    FTransform)

SPECIAL_DIRS = [(0, 0, -1), (1, 0, 0), (0, -1, 0), (1, 1, 0), (1, -1, 0), (-1, 1, 1), (1, 1, 1), (0, 1, -1), (-1, 0, -1), (1, 1e-4, 0),
                (0, 0, 0), (0.05, 0.02, -1), (-3, 0.4, 0.3)]


def world_to_row(w: FRaymarchWorldParameters) -> np.ndarray:
    c = w.to_c()
    return np.array([*c.translation, *c.rotation, *c.scale, *c.clip.center, *c.clip.direction], np.float64)


def world_from_row(r: np.ndarray) -> FRaymarchWorldParameters:
    t = FTransform(tuple(r[0:3]), tuple(r[3:7]), tuple(r[7:10]))
    return FRaymarchWorldParameters(t, FClippingPlaneParameters(tuple(r[10:13]), tuple(r[13:16])))


def hostmath_inputs():
    rng = np.random.default_rng(20261017)
    dims, dirs, inten, worlds = [], [], [], []
    fixed = [synth.identity_world(), synth.scaled_rotated_world(), synth.clipped_world()]
    for i, d in enumerate(SPECIAL_DIRS):
        for w in fixed[: 1 + (i % 3)]:
            dims.append((32, 48, 64) if i % 2 else (512, 512, 512)), dirs.append(d), inten.append(0.7), worlds.append(world_to_row(w))
    for l in synth.LIGHTS:
        for w in fixed:
            dims.append((512, 512, 512)), dirs.append(tuple(l.LightDirection)), inten.append(l.LightIntensity), worlds.append(world_to_row(w))
    for i in range(300):
        axis = rng.standard_normal(3)
        t = FTransform.from_axis_angle(tuple(axis), float(rng.uniform(-180, 180)), tuple(rng.uniform(-50, 50, 3)), tuple(rng.uniform(0.3, 3.0, 3)))
        w = FRaymarchWorldParameters(t, FClippingPlaneParameters(tuple(rng.uniform(-40, 40, 3)), tuple(rng.standard_normal(3))))
        dims.append(tuple(int(x) for x in rng.integers(5, 1100, 3))), dirs.append(tuple(rng.standard_normal(3) * 3))
        inten.append(float(rng.uniform(0.05, 1.5))), worlds.append(world_to_row(w if i % 5 else fixed[0]))
    return np.array(dims, np.int32), np.array(dirs, np.float64), np.array(inten, np.float32), np.array(worlds, np.float64)


def hostmath_case():
    dims, dirs, inten, worlds = hostmath_inputs()
    plans = []
    for d, l, i, w in zip(dims, dirs, inten, worlds):
        p = refpin.plan_dir_light(tuple(int(x) for x in d), FDirLightParameters(tuple(l), float(i)), world_from_row(w))
        plans.append(refpin.plan_to_vector(p))
    perms = np.stack([refpin.permutation_rows(f) for f in range(6)])
    return {"dims": dims, "dirs": dirs, "intensity": inten, "worlds": worlds, "plans": np.stack(plans), "permutation_rows": perms}


SHADER_DIMS = (24, 20, 16)


def shader_inputs():
    import oracle
    from tbraymarcherplugin_b200.raymarch_utils import FWindowingParameters
    data = synth.perlin_ct_volume(SHADER_DIMS)
    return data, oracle.prepare_tf(synth.soft_ct_curve()), FWindowingParameters(0.45, 0.5, True, False)


def shader_sequence(vol, world, out, tag):
    """Full reset with the four benchmark lights, removal of L2, an in-place ChangeDirLight (same major axes), one that falls back to
    Remove + Add (different axes), then a lit frame — the states tests/test_ref_shaders_cpu.py replays through the oracle."""
    for l in synth.LIGHTS:
        vol.add_dir_light(l, True, world)
    out[f"{tag}_reset"] = vol.light.copy()
    vol.add_dir_light(synth.LIGHTS[1], False, world)
    out[f"{tag}_removed"] = vol.light.copy()
    vol.change_dir_light(synth.LIGHTS[0], synth.rotate_about_z(synth.LIGHTS[0], 5.0), world)
    out[f"{tag}_changed"] = vol.light.copy()
    vol.change_dir_light(synth.LIGHTS[2], synth.rotate_about_z(synth.LIGHTS[2], 80.0), world)
    out[f"{tag}_changed_fallback"] = vol.light.copy()


SHADER_WORLDS = {"identity": synth.identity_world, "scaled_rotated": synth.scaled_rotated_world, "clipped": synth.clipped_world}


def shaders_case():
    """Outputs of the reference's OWN shaders (AddDirLightShader.usf, ChangeDirLightShader.usf, WindowedRaymarchMaterials.usf, compiled
    for the CPU by oracle/ref.mk) driven by the reference's own host math."""
    data, tf, win = shader_inputs()
    out = {}
    for name, mk in SHADER_WORLDS.items():
        for light32 in (True, False):
            vol = refpin.RefVolume(data, tf, win, light32=light32)
            tag = f"{name}_{'r32f' if light32 else 'g8'}"
            shader_sequence(vol, mk(), out, tag)
            if light32:
                cam = synth.benchmark_camera(40, 24, jitter=True, frame=3)
                out[f"{tag}_setup"] = vol.raymarch(-1, cam, mk(), 48.0)
                out[f"{tag}_lit"] = vol.raymarch(0, cam, mk(), 48.0)
    return out


MATERIAL_DIMS = [(20, 9, 5), (32, 40, 24)]
MATERIAL_WINDOWS = {"ct": (0.45, 0.5, True, False), "full": (0.5, 1.0, True, True)}
MATERIAL_WORLDS = {"identity": synth.identity_world, "clipped": synth.clipped_world}
MANDELBULB_VIEW = (32, 24)
SDF_CASE = dict(dims=(20, 16, 12), center=(0.1, 0.0, -0.05), extent=2.4, power=8.0)


def material_camera():
    return synth.benchmark_camera(40, 24, jitter=True, frame=2)


def mandelbulb_params():
    from tbraymarcherplugin_b200.raymarch_utils import FMandelbulbParameters
    return FMandelbulbParameters(MaxSteps=64.0, MaxIterations=8.0)


def materials_case():
    """The reference's GenerateOctreeShader.usf, PerformWindowedIntensityRaymarch, PerformWindowedRaymarchOctree, SDFMarcher.usf (distance and
    normal variants) and CalculateMandelbulbSDF.usf, compiled for the CPU."""
    import oracle
    from tbraymarcherplugin_b200.raymarch_utils import FWindowingParameters
    tf = oracle.prepare_tf(synth.soft_ct_curve())
    out = {}
    for dims in MATERIAL_DIMS:
        tag = "x".join(map(str, dims))
        data = synth.perlin_ct_volume(dims)
        mips = refpin.generate_octree(data)
        for m, a in enumerate(mips):
            out[f"octree_{tag}_mip{m}"] = a
        for wname, wv in MATERIAL_WINDOWS.items():
            for world_name, mkw in MATERIAL_WORLDS.items():
                vol = refpin.RefVolume(data, tf, FWindowingParameters(*wv))
                out[f"intensity_{tag}_{wname}_{world_name}"] = vol.raymarch(1, material_camera(), mkw(), 40.0)
                for mip in (0, 2):
                    out[f"octree_march_{tag}_{wname}_{world_name}_mip{mip}"] = vol.raymarch(2, material_camera(), mkw(), 40.0, octree=mips, octree_mip=mip)
    cam = synth.benchmark_camera(*MANDELBULB_VIEW, jitter=False)
    out["mandelbulb_distance"] = refpin.mandelbulb_march(0, mandelbulb_params(), cam, synth.identity_world())
    out["mandelbulb_normal"] = refpin.mandelbulb_march(1, mandelbulb_params(), cam, synth.identity_world(), 0.01)
    out["mandelbulb_sdf_g16"] = refpin.mandelbulb_sdf(g16=True, **SDF_CASE)
    out["mandelbulb_sdf_r32f"] = refpin.mandelbulb_sdf(g16=False, **SDF_CASE)
    return out


def ingest_inputs():
    """One array per EVolumeVoxelFormat: random values plus the type's extremes for the narrow types."""
    rng = np.random.default_rng(77)
    arrays = {}
    for fmt, dt in refpin.VOXEL_DTYPES.items():
        if fmt == 6:
            a = (rng.standard_normal(600) * 1500.0).astype(np.float32)
        else:
            info = np.iinfo(dt)
            a = rng.integers(max(info.min, -40000), min(info.max, 90000), 600, endpoint=True).astype(dt)
            if np.dtype(dt).itemsize == 1:
                a[:2] = (info.min, info.max)
        arrays[fmt] = a
    arrays[16] = (-np.abs(rng.standard_normal(300)) - 1.0).astype(np.float32)  # all-negative floats: InMax stays FLT_MIN (TextureUtilities.h:111)
    return arrays


def ingest_case():
    """UVolumeTextureToolkit::NormalizeArrayByFormat / ConvertArrayToFloat and FVolumeInfo::Normalize* of the reference."""
    out = {}
    for key, a in ingest_inputs().items():
        fmt = key % 10
        n, lo, hi = refpin.normalize_array(fmt, a)
        out[f"in_{key}"], out[f"normalized_{key}"], out[f"minmax_{key}"] = a, n, np.array([lo, hi], np.float32)
        if fmt != 6:
            out[f"float_{key}"] = refpin.convert_to_float(fmt, a)
    vals = np.array([-1000.0, 0.0, 37.5, 3000.0, 1e-3], np.float32)
    out["info_values"] = vals
    out["info_maps"] = np.array([[refpin.volume_info_map(w, True, -1000.0, 3000.0, float(v)) for v in vals] for w in range(4)], np.float32)
    out["info_maps_raw"] = np.array([[refpin.volume_info_map(w, False, -1000.0, 3000.0, float(v)) for v in vals] for w in range(4)], np.float32)
    return out


LOADER_HEADERS = [
    "ObjectType = Image\nNDims = 3\nDimSize = 64 48 20\nElementSpacing = 0.5 0.5 1.25\nElementType = MET_SHORT\nElementDataFile = ct_head.raw\n",
    "NDims = 3\r\nDimSize = 7 8 9\r\nElementSize = 1 2 3\r\nElementType = MET_UCHAR\r\nElementDataFile = a.raw\r\n",
    "ElementDataFile = a.raw\nElementType = MET_FLOAT\nDimSize = 1 2 3\nElementSpacing = 1e-1 .5 2\n",
    "DimSize = 10 10 10\nElementSpacing = 1 1 1\nElementType = MET_USHORT\nCompressedData = True\nCompressedDataSize = 12345\nElementDataFile = z.zraw\n",
    "DimSize\t=\t3   4\n5\nElementSpacing = 1 1 1\nElementType = MET_INT\nElementDataFile = x.raw",
    "DimSize = 4 4 4\nElementType = MET_UCHAR\nElementDataFile = x.raw\n",                        # no spacing
    "ElementSpacing = 1 1 1\nElementType = MET_UCHAR\nElementDataFile = x.raw\n",                  # no DimSize
    "DimSize = 4 4 4\nElementSpacing = 1 1 1\nElementType = MET_DOUBLE\nElementDataFile = x.raw\n",  # unknown element type
    "DimSize = 4 4 4\nElementSpacing = 1 1 1\nElementType = MET_CHAR\n",                            # no data file
    "",
    "DimSize = 4 4 4\nElementSpacing = 1 1 1\nElementType = MET_UINT\nElementDataFile = with space.raw\n",  # a name is ONE word
]
INFO_FIELDS = ["parse_ok", "dims", "spacing", "world_dims", "original_format", "actual_format", "bytes_per_voxel", "is_signed", "is_normalized",
               "min_value", "max_value", "is_compressed", "compressed_bytes"]


def info_to_vector(i) -> np.ndarray:
    v = []
    for f in INFO_FIELDS:
        x = getattr(i, f)
        v.extend(list(x) if hasattr(x, "__len__") else [x])
    return np.array(v, np.float64)


def loaders_case():
    """The reference's own UMHDLoader::ParseVolumeInfoFromHeader on a set of headers, and the outcome of IVolumeLoader::ConvertData +
    FVolumeInfo::VoxelFormatToPixelFormat (through UMHDLoader::CreateVolumeFromFile) for every element type and flag combination."""
    import tempfile

    out = {}
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        parsed, names = [], []
        for h in LOADER_HEADERS:
            (d / "h.mhd").write_text(h, newline="")
            i = refpin.mhd_parse_file(d / "h.mhd")
            parsed.append(info_to_vector(i))
            names.append(i.data_file.decode())
        out["parsed"] = np.stack(parsed)
        out["data_files"] = np.array(names)
        table = []
        for fmt, dt in refpin.VOXEL_DTYPES.items():
            met = ["MET_UCHAR", "MET_CHAR", "MET_USHORT", "MET_SHORT", "MET_UINT", "MET_INT", "MET_FLOAT"][fmt]
            raw = (np.arange(2 * 3 * 4) % 7).astype(dt)
            (d / "v.raw").write_bytes(raw.tobytes())
            (d / "v.mhd").write_text(f"DimSize = 4 3 2\nElementSpacing = 1 1 1\nElementType = {met}\nElementDataFile = v.raw\n")
            for nrm in (0, 1):
                for flt in (0, 1):
                    info, tex, bulk = refpin.mhd_create_volume(d / "v.mhd", nrm, flt)
                    table.append([fmt, nrm, flt, tex, info.actual_format, info.bytes_per_voxel, info.is_normalized, len(bulk)])
        out["conversion_table"] = np.array(table, np.int64)
    return out


CASES = {"ref_hostmath": hostmath_case, "ref_shaders": shaders_case, "ref_materials": materials_case, "ref_ingest": ingest_case,
         "ref_loaders": loaders_case}

# the reference files the build compiles (oracle/ref.mk), relative to the reference checkout: their digests tie the golden vectors to the sources
REFERENCE_SOURCES = [
    "Source/Raymarcher/Private/Rendering/LightingShaderUtils.cpp", "Source/Raymarcher/Public/Rendering/LightingShaderUtils.h",
    "Source/Raymarcher/Public/Rendering/RaymarchTypes.h",
    "Source/Raymarcher/Shaders/Private/AddDirLightShader.usf", "Source/Raymarcher/Shaders/Private/ChangeDirLightShader.usf",
    "Source/Raymarcher/Shaders/Private/RaymarcherCommon.usf", "Source/Raymarcher/Shaders/Private/WindowedSampling.usf",
    "Source/Raymarcher/Shaders/Private/RaymarchMaterialCommon.usf", "Source/Raymarcher/Shaders/Private/WindowedRaymarchMaterials.usf",
    "Source/Raymarcher/Shaders/Private/GenerateOctreeShader.usf", "Source/Raymarcher/Shaders/Private/OctreeCommon.usf",
    "Source/FractalMarcher/Shaders/Private/SDFMarcher.usf", "Source/FractalMarcher/Shaders/Private/CalculateMandelbulbSDF.usf",
    "Source/VolumeTextureToolkit/Public/TextureUtilities.h", "Source/VolumeTextureToolkit/Public/VolumeAsset/VolumeInfo.h",
    "Source/VolumeTextureToolkit/Private/VolumeAsset/VolumeInfo.cpp",
    "Source/VolumeTextureToolkit/Private/VolumeAsset/Loaders/MHDLoader.cpp", "Source/VolumeTextureToolkit/Private/VolumeAsset/Loaders/VolumeLoader.cpp",
]


def reference_source_digests(root="/root/reference") -> dict:
    import hashlib

    return {rel: hashlib.sha256((Path(root) / rel).read_bytes()).hexdigest() for rel in REFERENCE_SOURCES}


if __name__ == "__main__":
    only = set(sys.argv[1:])
    if not only or "sources" in only:
        import json

        (HERE / "ref_sources.json").write_text(json.dumps({"reference": "tommybazar/TBRaymarcherPlugin @ e06b824 (SURVEY.md)",
                                                           "sha256": reference_source_digests()}, indent=1) + "\n")
        print("ref_sources.json written")
    for name, fn in CASES.items():
        if only and name not in only:
            continue
        arrays = fn()
        np.savez_compressed(HERE / f"{name}.npz", **arrays)
        print(name, {k: v.shape for k, v in arrays.items()})
