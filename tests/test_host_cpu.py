"""CPU-side tests of the product's host logic: the C-ABI library loads and exports every declared symbol, its host
parameter math (csrc/host_plan.hpp) agrees bit-for-bit with the oracle's restatement of LightingShaderUtils.cpp, and
argument validation / error behaviour follows the reference. No compute calls (there is no GPU here)."""
import ctypes as C
import math
import re
from pathlib import Path

import numpy as np
import pytest

import oracle
from tbraymarcherplugin_b200 import _capi, synth
from tbraymarcherplugin_b200.raymarch_utils import (FBasicRaymarchRenderingResources, FClippingPlaneParameters, FDirLightParameters,
                                                    FRaymarchWorldParameters, FTransform, FWindowingParameters, URaymarchUtils,
                                                    plan_dir_light)

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_symbol_the_header_declares():
    header = (ROOT / "include" / "tbrm.h").read_text()
    declared = set(re.findall(r"\b(tbrm_[a-z0-9_]+)\s*\(", header))
    lib = _capi.load()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"libtbrm.so does not export {name}"
    assert declared == set(_capi.PROTOTYPES), declared ^ set(_capi.PROTOTYPES)
    assert lib.tbrm_abi_version() == 1


def test_struct_layouts_match_between_binding_and_oracle():
    assert C.sizeof(_capi.LightPlan) == C.sizeof(oracle.LightPlan)
    assert C.sizeof(_capi.PassPlan) == C.sizeof(oracle.Pass)


def _random_world(rng):
    axis = rng.standard_normal(3)
    t = FTransform.from_axis_angle(tuple(axis), float(rng.uniform(-180, 180)), tuple(rng.uniform(-50, 50, 3)), tuple(rng.uniform(0.3, 3.0, 3)))
    c = FClippingPlaneParameters(tuple(rng.uniform(-40, 40, 3)), tuple(rng.standard_normal(3)))
    return FRaymarchWorldParameters(t, c)


def _plans_equal(a, b):
    for f, _ in _capi.LightPlan._fields_:
        va, vb = getattr(a, f), getattr(b, f)
        if f == "passes":
            for pa, pb in zip(va, vb):
                for g, _ in _capi.PassPlan._fields_:
                    xa, xb = getattr(pa, g), getattr(pb, g)
                    xa = list(xa) if hasattr(xa, "__len__") else xa
                    xb = list(xb) if hasattr(xb, "__len__") else xb
                    assert xa == xb, (g, xa, xb)
        else:
            va = list(va) if hasattr(va, "__len__") else va
            vb = list(vb) if hasattr(vb, "__len__") else vb
            assert va == vb, (f, va, vb)


def test_host_plan_matches_oracle_bit_for_bit():
    rng = np.random.default_rng(7)
    worlds = [synth.identity_world(), synth.scaled_rotated_world(), synth.clipped_world()] + [_random_world(rng) for _ in range(40)]
    for i, world in enumerate(worlds):
        dims = tuple(int(d) for d in rng.integers(5, 300, 3))
        win = FWindowingParameters(float(rng.uniform(0, 1)), float(rng.uniform(0.05, 1.5)), bool(i & 1), bool(i & 2))
        light = FDirLightParameters(tuple(rng.standard_normal(3) * 3), float(rng.uniform(0.1, 1.5)))
        for exact in (False, True):
            _plans_equal(plan_dir_light(dims, win, light, world, exact), oracle.plan_dir_light(dims, win, light, world, exact))
    for light in synth.LIGHTS:  # the benchmark lights, identity transform
        _plans_equal(plan_dir_light((512, 512, 512), FWindowingParameters(), light, worlds[0]),
                     oracle.plan_dir_light((512, 512, 512), FWindowingParameters(), light, worlds[0]))


def test_plan_semantics():
    w = FWindowingParameters()
    world = synth.identity_world()
    # axis-aligned light: weight 1 on +Z, one pass, sweep from the top slice downwards (LightingShaderUtils.cpp:66-70,181-187)
    p = plan_dir_light((32, 48, 64), w, FDirLightParameters((0, 0, -1), 0.5), world)
    assert p.add_passes == 1 and p.passes[0].face == 4 and p.passes[0].weight == 1.0 and p.passes[1].weight == 0.0
    assert p.passes[0].dirn == -1 and p.passes[0].start == 63 and p.passes[0].stop == -1 and list(p.passes[0].td) == [32, 48, 64]
    assert p.passes[0].light_alpha == 0.5 and list(p.passes[0].uv_offset) == [0.0, 0.0]
    assert p.passes[0].step_size == pytest.approx(1 / 64) and list(p.passes[0].uvw_offset) == pytest.approx([0, 0, 1 / 32])
    # L1: faces -X then -Y, weights 0.8 / 0.2 (third axis folded into the second), transposed dims (Y,Z,X) / (X,Z,Y)
    p = plan_dir_light((32, 48, 64), w, synth.LIGHTS[0], world)
    assert p.add_passes == 2 and [q.face for q in p.passes] == [1, 3]
    assert p.passes[0].weight + p.passes[1].weight == pytest.approx(1.0)
    assert list(p.passes[0].td) == [48, 64, 32] and list(p.passes[1].td) == [32, 64, 48]
    assert p.passes[0].dirn == 1 and p.passes[0].start == 0 and p.passes[0].stop == 32
    # no clipping plane: "ridiculously far and facing away"
    assert list(p.clip_center) == pytest.approx([0.5, 0.5, 100000.5]) and list(p.clip_dir) == [0.0, 0.0, -1.0]
    # border colours: 8-bit round trip by default, exact on request
    assert p.passes[0].border != p.passes[0].light_alpha
    assert abs(p.passes[0].border - p.passes[0].light_alpha) < 4e-3
    pe = plan_dir_light((32, 48, 64), w, synth.LIGHTS[0], world, border_exact=True)
    assert pe.passes[0].border == pe.passes[0].light_alpha and pe.data_border == 0.0
    assert plan_dir_light((8, 8, 8), w, FDirLightParameters((0, 0, 0), 1.0), world).zero_direction == 1


def test_uninitialised_resources_report_light_not_added():
    res = FBasicRaymarchRenderingResources()  # no handle: every pointer the reference checks is null
    assert URaymarchUtils.AddDirLightToSingleVolume(res, synth.LIGHTS[0], True, synth.identity_world()) is False
    assert URaymarchUtils.ChangeDirLightInSingleVolume(res, synth.LIGHTS[0], synth.LIGHTS[1], synth.identity_world()) is False
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)  # silently returns


def test_create_without_a_device_fails_loudly():
    lib = _capi.load()
    if lib.tbrm_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_capi.TbrmError) as e:
        URaymarchUtils.InitializeRaymarchResources((8, 8, 8))
    assert e.value.status == _capi.TBRM_ERR_NO_DEVICE


def test_argument_validation():
    lib = _capi.load()
    h = C.c_void_p()
    assert lib.tbrm_create(0, (C.c_int32 * 3)(0, 4, 4), 0, 2, 0, C.byref(h)) == _capi.TBRM_ERR_INVALID_ARGUMENT
    assert lib.tbrm_create(0, (C.c_int32 * 3)(4, 4, 4), 0, 1, 0, C.byref(h)) == _capi.TBRM_ERR_INVALID_ARGUMENT  # G16 light volume
    assert b"light volume" in lib.tbrm_last_error()
    assert lib.tbrm_flush(None) == _capi.TBRM_ERR_INVALID_ARGUMENT


def test_synthetic_volumes_are_deterministic_and_shaped():
    s = synth.sphere_volume((16, 12, 10))
    assert s.shape == (10, 12, 16) and s.dtype == np.uint8 and s.max() > 200 and s[0, 0, 0] == 0
    p1, p2 = synth.perlin_ct_volume((24, 24, 24)), synth.perlin_ct_volume((24, 24, 24))
    assert np.array_equal(p1, p2) and p1[0, 0, 0] == 0 and 60 < p1[12, 12, 12] < 200
    assert not np.array_equal(p1, synth.perlin_ct_volume((24, 24, 24), seed=1))
    c = synth.soft_ct_curve()
    assert c.shape == (256, 4) and c[0, 3] == 0 and c[255, 3] == pytest.approx(0.15)


def test_cpp_host_mirror_compiles_against_the_c_abi(tmp_path):
    """csrc/RaymarchUtils.hpp (the reference's signatures over the C ABI) is plain C++17 and links against libtbrm.so."""
    import subprocess

    from tbraymarcherplugin_b200 import build

    src = tmp_path / "t.cpp"
    src.write_text('''
        #include "tbraymarcherplugin_b200/csrc/RaymarchUtils.hpp"
        using namespace tbrm_ue;
        int main() {
            FBasicRaymarchRenderingResources res;  // uninitialised: every resource pointer is null
            bool added = true;
            URaymarchUtils::AddDirLightToSingleVolume(res, FDirLightParameters(FVector(1, 0, 0), 1.0f), true, FRaymarchWorldParameters(), added, true);
            if (added) return 1;  // LightAdded must be false (RaymarchUtils.cpp:39-45)
            added = true;
            URaymarchUtils::ChangeDirLightInSingleVolume(res, FDirLightParameters(), FDirLightParameters(), FRaymarchWorldParameters(), added);
            URaymarchUtils::ClearResourceLightVolumes(res, 0.0f);
            return added ? 2 : 0;
        }''')
    exe = tmp_path / "t"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-I", str(ROOT), str(src), "-o", str(exe), str(build.LIB_PATH), f"-Wl,-rpath,{build.PKG_DIR}"], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_cpp_volume_policy_mirror(tmp_path):
    """csrc/RaymarchVolume.hpp (ARaymarchVolume::Tick policy in C++) with a recording operator surface: the same decisions as
    tests/test_volume_policy_cpu.py checks for the Python twin."""
    import subprocess

    src = tmp_path / "p.cpp"
    src.write_text('''
        #include <cstdio>
        #include "tbraymarcherplugin_b200/csrc/RaymarchVolume.hpp"
        using namespace tbrm_ue;
        static std::vector<std::string> calls;
        struct Rec {
            static void ClearResourceLightVolumes(FBasicRaymarchRenderingResources, float) { calls.push_back("clear"); }
            static void AddDirLightToSingleVolume(const FBasicRaymarchRenderingResources&, const FDirLightParameters&, bool, FRaymarchWorldParameters, bool& ok, bool) { calls.push_back("add"); ok = true; }
            static void ChangeDirLightInSingleVolume(FBasicRaymarchRenderingResources&, FDirLightParameters o, FDirLightParameters n, FRaymarchWorldParameters, bool& ok, bool) {
                calls.push_back(o.LightIntensity == 1.0f && n.LightIntensity != 1.0f ? "change" : "change?"); ok = true; }
            static void GenerateOctree(FBasicRaymarchRenderingResources&) { calls.push_back("octree"); }
        };
        #define CHECK(c) do { if (!(c)) { std::printf("failed: %s (line %d)\\n", #c, __LINE__); return 1; } } while (0)
        int main() {
            ARaymarchLight L[4];
            for (int i = 0; i < 4; ++i) L[i].ForwardVector = FVector(1, 0.1 * i, -0.3);
            ARaymarchVolume<Rec> vol;
            vol.RaymarchResources.bIsInitialized = true;
            for (auto& l : L) vol.LightsArray.push_back(&l);
            vol.OnConstruction();
            CHECK(vol.Tick().action == FTickReport::None && calls.empty());
            L[2].LightIntensity = 0.5f;                       // one light changed: incremental, old parameters from the map
            CHECK(vol.Tick().action == FTickReport::Incremental && calls.size() == 1 && calls[0] == "change");
            CHECK(vol.Tick().action == FTickReport::None);
            calls.clear();
            L[0].LightIntensity = 0.2f; L[1].LightIntensity = 0.3f;   // 2 of 4 changed: "> 1 && >= half" -> full reset
            CHECK(vol.Tick().action == FTickReport::Reset && calls.size() == 5 && calls[0] == "clear" && calls[4] == "add");
            calls.clear();
            CHECK(vol.Tick().action == FTickReport::Reset);  // reference quirk: the map is stale after a reset, the rule fires again
            vol.bRefreshLightMapOnReset = true;
            CHECK(vol.Tick().action == FTickReport::Reset && vol.Tick().action == FTickReport::None);
            calls.clear();
            vol.ComponentTransform.Translation = FVector(5, 0, 0);  // world change: full reset
            CHECK(vol.Tick().action == FTickReport::Reset && !vol.bRequestedRecompute);
            vol.ComponentTransform.Translation = FVector(5 + 5e-5, 0, 0);  // within FTransform::Equals' tolerance
            CHECK(vol.Tick().action == FTickReport::None);
            vol.SelectRaymarchMaterial = ERaymarchMaterial::Intensity;
            L[3].LightIntensity = 0.1f;
            calls.clear();
            CHECK(vol.Tick().action == FTickReport::None && calls.empty());
            vol.bRequestedOctreeRebuild = true;                  // a new volume was loaded (RaymarchVolume.cpp:553-554)
            CHECK(!vol.Tick().octree_rebuilt && calls.empty());  // not under the intensity material
            vol.SelectRaymarchMaterial = ERaymarchMaterial::Octree;
            CHECK(vol.Tick().octree_rebuilt && calls.size() == 1 && calls[0] == "octree" && !vol.bRequestedOctreeRebuild);
            CHECK(!vol.Tick().octree_rebuilt && calls.size() == 1);
            vol.RaymarchResources.bIsInitialized = false;
            CHECK(vol.Tick().action == FTickReport::NotInitialized);
            return 0;
        }''')
    exe = tmp_path / "p"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-I", str(ROOT), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout


def test_exact_division_free_sequences():
    """The fast kernels replace v/255 by fma(v, c_hi, v*c_lo) and x/W by Markstein's sequence; both must equal IEEE division."""
    v = np.arange(256, dtype=np.float32)
    c_hi, c_lo = np.float32(0.003921568859368563), np.float32(-2.319175823606301e-10)
    got = (v.astype(np.float64) * float(c_hi) + (v * c_lo).astype(np.float64)).astype(np.float32)  # fma: exact product, one rounding
    assert np.array_equal(got, v / np.float32(255.0))
    v = np.arange(65536, dtype=np.float32)  # UNORM16 texels of the octree march (decode_u16_exact, csrc/materials.cuh)
    c_hi, c_lo = np.float32(1.5259021893143654e-05), np.float32(3.5527678889091252e-15)
    got = (v.astype(np.float64) * float(c_hi) + (v * c_lo).astype(np.float64)).astype(np.float32)
    assert np.array_equal(got, v / np.float32(65535.0))
    rng = np.random.default_rng(3)
    for w in (0.5, 1.0, 0.4, 0.3, 0.9, 1.7):
        w32 = np.float32(w)
        rw = np.float32(1.0) / w32
        x = np.concatenate([rng.uniform(-1.5, 2.0, 20000), (np.arange(-20, 280) / 255.0)]).astype(np.float32)
        q = x * rw
        e = (-(q.astype(np.float64)) * float(w32) + x.astype(np.float64)).astype(np.float32)
        q2 = (e.astype(np.float64) * float(rw) + q.astype(np.float64)).astype(np.float32)
        assert np.array_equal(q2, x / w32), w


def test_small_helpers_of_the_function_library():
    """RaymarchUtils.cpp:219-252: coordinate helpers and TransformToMatrix (FTransform::ToMatrixWithScale: row-vector convention)."""
    assert URaymarchUtils.LocalToTextureCoords((-1.0, 0.0, 1.0)) == (0.0, 0.5, 1.0)
    assert URaymarchUtils.TextureToLocalCoords((0.0, 0.5, 1.0)) == (-1.0, 0.0, 1.0)
    assert URaymarchUtils.GetVolumeTextureDimensions(None) == (0, 0, 0)
    assert URaymarchUtils.GetVolumeTextureDimensions(FBasicRaymarchRenderingResources()) == (0, 0, 0)
    t = FTransform.from_axis_angle((0, 0, 1), 90.0, (5.0, 6.0, 7.0), (2.0, 3.0, 4.0))
    m = URaymarchUtils.TransformToMatrix(t)
    p = np.array([1.0, 0.0, 0.0, 1.0]) @ m  # scale, rotate +90 degrees about Z, translate
    assert np.allclose(p[:3], (5.0, 8.0, 7.0)) and np.allclose(m[3], (5, 6, 7, 1))
    r = URaymarchUtils.TransformToMatrix(t, WithScaling=False)
    assert np.allclose(r[:3, :3] @ r[:3, :3].T, np.eye(3)) and np.allclose(np.linalg.det(r[:3, :3]), 1.0)
    # consistent with the library's InverseTransformPosition (the WorldToLocal the kernels use): world -> local -> world
    lp = np.linalg.inv(m)
    assert np.allclose((np.array([5.0, 8.0, 7.0, 1.0]) @ lp)[:3], (1.0, 0.0, 0.0))


def test_joined_lights_entry_point_validates_like_add_dir_light():
    lib = _capi.load()
    n = C.c_int(7)
    st = _capi.SweepStats()
    w = synth.identity_world().to_c()
    arr = (_capi.DirLight * 2)(synth.LIGHTS[0].to_c(), synth.LIGHTS[1].to_c())
    # a missing resource set: status NOT_INITIALIZED and *lights_added = 0 (the reference's `LightAdded = false`, RaymarchUtils.cpp:39-49)
    assert lib.tbrm_add_dir_lights_joined(None, arr, 2, 1, C.byref(w), C.byref(n), C.byref(st)) == _capi.TBRM_ERR_NOT_INITIALIZED
    assert n.value == 0 and st.passes == 0
    assert URaymarchUtils.AddDirLightsToSingleVolumeJoined(FBasicRaymarchRenderingResources(), synth.LIGHTS, True, synth.identity_world()) is False


def test_kernel_source_of_the_power8_mandelbulb_iteration_equals_the_oracle_twin_on_the_host():
    """csrc/mandelbulb.cu's mandelbulb_sdf_p8 is one __host__ __device__ function: its HOST compilation (tbrm_debug_mandelbulb_sdf_p8) must give
    the oracle twin's bits (same +, -, *, /, sqrt, and here even the same libm log) — on a lattice through the bulb, next to its surface,
    at the origin (NaN either way) and far outside. What runs on the GPU is this source compiled for sm_100a with --fmad=false."""
    lib, ora = _capi.load(), oracle.lib()
    ora.tbo_mandelbulb_sdf_at.restype = C.c_float
    ora.tbo_mandelbulb_sdf_at.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    rng = np.random.default_rng(8)
    g = np.linspace(-1.3, 1.3, 14, dtype=np.float32)
    pts = [np.array(p, np.float32) for p in np.stack(np.meshgrid(g, g, g), -1).reshape(-1, 3)]
    pts += [v / np.linalg.norm(v) * np.float32(r) for v, r in zip(rng.standard_normal((400, 3)).astype(np.float32), rng.uniform(0.7, 1.25, 400))]
    pts += [np.zeros(3, np.float32), np.array([0, 0, 1.0], np.float32), np.array([5, -7, 3], np.float32), np.array([1e-30, 0, 0], np.float32)]
    n_inside = 0
    for iterations, bailout in ((16, 2.4), (50, 2.0)):
        for p in pts:
            pos = (C.c_float * 3)(*p)
            a_it, b_it = C.c_uint32(), C.c_uint64()
            a = lib.tbrm_debug_mandelbulb_sdf_p8(pos, bailout, iterations, C.byref(a_it))
            b = ora.tbo_mandelbulb_sdf_at(pos, bailout, 8.0, iterations, 1, C.byref(b_it))
            assert a_it.value == b_it.value and (np.float32(a).tobytes() == np.float32(b).tobytes() or (np.isnan(a) and np.isnan(b))), (p, a, b)
            n_inside += a_it.value == iterations
    assert n_inside > 50  # the lattice does reach the inside of the bulb


def test_unorm8_to_unorm16_through_the_float_round_trip_is_byte_replication():
    """GenerateOctreeShader.usf:36-49 stores Volume.Load(p).r into a UNORM16 UAV: floor(saturate(b / 255) * 65535 + .5) in fp32. For every byte
    that is b * 257 — what octree_build_u8x16_kernel computes with one byte permute (csrc/materials.cuh)."""
    b = np.arange(256, dtype=np.float32)
    v = np.clip(b / np.float32(255.0), np.float32(0), np.float32(1)) * np.float32(65535.0) + np.float32(0.5)
    assert v.dtype == np.float32 and np.array_equal(np.floor(v).astype(np.int64), np.arange(256) * 257)
    w = np.arange(65536, dtype=np.float32)  # UNORM16 data: the same round trip is the identity (octree_texel, csrc/materials.cuh)
    v = np.clip(w / np.float32(65535.0), np.float32(0), np.float32(1)) * np.float32(65535.0) + np.float32(0.5)
    assert np.array_equal(np.floor(v).astype(np.int64), np.arange(65536))
