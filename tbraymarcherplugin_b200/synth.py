"""Deterministic synthetic inputs of SURVEY.md §8(d): volumes, transfer functions, lights, cameras, configs.

The volume generators are the numpy twins of csrc/synth.cu (fp64, same operation order) and reproduce the device
generators bit-for-bit; they are used where no GPU is available (CPU tests, the CPU baseline's inputs).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

from .raymarch_utils import (FCamera, FClippingPlaneParameters, FDirLightParameters, FMandelbulbParameters, FRaymarchWorldParameters,
                             FTransform, FWindowingParameters)

PERLIN_SEED = 0x5EED1234


def _lowbias32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def _lattice(x, y, z, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        h = x.astype(np.uint32) + np.uint32(374761393) * y.astype(np.uint32) + np.uint32(668265263) * z.astype(np.uint32) + np.uint32(seed & 0xFFFFFFFF)
        return _lowbias32(h).astype(np.float64) / 4294967296.0


def sphere_volume(dims: Tuple[int, int, int]) -> np.ndarray:
    """Sphere R8 (cfg 1): v = round(255*max(0, 1 - |uvw-0.5|/0.4)); returns uint8 [Z,Y,X]."""
    X, Y, Z = dims
    u = ((np.arange(X, dtype=np.float64) + 0.5) / X)[None, None, :] - 0.5
    v = ((np.arange(Y, dtype=np.float64) + 0.5) / Y)[None, :, None] - 0.5
    w = ((np.arange(Z, dtype=np.float64) + 0.5) / Z)[:, None, None] - 0.5
    d = np.sqrt(u * u + v * v + w * w)
    val = np.maximum(0.0, 1.0 - d / 0.4)
    return np.floor(255.0 * val + 0.5).astype(np.uint8)


def perlin_ct_volume(dims: Tuple[int, int, int], seed: int = PERLIN_SEED) -> np.ndarray:
    """CT-like 4-octave value noise with an ellipsoid body mask (cfg 2-4); returns uint8 [Z,Y,X]."""
    X, Y, Z = dims
    out = np.empty((Z, Y, X), dtype=np.uint8)
    u = ((np.arange(X, dtype=np.float64) + 0.5) / X)[None, :]
    v = ((np.arange(Y, dtype=np.float64) + 0.5) / Y)[:, None]
    u, v = np.broadcast_arrays(u, v)
    for iz in range(Z):
        w = np.full_like(u, (iz + 0.5) / Z)
        total = np.zeros_like(u)
        amp, norm = 1.0, 0.0
        for o in range(4):
            cells = float(4 << o)
            px, py, pz = u * cells, v * cells, w * cells
            fx0, fy0, fz0 = np.floor(px), np.floor(py), np.floor(pz)
            ix, iy, izc = fx0.astype(np.int64), fy0.astype(np.int64), fz0.astype(np.int64)
            fx, fy, fz = px - fx0, py - fy0, pz - fz0
            fx = fx * fx * (3.0 - 2.0 * fx)
            fy = fy * fy * (3.0 - 2.0 * fy)
            fz = fz * fz * (3.0 - 2.0 * fz)
            s = (seed + o * 0x9E3779B9) & 0xFFFFFFFF
            c000, c100 = _lattice(ix, iy, izc, s), _lattice(ix + 1, iy, izc, s)
            c010, c110 = _lattice(ix, iy + 1, izc, s), _lattice(ix + 1, iy + 1, izc, s)
            c001, c101 = _lattice(ix, iy, izc + 1, s), _lattice(ix + 1, iy, izc + 1, s)
            c011, c111 = _lattice(ix, iy + 1, izc + 1, s), _lattice(ix + 1, iy + 1, izc + 1, s)
            x00, x10 = c000 + fx * (c100 - c000), c010 + fx * (c110 - c010)
            x01, x11 = c001 + fx * (c101 - c001), c011 + fx * (c111 - c011)
            y0, y1 = x00 + fy * (x10 - x00), x01 + fy * (x11 - x01)
            total = total + amp * (y0 + fz * (y1 - y0))
            norm = norm + amp
            amp = amp * 0.5
        noise = total / norm
        eu, ev, ew = (u - 0.5) / 0.45, (v - 0.5) / 0.40, (w - 0.5) / 0.48
        e = np.sqrt(eu * eu + ev * ev + ew * ew)
        mask = np.minimum(1.0, np.maximum(0.0, (1.0 - e) / 0.02))
        out[iz] = np.floor(255.0 * (noise * mask) + 0.5).astype(np.uint8)
    return out


# ---- transfer functions -----------------------------------------------------------------------------------------
def default_ramp_curve() -> np.ndarray:
    """MakeDefaultTFTexture (RaymarchUtils.cpp:113-141): (t, t, t, 1)."""
    t = (np.arange(256, dtype=np.float32) / np.float32(255.0)).astype(np.float32)
    return np.stack([t, t, t, np.ones_like(t)], axis=1).astype(np.float32)


def soft_ct_curve() -> np.ndarray:
    """'soft_ct' colour curve sampled at i/255 like ColorCurveToTexture (RaymarchUtils.cpp:153-162)."""
    keys = np.array([0.0, 0.25, 0.5, 1.0])
    vals = np.array([[0, 0, 0, 0], [0.8, 0.3, 0.2, 0.005], [0.9, 0.8, 0.6, 0.03], [1.0, 1.0, 1.0, 0.15]], dtype=np.float64)
    t = np.arange(256, dtype=np.float64) / 255.0
    return np.stack([np.interp(t, keys, vals[:, c]) for c in range(4)], axis=1).astype(np.float32)


# ---- lights / world / camera ---------------------------------------------------------------------------------------
def _unit(v):
    n = math.sqrt(sum(c * c for c in v))
    return tuple(c / n for c in v)


LIGHTS: List[FDirLightParameters] = [
    FDirLightParameters(_unit((1.0, 0.4, -0.3)), 1.0),
    FDirLightParameters(_unit((-0.2, -1.0, -0.5)), 0.6),
    FDirLightParameters(_unit((0.3, 0.2, -1.0)), 0.8),
    FDirLightParameters((0.0, 0.0, -1.0), 0.5),
]


def rotate_about_z(light: FDirLightParameters, degrees: float) -> FDirLightParameters:
    a = math.radians(degrees)
    x, y, z = light.LightDirection
    return FDirLightParameters((x * math.cos(a) - y * math.sin(a), x * math.sin(a) + y * math.cos(a), z), light.LightIntensity)


def identity_world() -> FRaymarchWorldParameters:
    return FRaymarchWorldParameters(FTransform(), FClippingPlaneParameters())


def scaled_rotated_world() -> FRaymarchWorldParameters:
    """The parity case with scale (1,1.3,0.7) and a 30 degree yaw."""
    return FRaymarchWorldParameters(FTransform.from_axis_angle((0, 0, 1), 30.0, scale=(1.0, 1.3, 0.7)), FClippingPlaneParameters())


def clipped_world() -> FRaymarchWorldParameters:
    """Clip plane through the volume centre with direction normalize(1,1,0)."""
    return FRaymarchWorldParameters(FTransform(), FClippingPlaneParameters((0.0, 0.0, 0.0), _unit((1.0, 1.0, 0.0))))


def benchmark_camera(width: int, height: int, jitter: bool = True, frame: int = 0) -> FCamera:
    return FCamera((-0.9, -0.5, 0.7), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 60.0, width, height, 0.0, frame, jitter)


CONFIGS: Dict[str, dict] = {
    # BASELINE.json configs[0..4]
    "cfg1": dict(volume="sphere", n=256, view=(512, 512), steps=256, lights=[0], windowing=FWindowingParameters(), tf="default_ramp"),
    "cfg2": dict(volume="perlin", n=512, view=(1920, 1080), steps=512, lights=[0, 1],
                 windowing=FWindowingParameters(0.45, 0.5, True, False), tf="soft_ct"),
    "cfg3": dict(volume="perlin", n=512, view=None, steps=None, lights=[0, 1, 2, 3],
                 windowing=FWindowingParameters(0.45, 0.5, True, False), tf="soft_ct", updates=16, degrees=5.0),
    "cfg4": dict(volume="perlin", n=1024, view=(3840, 2160), steps=768, lights=[0, 1, 2],
                 windowing=FWindowingParameters(0.45, 0.5, True, False), tf="soft_ct"),
    "cfg5": dict(volume=None, view=(1920, 1080), mandelbulb=FMandelbulbParameters()),
}


def make_volume(kind: str, n: int) -> np.ndarray:
    return sphere_volume((n, n, n)) if kind == "sphere" else perlin_ct_volume((n, n, n))


def make_curve(name: str) -> np.ndarray:
    return default_ramp_curve() if name == "default_ramp" else soft_ct_curve()
