"""Generate the polynomial coefficients of the deterministic fp32 pow used by BOTH the oracle
(oracle/tbrm_oracle.cpp: det_log2 / det_exp2) and the CUDA kernels (csrc/tbrm_math.cuh).

HLSL defines pow(x,y) = exp2(y*log2(x)) with implementation-defined precision.
To make oracle-vs-kernel parity bit-exact we fix one concrete fp32 evaluation order:

  log2(x):  x = m * 2^e with m in [sqrt(1/2), sqrt(2));  t = m - 1;  log2 = e + t * P(t)     (Horner, fmaf)
  exp2(z):  n = floor(z + 0.5);  f = z - n in [-0.5, 0.5];  exp2 = Q(f) * 2^n               (Horner, fmaf)

Coefficients are a Chebyshev-node least-squares fit in float64, rounded to float32; the script then measures
the fp32 error of the whole pow against float64 and prints C initialisers.  Run: python oracle/gen_pow_coeffs.py
"""
import numpy as np

def cheb_nodes(a, b, n):
    k = np.arange(n)
    return 0.5 * (a + b) + 0.5 * (b - a) * np.cos(np.pi * (2 * k + 1) / (2 * n))

def fit(fn, a, b, deg, n=4000, fixed0=None):
    x = cheb_nodes(a, b, n)
    y = fn(x)
    if fixed0 is None:
        V = np.vander(x, deg + 1, increasing=True)
        c, *_ = np.linalg.lstsq(V, y, rcond=None)
        return c
    V = np.vander(x, deg + 1, increasing=True)[:, 1:]
    c, *_ = np.linalg.lstsq(V, y - fixed0, rcond=None)
    return np.concatenate([[fixed0], c])

LOG_DEG, EXP_DEG = 9, 6
lo, hi = np.sqrt(0.5) - 1.0, np.sqrt(2.0) - 1.0
def P(t):
    out = np.full_like(t, 1.0 / np.log(2.0))
    nz = np.abs(t) > 1e-12
    out[nz] = np.log2(1.0 + t[nz]) / t[nz]
    return out
cl = fit(P, lo, hi, LOG_DEG).astype(np.float32)
ce = fit(lambda f: np.exp2(f), -0.5, 0.5, EXP_DEG, fixed0=1.0).astype(np.float32)

def f32(x): return np.asarray(x, dtype=np.float32)
def fma(a, b, c):  # emulate fmaf with float64 (exact product of two f32 fits f64; one rounding to f32)
    return f32(a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64))

def det_log2(x):
    x = f32(x); bits = x.view(np.int32)
    e = ((bits >> 23) & 0xff) - 127
    m = ((bits & 0x007fffff) | 0x3f800000).view(np.float32)
    big = m > f32(1.41421356)
    m = np.where(big, m * f32(0.5), m); e = np.where(big, e + 1, e)
    t = f32(m - f32(1.0))
    p = np.full_like(t, cl[-1])
    for c in cl[-2::-1]:
        p = fma(p, t, np.full_like(t, c))
    return fma(t, p, f32(e))

def det_exp2(z):
    z = f32(z); n = np.floor(f32(z + f32(0.5))); f = f32(z - n)
    q = np.full_like(f, ce[-1])
    for c in ce[-2::-1]:
        q = fma(q, f, np.full_like(f, c))
    ni = n.astype(np.int32)
    out = (q.view(np.int32) + (ni << 23)).view(np.float32)
    return np.where(n < -125, f32(0), out)

def det_pow(x, y):
    return np.where(x <= 0, f32(0), det_exp2(f32(det_log2(x) * f32(y))))

if __name__ == "__main__":
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.random(2_000_000), 1 - 10 ** rng.uniform(-7, 0, 1_000_000), np.arange(0, 256) / 255.0]).astype(np.float32)
    x = x[(x > 0) & (x <= 1)]
    worst = 0.0
    for y in [0.05, 0.1953125, 0.2, 0.39, 1.0, 3.0, 17.3, 100.0]:
        got = det_pow(x, np.float32(y)).astype(np.float64)
        ref = np.power(x.astype(np.float64), np.float64(np.float32(y)))
        err = np.abs(got - ref)
        worst = max(worst, err.max())
        print(f"y={y:9.4f}  max abs err {err.max():.3e}  max rel err {np.max(err / np.maximum(ref, 1e-30) * (ref > 1e-30)):.3e}")
    l = det_log2(x).astype(np.float64); lr = np.log2(x.astype(np.float64))
    print("log2 max abs err", np.abs(l - lr).max(), " max rel", np.max(np.abs(l - lr) / np.maximum(np.abs(lr), 1e-30)))
    assert det_pow(f32([1.0]), f32(0.3))[0] == 1.0 and det_pow(f32([0.0]), f32(0.3))[0] == 0.0
    print("worst abs err of pow:", worst)
    print("static const float kLog2P[%d] = {%s};" % (len(cl), ", ".join(f"{float(c)!r}f" for c in cl)))
    print("static const float kExp2Q[%d] = {%s};" % (len(ce), ", ".join(f"{float(c)!r}f" for c in ce)))
