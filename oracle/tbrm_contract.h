// tbrm_contract.h — the one piece of the arithmetic contract that is shared source between the CPU oracle
// (tbrm_oracle.cpp) and the runner of the reference's shaders (ref_shaders.cpp): HLSL's pow(x,y) = exp2(y*log2(x)) has
// implementation-defined precision, so the contract fixes ONE evaluation of it — fixed polynomials
// from oracle/gen_pow_coeffs.py, |error| < 1e-7 vs libm. TEST INFRASTRUCTURE like everything under oracle/; the CUDA kernels
// carry their own copy of the same polynomials (csrc/tbrm_math.cuh).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace tbrm_contract {

const float kLog2P[10] = {1.4426950216293335f,  -0.7213473320007324f, 0.4808982014656067f,  -0.3606966435909271f,
                          0.2885688841342926f,  -0.23961904644966125f, 0.2045752853155136f, -0.19102497398853302f,
                          0.18631209433078766f, -0.11020159721374512f};
const float kExp2Q[7] = {1.0f,
                         0.6931471824645996f,
                         0.24022650718688965f,
                         0.05550327152013779f,
                         0.009618035517632961f,
                         0.0013400432653725147f,
                         0.00015467364573851228f};

inline float det_log2(float x) {  // x normal, > 0
    int32_t bits;
    std::memcpy(&bits, &x, 4);
    int32_t e = ((bits >> 23) & 0xff) - 127;
    int32_t mb = (bits & 0x007fffff) | 0x3f800000;
    float m;
    std::memcpy(&m, &mb, 4);
    if (m > 1.41421356f) {
        m = m * 0.5f;
        e += 1;
    }
    float t = m - 1.0f;
    float p = kLog2P[9];
    for (int i = 8; i >= 0; --i) p = fmaf(p, t, kLog2P[i]);
    return fmaf(t, p, (float) e);
}
inline float det_exp2(float z) {
    float n = floorf(z + 0.5f);
    if (n < -125.0f) return 0.0f;
    float f = z - n;
    float q = kExp2Q[6];
    for (int i = 5; i >= 0; --i) q = fmaf(q, f, kExp2Q[i]);
    int32_t qb;
    std::memcpy(&qb, &q, 4);
    qb += ((int32_t) n) << 23;
    float r;
    std::memcpy(&r, &qb, 4);
    return r;
}
// pow(x, y) for y > 0, x <= 1 (the only use: 1 - pow(1 - a, StepSize), WindowedSampling.usf:35)
inline float det_pow(float x, float y) {
    if (!(x >= 1.17549435e-38f)) return 0.0f;
    return det_exp2(y * det_log2(x));
}

}  // namespace tbrm_contract
