// tbrm_math.cuh — fp32 arithmetic contract of the hot path (device side).
//
// The translation unit is compiled with --fmad=false: every +,-,*,/ below is one correctly rounded IEEE op,
// fused multiply-adds happen only where __fmaf_rn is written. DESIGN.md §4 states the contract; the CPU oracle
// implements the same contract independently, which is what makes parity bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tbrm {

__device__ __forceinline__ float lerpf(float a, float b, float t) { return __fmaf_rn(t, b - a, a); }
__device__ __forceinline__ float saturatef(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }  // NaN -> 0
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return ((ax * bx) + (ay * by)) + (az * bz);
}

// pow(x,y) = exp2(y*log2(x)) for 0 <= x <= 1, y > 0 with fixed polynomials (oracle/gen_pow_coeffs.py):
//   log2: x = m*2^e, m in [sqrt(1/2), sqrt(2)), t = m-1, log2 = e + t*P(t);  exp2: n = floor(z+.5), f = z-n, Q(f)*2^n
__device__ __forceinline__ float det_log2(float x) {
    int bits = __float_as_int(x);
    int e = ((bits >> 23) & 0xff) - 127;
    float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
    if (m > 1.41421356f) {
        m = m * 0.5f;
        e += 1;
    }
    float t = m - 1.0f;
    float p = -0.11020159721374512f;
    p = __fmaf_rn(p, t, 0.18631209433078766f);
    p = __fmaf_rn(p, t, -0.19102497398853302f);
    p = __fmaf_rn(p, t, 0.2045752853155136f);
    p = __fmaf_rn(p, t, -0.23961904644966125f);
    p = __fmaf_rn(p, t, 0.2885688841342926f);
    p = __fmaf_rn(p, t, -0.3606966435909271f);
    p = __fmaf_rn(p, t, 0.4808982014656067f);
    p = __fmaf_rn(p, t, -0.7213473320007324f);
    p = __fmaf_rn(p, t, 1.4426950216293335f);
    return __fmaf_rn(t, p, (float) e);
}
__device__ __forceinline__ float det_exp2(float z) {
    float n = floorf(z + 0.5f);
    if (n < -125.0f) return 0.0f;
    float f = z - n;
    float q = 0.00015467364573851228f;
    q = __fmaf_rn(q, f, 0.0013400432653725147f);
    q = __fmaf_rn(q, f, 0.009618035517632961f);
    q = __fmaf_rn(q, f, 0.05550327152013779f);
    q = __fmaf_rn(q, f, 0.24022650718688965f);
    q = __fmaf_rn(q, f, 0.6931471824645996f);
    q = __fmaf_rn(q, f, 1.0f);
    return __int_as_float(__float_as_int(q) + (((int) n) << 23));
}
__device__ __forceinline__ float det_pow(float x, float y) {
    if (!(x >= 1.17549435e-38f)) return 0.0f;
    return det_exp2(y * det_log2(x));
}

// one axis of a linear SampleLevel: x = u*N - 0.5; taps floor(x), floor(x)+1
__device__ __forceinline__ void axis_taps(float u, int N, int& i0, float& f) {
    float x = u * (float) N - 0.5f;
    float fl = floorf(x);
    f = x - fl;
    fl = fminf(fmaxf(fl, -4.0f), (float) N + 4.0f);
    i0 = (int) fl;
}

__device__ __forceinline__ uint8_t quant8(float v) { return (uint8_t) floorf(saturatef(v) * 255.0f + 0.5f); }

// windowing uniform block (FWindowingParameters::ToLinearColor, VolumeInfo.h:49-52)
struct Windowing {
    float center, width, low, high;
};

// GetTransferFuncPosition + cut-offs (WindowedSampling.usf:14-29). Returns false when the sample is cut off.
__device__ __forceinline__ bool tf_position(float v, const Windowing& w, float& pos) {
    pos = (v - w.center + (w.width / 2.0f)) / w.width;
    return !((pos < 0.0f && w.low > 0.0f) || (pos > 1.0f && w.high > 0.0f));
}
// bilinear clamp lookup in the collapsed 256-entry TF: taps and weight
__device__ __forceinline__ void tf_taps(float pos, int& i0, int& i1, float& f) {
    float x = pos * 256.0f - 0.5f;
    float fl = floorf(x);
    f = x - fl;
    i0 = (int) fminf(fmaxf(fl, 0.0f), 255.0f);
    i1 = (int) fminf(fmaxf(fl + 1.0f, 0.0f), 255.0f);
}
// opacity correction for the step length (WindowedSampling.usf:34-35)
__device__ __forceinline__ float step_opacity(float a, float step) {
    a = saturatef(a);
    return 1.0f - det_pow(1.0f - a, step);
}

}  // namespace tbrm
