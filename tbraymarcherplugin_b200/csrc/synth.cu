// synth.cu — deterministic synthetic volumes of SURVEY.md §8(d), generated on the device in fp64 so that the numpy
// twin (tbraymarcherplugin_b200/synth.py) reproduces them bit-for-bit. Benchmark / test harness input, not the
// reference's code.
#include "tbrm_internal.hpp"

namespace tbrm {

__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ double lattice(int x, int y, int z, uint32_t seed) {
    const uint32_t h = lowbias32((uint32_t) x + 374761393u * (uint32_t) y + 668265263u * (uint32_t) z + seed);
    return (double) h / 4294967296.0;
}

__global__ void synth_kernel(int kind, int X, int Y, int Z, uint32_t seed, uint8_t* __restrict__ out) {
    const size_t n = (size_t) X * Y * Z;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const int x = (int) (i % X), y = (int) ((i / X) % Y), z = (int) (i / ((size_t) X * Y));
        const double u = ((double) x + 0.5) / (double) X, v = ((double) y + 0.5) / (double) Y, w = ((double) z + 0.5) / (double) Z;
        double val;
        if (kind == TBRM_SYNTH_SPHERE) {
            const double du = u - 0.5, dv = v - 0.5, dw = w - 0.5;
            const double d = sqrt(du * du + dv * dv + dw * dw);
            val = fmax(0.0, 1.0 - d / 0.4);
        } else {
            double sum = 0.0, amp = 1.0, norm = 0.0;
            for (int o = 0; o < 4; ++o) {
                const double cells = (double) (4 << o);
                const double px = u * cells, py = v * cells, pz = w * cells;
                const double fx0 = floor(px), fy0 = floor(py), fz0 = floor(pz);
                const int ix = (int) fx0, iy = (int) fy0, iz = (int) fz0;
                double fx = px - fx0, fy = py - fy0, fz = pz - fz0;
                fx = fx * fx * (3.0 - 2.0 * fx);
                fy = fy * fy * (3.0 - 2.0 * fy);
                fz = fz * fz * (3.0 - 2.0 * fz);
                const uint32_t s = seed + (uint32_t) o * 0x9E3779B9u;
                const double c000 = lattice(ix, iy, iz, s), c100 = lattice(ix + 1, iy, iz, s);
                const double c010 = lattice(ix, iy + 1, iz, s), c110 = lattice(ix + 1, iy + 1, iz, s);
                const double c001 = lattice(ix, iy, iz + 1, s), c101 = lattice(ix + 1, iy, iz + 1, s);
                const double c011 = lattice(ix, iy + 1, iz + 1, s), c111 = lattice(ix + 1, iy + 1, iz + 1, s);
                const double x00 = c000 + fx * (c100 - c000), x10 = c010 + fx * (c110 - c010);
                const double x01 = c001 + fx * (c101 - c001), x11 = c011 + fx * (c111 - c011);
                const double y0 = x00 + fy * (x10 - x00), y1 = x01 + fy * (x11 - x01);
                sum = sum + amp * (y0 + fz * (y1 - y0));
                norm = norm + amp;
                amp = amp * 0.5;
            }
            const double noise = sum / norm;
            const double eu = (u - 0.5) / 0.45, ev = (v - 0.5) / 0.40, ew = (w - 0.5) / 0.48;
            const double e = sqrt(eu * eu + ev * ev + ew * ew);
            const double mask = fmin(1.0, fmax(0.0, (1.0 - e) / 0.02));
            val = noise * mask;
        }
        out[i] = (uint8_t) floor(255.0 * val + 0.5);
    }
}

cudaError_t synth_volume_u8(cudaStream_t stream, int kind, const int32_t dims[3], uint32_t seed, uint8_t* d_out) {
    synth_kernel<<<148 * 8, 256, 0, stream>>>(kind, dims[0], dims[1], dims[2], seed, d_out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace tbrm
