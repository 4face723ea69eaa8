// sweep_tma.cuh — TMA-staged fused plane-sweep, the fast path of the illumination sweep on sm_100a.
//
// Same schedule as sweep_fused.cuh (one cooperative launch per axis pass, tiles of the buffer plane walk all slices,
// propagated light exchanged through an L2-resident ring with per-tile release/acquire flags), plus:
//   * light-volume bricks (SB slices x tile) are streamed HBM -> SMEM by TMA (cp.async.bulk.tensor.3d) through a
//     3-stage mbarrier pipeline, updated in place in SMEM and written back by a TMA store: the light volume is read
//     once and written once per pass, fully asynchronously, whatever the sweep axis;
//   * the data-volume brick the trilinear taps need (tile + 1 voxel apron, SB + apron slices) rides in the same stage;
//     sweeps along X read an axis-permuted replica of the data volume so that every axis streams x-fastest bricks;
//   * all per-coordinate sampler arithmetic (GetUVW + UVWOffset, tap index, weight, saturate gate, read-buffer UVs) is
//     tabulated once per pass on the host in the exact fp32 order of the shader, so the kernel does no divisions for
//     addressing; a thread owns 2 adjacent pixels and shares their taps;
//   * UNORM8 decode v/255 and the TF-position division use Markstein's correctly rounded 3-op sequence (identical
//     results to IEEE division; tests/test_host_cpu.py proves the decode exhaustively).
// Covers: U8 data, R32F light volume, full-resolution light volume, X % 16 == 0. Everything else takes sweep_fused.cuh.
#pragma once
#include <cuda.h>

#include <vector>

namespace tbrm {

constexpr int kTmaThreads = 256;
constexpr int kTW = 64, kTH = 8;  // tile: 64 x 8 pixels, 2 adjacent pixels per thread
constexpr int kSB = 4;            // slices per pipeline stage
constexpr int kStages = 3;
constexpr int kRingDepth = 16;
constexpr int kHaloPerThread = 2;          // footprint cells owned by other tiles, fetched per thread
constexpr int kFpW = kTW + 4, kFpH = kTH + 4;  // SMEM footprint of the previous slice

struct AxisTab {  // per native coordinate c of one axis (device pointers)
    const float* S;   // GetUVW(c) + UVWOffset
    const float* f;   // trilinear weight
    const int2* meta; // {tap index i0, S == saturate(S)}
};
struct LightTabs {
    AxisTab ax[3];
    const int2* bx;  // per px: {i0, float bits of fx} of the read-buffer bilinear
    const int2* by;  // per py
};

struct TmaParams {
    SweepUniforms U;
    LightTabs A;      // the (added) light
    int ntx, nty;
    float* ring;
    unsigned int* flags;
    int dmin[3], dext[3];   // data-box offset / extent along transposed (p,q,s): box_p0 = tile_p0 + dmin[0], ...
    int ds_q, ds_s;         // SMEM strides (bytes) of the data box along q and s
    int ls_p, ls_q, ls_s;   // SMEM strides (floats) of the light box
    int bmin[2], bext[2];   // footprint offset / extent in the buffer plane
    int data_dims_t[3];     // data dims in transposed (p,q,s) order
    int stage_bytes, light_bytes, data_bytes;
    unsigned int epoch;     // distinguishes the ring tags of successive passes
};

// ---- PTX helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(c2), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace tbrm

#include "sweep_tma_kernel.cuh"

namespace tbrm {

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled) p;
    }
    return fn;
}

static bool make_map3(CUtensorMap* m, CUtensorMapDataType type, size_t elem, void* base, const int dims[3], const int box[3]) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t gd[3] = {(cuuint64_t) dims[0], (cuuint64_t) dims[1], (cuuint64_t) dims[2]};
    cuuint64_t gs[2] = {(cuuint64_t) dims[0] * elem, (cuuint64_t) dims[0] * dims[1] * elem};
    cuuint32_t bd[3] = {(cuuint32_t) box[0], (cuuint32_t) box[1], (cuuint32_t) box[2]};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, type, 3, base, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Per-coordinate sampler tables of one light in one pass, in the exact fp32 order of the shader / oracle
// (this file is compiled with -ffp-contract=off, so host floats round like the device's).
struct HostTabs {
    std::vector<float> S[3], f[3];
    std::vector<int2> meta[3];
    std::vector<int2> bx, by;
    int dmin[3], dmax[3];  // min / max of (i0 - c) per native axis
    int bmin[2], bmax[2];
    bool pairs_ok = true;  // i0(c+1) == i0(c) + 1 everywhere along each axis
};

static void build_tabs(const SweepUniforms& u, const LightPass& L, HostTabs& T) {
    for (int a = 0; a < 3; ++a) {
        const int n = u.ldims[a], nd = u.ddims[a];
        T.S[a].resize(n), T.f[a].resize(n), T.meta[a].resize(n);
        T.dmin[a] = 1 << 30, T.dmax[a] = -(1 << 30);
        for (int c = 0; c < n; ++c) {
            const float s = ((float) c + 0.5f) / (float) n + L.uvw_off[a];  // GetUVW + UVWOffset
            const float x = s * (float) nd - 0.5f;
            float fl = floorf(x);
            const float fr = x - fl;
            fl = fminf(fmaxf(fl, -4.0f), (float) nd + 4.0f);
            const int i0 = (int) fl;
            const float sat = fminf(fmaxf(s, 0.0f), 1.0f);
            T.S[a][c] = s, T.f[a][c] = fr, T.meta[a][c] = make_int2(i0, s == sat ? 1 : 0);
            T.dmin[a] = std::min(T.dmin[a], i0 - c), T.dmax[a] = std::max(T.dmax[a], i0 - c);
            if (c > 0 && T.meta[a][c - 1].x + 1 != i0) T.pairs_ok = false;
        }
    }
    const int tx = u.td[0], ty = u.td[1];
    T.bx.resize(tx), T.by.resize(ty);
    for (int d = 0; d < 2; ++d) {
        const int n = d ? ty : tx;
        std::vector<int2>& out = d ? T.by : T.bx;
        T.bmin[d] = 1 << 30, T.bmax[d] = -(1 << 30);
        for (int c = 0; c < n; ++c) {
            const float uu = ((float) c + 0.5f) / (float) n + L.uv_off[d];
            const float x = uu * (float) n - 0.5f;
            float fl = floorf(x);
            const float fr = x - fl;
            fl = fminf(fmaxf(fl, -4.0f), (float) n + 4.0f);
            const int i0 = (int) fl;
            int bits;
            memcpy(&bits, &fr, 4);
            out[c] = make_int2(i0, bits);
            T.bmin[d] = std::min(T.bmin[d], i0 - c), T.bmax[d] = std::max(T.bmax[d], i0 - c);
            if (d == 0 && c > 0 && out[c - 1].x + 1 != i0) T.pairs_ok = false;
        }
    }
}

// conservative test that the clip weight is exactly 1 for every voxel (then the kernel skips the clip arithmetic)
static bool clip_is_inactive(const SweepUniforms& u, const HostTabs& T) {
    // dist = dot(S - P, D) over the box of all S; weight = clamp(0.5 + 0.577 * |dist| * |D o res| * sign, 0, 1)
    double lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        lo[a] = hi[a] = T.S[a][0];
        for (float s : T.S[a]) lo[a] = std::min(lo[a], (double) s), hi[a] = std::max(hi[a], (double) s);
    }
    double dmin = 0.0;
    for (int a = 0; a < 3; ++a) {
        const double d = u.clip_dir[a];
        const double c0 = (lo[a] - u.clip_center[a]) * d, c1 = (hi[a] - u.clip_center[a]) * d;
        dmin += std::min(c0, c1);
    }
    double scale = 0.0;
    for (int a = 0; a < 3; ++a) scale += (double) u.clip_dir[a] * u.ldims[a] * (double) u.clip_dir[a] * u.ldims[a];
    scale = std::sqrt(scale);
    if (!(dmin > 0.0) || !(scale > 0.0)) return false;
    // need 0.5 + 0.577 * dmin * scale >= 1 with a wide margin for the fp32 evaluation (which cancels S - (S + D*dist))
    return 0.57735026919 * dmin * scale > 4.0 && dmin < 1e30;
}

static cudaError_t upload(tbrm_resources& r, const void* src, size_t bytes, size_t& off, const void** dptr) {
    off = (off + 15) & ~(size_t) 15;
    *dptr = (const char*) r.tables + off;
    cudaError_t e = cudaMemcpyAsync((char*) r.tables + off, src, bytes, cudaMemcpyHostToDevice, r.stream);
    off += bytes;
    return e;
}

template <int AXIS>
static cudaError_t tma_launch(tbrm_resources& r, const CUtensorMap& lm, const CUtensorMap& dm, const TmaParams& P, bool clip, int ntiles,
                              size_t smem) {
    const float4* tf = r.tf;
    void* args[] = {(void*) &lm, (void*) &dm, (void*) &P, (void*) &tf};
    const void* k = clip ? (const void*) sweep_tma_kernel<AXIS, true> : (const void*) sweep_tma_kernel<AXIS, false>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    return cudaLaunchCooperativeKernel(k, dim3(ntiles), dim3(kTmaThreads), args, smem, r.stream);
}

__global__ void permute_yzx_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int X, int Y, int Z) {
    // dst[x][z][y] (y fastest) = src[z][y][x]; 32x32 tile transpose of the (x,y) plane for each z
    __shared__ uint8_t tile[32][33];
    const int z = blockIdx.z, bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int x = bx + threadIdx.x, y = by + j;
        if (x < X && y < Y) tile[j][threadIdx.x] = src[(size_t) x + (size_t) X * ((size_t) y + (size_t) Y * z)];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int y = by + threadIdx.x, x = bx + j;
        if (x < X && y < Y) dst[(size_t) y + (size_t) Y * ((size_t) z + (size_t) Z * x)] = tile[threadIdx.x][j];
    }
}

// the (y,z,x)-ordered replica of the data volume used by sweeps along X; rebuilt lazily after an upload
static cudaError_t ensure_replica(tbrm_resources& r) {
    if (r.data_yzx_valid) return cudaSuccess;
    const size_t bytes = r.data_voxels();
    cudaError_t e;
    if (!r.data_yzx && (e = cudaMalloc(&r.data_yzx, bytes)) != cudaSuccess) return e;
    const dim3 block(32, 8), grid((r.ddims[0] + 31) / 32, (r.ddims[1] + 31) / 32, r.ddims[2]);
    permute_yzx_kernel<<<grid, block, 0, r.stream>>>((const uint8_t*) r.data, (uint8_t*) r.data_yzx, r.ddims[0], r.ddims[1], r.ddims[2]);
    count_launch();
    r.data_yzx_valid = true;
    return cudaGetLastError();
}

cudaError_t sweep_pass_tma(tbrm_resources& r, const SweepUniforms& u, bool change, int* launches, bool* handled) {
    *handled = false;
    if (change || r.data_fmt != TBRM_FMT_G8 || r.light_fmt != TBRM_FMT_R32F || r.half_res) return cudaSuccess;
    const int X = r.ddims[0], Y = r.ddims[1], Z = r.ddims[2];
    if (X % 16 != 0 || (u.axis == 0 && Y % 16 != 0)) return cudaSuccess;  // TMA global strides must be multiples of 16 bytes
    if (((uintptr_t) r.data & 15) || ((uintptr_t) r.light & 15)) return cudaSuccess;
    if (!get_encode()) return cudaSuccess;
    int dev = r.device, sms = 0, coop = 0;
    cudaError_t e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev)) != cudaSuccess) return e;
    if (!coop) return cudaSuccess;

    HostTabs T;
    build_tabs(u, u.a, T);
    if (!T.pairs_ok) return cudaSuccess;
    const int tx = u.td[0], ty = u.td[1];
    TmaParams P;
    memset(&P, 0, sizeof(P));
    P.U = u;
    P.ntx = (tx + kTW - 1) / kTW, P.nty = (ty + kTH - 1) / kTH;
    const int ntiles = P.ntx * P.nty;
    // transposed axis order (p,q,s) in native axes
    const int pa = u.axis == 0 ? 1 : 0, qa = u.axis == 2 ? 1 : 2, sa = u.axis;
    const int nat[3] = {pa, qa, sa};
    for (int t = 0; t < 3; ++t) {
        P.dmin[t] = T.dmin[nat[t]];
        P.data_dims_t[t] = r.ddims[nat[t]];
    }
    // data box extents: tile (+1 tap, + spread of i0 - c), inner extent rounded up to 16 bytes (+4 for the funnel read)
    // TMA needs the box origin 16-byte aligned along the innermost dimension: round the u8 box start down to 16 voxels
    // (tile origins are multiples of 64), and widen the box accordingly.
    const int dmin_p_al = (int) floorf((float) T.dmin[pa] / 16.0f) * 16;
    P.dmin[0] = dmin_p_al;
    const int ext_p = kTW + (T.dmax[pa] - dmin_p_al) + 1, ext_q = kTH + (T.dmax[qa] - T.dmin[qa]) + 1,
              ext_s = kSB + (T.dmax[sa] - T.dmin[sa]) + 1;
    P.dext[0] = (ext_p + 4 + 15) / 16 * 16, P.dext[1] = ext_q, P.dext[2] = ext_s;
    if (P.dext[0] > 256 || P.dext[1] > 256 || P.dext[2] > 256) return cudaSuccess;
    for (int d = 0; d < 2; ++d) P.bmin[d] = T.bmin[d], P.bext[d] = (d ? kTH : kTW) + (T.bmax[d] - T.bmin[d]) + 1;
    if (P.bext[0] > kFpW || P.bext[1] > kFpH) return cudaSuccess;
    if (u.td[2] >= 65000) return cudaSuccess;  // slice index + 1 must fit the 16-bit tag field
    // halo cells (footprint cells of other tiles) must fit kHaloPerThread per thread
    {
        const int ow = std::max(0, std::min(P.bmin[0] + P.bext[0], kTW) - std::max(P.bmin[0], 0));
        const int oh = ow > 0 ? std::max(0, std::min(P.bmin[1] + P.bext[1], kTH) - std::max(P.bmin[1], 0)) : 0;
        if (P.bext[0] * P.bext[1] - ow * oh > kHaloPerThread * kTmaThreads) return cudaSuccess;
    }
    // dependency lists must fit
    {
        const long long nx = (long long) (P.bext[0] + kTW - 1) / kTW + 1, ny = (long long) (P.bext[1] + kTH - 1) / kTH + 1;
        if (nx * ny - 1 > kFusedMaxDeps) return cudaSuccess;
    }
    // tensor maps. Light: native (X,Y,Z) fp32, box = tile x SB along the sweep axis.
    int lbox[3], ldims[3] = {r.ldims[0], r.ldims[1], r.ldims[2]};
    lbox[pa] = kTW, lbox[qa] = kTH, lbox[sa] = kSB;
    CUtensorMap lm, dm;
    if (!make_map3(&lm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, r.light, ldims, lbox)) return cudaSuccess;
    // light SMEM strides: box is stored with native x fastest, then y, then z
    {
        const int str[3] = {1, lbox[0], lbox[0] * lbox[1]};
        P.ls_p = str[pa], P.ls_q = str[qa], P.ls_s = str[sa];
    }
    if (u.axis == 0) {
        if ((e = ensure_replica(r)) != cudaSuccess) return e;
        const int dd[3] = {Y, Z, X}, db[3] = {P.dext[0], P.dext[1], P.dext[2]};
        if (!make_map3(&dm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, r.data_yzx, dd, db)) return cudaSuccess;
        P.ds_q = P.dext[0], P.ds_s = P.dext[0] * P.dext[1];
    } else {
        int dd[3] = {X, Y, Z}, db[3];
        db[pa] = P.dext[0], db[qa] = P.dext[1], db[sa] = P.dext[2];
        if (!make_map3(&dm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, r.data, dd, db)) return cudaSuccess;
        const int str[3] = {1, db[0], db[0] * db[1]};
        P.ds_q = str[qa], P.ds_s = str[sa];
    }
    P.light_bytes = kTW * kTH * kSB * 4;
    P.data_bytes = P.dext[0] * P.dext[1] * P.dext[2];
    P.stage_bytes = (P.light_bytes + P.data_bytes + 16 + 127) / 128 * 128;
    const size_t smem = (size_t) kStages * P.stage_bytes + (2 * kFpW * kFpH + 256) * sizeof(float) + kStages * sizeof(uint64_t) + 16;

    const bool clip = !clip_is_inactive(u, T);
    const void* kern = nullptr;
    switch (u.axis) {
        case 0: kern = clip ? (const void*) sweep_tma_kernel<0, true> : (const void*) sweep_tma_kernel<0, false>; break;
        case 1: kern = clip ? (const void*) sweep_tma_kernel<1, true> : (const void*) sweep_tma_kernel<1, false>; break;
        default: kern = clip ? (const void*) sweep_tma_kernel<2, true> : (const void*) sweep_tma_kernel<2, false>; break;
    }
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) != cudaSuccess) {
        cudaGetLastError();
        return cudaSuccess;
    }
    int per_sm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTmaThreads, smem)) != cudaSuccess) return e;
    if ((long long) sms * per_sm < ntiles) return cudaSuccess;  // the plane does not fit one co-resident wave

    // scratch: ring, flags, tables
    const size_t ring_bytes = (size_t) kRingDepth * tx * ty * sizeof(unsigned long long);
    if (r.ring_bytes < ring_bytes) {
        if (r.ring) cudaStreamSynchronize(r.stream), cudaFree(r.ring);
        r.ring = nullptr, r.ring_bytes = 0;
        if ((e = cudaMalloc(&r.ring, ring_bytes)) != cudaSuccess) return e;
        r.ring_bytes = ring_bytes;
        r.ring_epoch = 0;
    }
    // ring tags are (epoch << 16 | slice + 1): a fresh epoch per pass means cells left by earlier passes never match.
    // Clear the ring when it is new, when the generic fused kernel (plain floats) used it since, or when the epoch wraps.
    if (r.ring_epoch == 0 || r.ring_epoch >= 0xffffu) {
        if ((e = cudaMemsetAsync(r.ring, 0, r.ring_bytes, r.stream)) != cudaSuccess) return e;
        r.ring_epoch = 0;
    }
    P.epoch = ++r.ring_epoch;
    if (r.flags_count < (size_t) ntiles * kFlagStride) {
        if (r.flags) cudaStreamSynchronize(r.stream), cudaFree(r.flags);
        r.flags = nullptr, r.flags_count = 0;
        if ((e = cudaMalloc((void**) &r.flags, (size_t) ntiles * kFlagStride * sizeof(unsigned int))) != cudaSuccess) return e;
        r.flags_count = (size_t) ntiles * kFlagStride;
    }
    size_t need = 256;
    for (int a = 0; a < 3; ++a) need += (size_t) u.ldims[a] * 16 + 48;
    need += (size_t) (tx + ty) * 8 + 32;
    if (r.tables_bytes < need) {
        if (r.tables) cudaStreamSynchronize(r.stream), cudaFree(r.tables);
        r.tables = nullptr, r.tables_bytes = 0;
        if ((e = cudaMalloc(&r.tables, need)) != cudaSuccess) return e;
        r.tables_bytes = need;
    }
    // the previous pass may still be reading the tables: stream order makes the async copies wait for it
    size_t off = 0;
    for (int a = 0; a < 3; ++a) {
        if ((e = upload(r, T.S[a].data(), T.S[a].size() * 4, off, (const void**) &P.A.ax[a].S)) != cudaSuccess) return e;
        if ((e = upload(r, T.f[a].data(), T.f[a].size() * 4, off, (const void**) &P.A.ax[a].f)) != cudaSuccess) return e;
        if ((e = upload(r, T.meta[a].data(), T.meta[a].size() * 8, off, (const void**) &P.A.ax[a].meta)) != cudaSuccess) return e;
    }
    if ((e = upload(r, T.bx.data(), T.bx.size() * 8, off, (const void**) &P.A.bx)) != cudaSuccess) return e;
    if ((e = upload(r, T.by.data(), T.by.size() * 8, off, (const void**) &P.A.by)) != cudaSuccess) return e;
    // pageable-memory async copies return once the source has been staged, so T may go out of scope after this call
    P.ring = (float*) r.ring;
    P.flags = r.flags;
    if ((e = cudaMemsetAsync(r.flags, 0, (size_t) ntiles * kFlagStride * sizeof(unsigned int), r.stream)) != cudaSuccess) return e;
    switch (u.axis) {
        case 0: e = tma_launch<0>(r, lm, dm, P, clip, ntiles, smem); break;
        case 1: e = tma_launch<1>(r, lm, dm, P, clip, ntiles, smem); break;
        default: e = tma_launch<2>(r, lm, dm, P, clip, ntiles, smem); break;
    }
    if (e != cudaSuccess) return e;
    count_launch();
    *launches += 1;
    *handled = true;
    return cudaSuccess;
}

}  // namespace tbrm
