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

#include <cstdio>
#include <cstdlib>
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

// What a launch that covers only part of a pass needs (see the SLAB notes on sweep_tma_kernel)
struct SlabParams {
    int tile_row0, tile_rows;        // tile rows of the buffer plane this launch walks
    int q_lo, q_hi;                  // buffer rows it owns
    int k_begin, k_end;              // slices (in sweep order) it walks
    int reach_lo, reach_hi;          // rows below q_lo / at or above q_hi a band's footprints read
    const unsigned long long* inbox; // [slice][reach_lo + reach_hi row slots][tx] LL cells: rows q_lo-reach_lo .. q_lo-1, q_hi .. q_hi+reach_hi-1
    unsigned long long* out_lo;      // inbox of the band below (rows < q_lo): local, or the neighbour GPU's through NVLink; or null
    unsigned long long* out_hi;      // inbox of the band above
    const unsigned long long* zin;   // plane of LL cells holding the upstream slab's last slice (k_begin > 0), or null
    unsigned long long* zout;        // the downstream slab's zin (k_end < slices), or null
    unsigned int* error;             // device word set when an exchange timed out
    unsigned long long timeout_ns;
};

// what a launch does with the light it propagates
enum { kModeAdd = 0,      // LightVolume += light * sign where |light| > 1e-3 (AddDirLight)
       kModeStore = 1,    // scratch volume = light (the removed light of a ChangeDirLight; light_map is the scratch volume's map)
       kModeCombine = 2 };// LightVolume += light - scratch where |light - scratch| > 1e-3 (the added light of a ChangeDirLight)

// tensor maps of the OTHER ranks' light volumes (push-gather): the finished light brick is stored into them as well
constexpr int kMaxPush = 15;
struct alignas(64) PushMaps {
    CUtensorMap m[kMaxPush];
};

struct TmaParams {
    SweepUniforms U;
    SlabParams S;
    int mode;
    int data_off;           // offset of the data brick inside a stage (after the light brick and, when combining, the scratch brick)
    LightTabs A;      // the (added) light
    int ntx, nty;
    float* ring;
    unsigned int* flags;
    int dmin[3], dext[3];   // data-box offset / extent along transposed (p,q,s): box_p0 = dk[0] * tile_p0 + dmin[0], ...
    int dk[3];              // data voxels per light voxel along (p,q,s): 1, or 2 for a half-resolution light volume (one-pixel form only)
    int ds_q, ds_s;         // SMEM strides (bytes) of the data box along q and s
    int ls_p, ls_q, ls_s;   // SMEM strides (elements) of the light box
    int ss_p, ss_q, ss_s;   // ... of the scratch box of a combining launch (differs from the light box's for a G8 sweep along X)
    int bmin[2], bext[2];   // footprint offset / extent in the buffer plane
    int data_dims_t[3];     // data dims in transposed (p,q,s) order
    int stage_bytes, light_bytes, data_bytes;
    unsigned int epoch;     // distinguishes the ring tags of successive passes
    unsigned int cut_lo_mode, cut_lo_add;  // exact empty-space skip on tap bytes (see window_cut_byte); mode 0 = off
    int exp_flags;          // experiment switches (tbrm_options.reserved[0])
    // second generation (sweep_split.cuh)
    const float* tvol;            // T = 1 - occlusion bricks written by occlusion_kernel: [tile][native block][slot][row][p]
    const unsigned char* tones;   // per brick: every T is exactly 1 (the brick was not written)
    const float* scratch;   // the removed light's propagated light (kModeCombine), light-volume layout
    long long* dbg;         // TBRM_CHAIN_TIMERS builds: per-warp section timers of the chain kernel
    unsigned int* error;    // device word set when a wait on another tile / launch timed out
    unsigned long long timeout_ns;
    unsigned int poll_limit;  // unsharded launches: polls after which a wait on another tile gives up (timeout_ns / 256 ns)
    int n_push;               // peers the light bricks are pushed to (0: none)
    int q_org;                // first generation: buffer row of tile row 0 (0, or negative where a slab does not start on a multiple of the tile height)
};

// ---- PTX helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among `nthreads` threads of the block (whole warps) on hardware barrier `id` (0 is __syncthreads' own)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
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
// the same wait where the phase is expected to complete much later (a producer that is a block ahead): the hardware suspends the
// thread for up to the hinted time instead of returning to the polling loop
__device__ __forceinline__ void mbar_wait_long(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned, size a multiple of 16), completing on an mbarrier like the tensor loads
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
// Ampere-style asynchronous copy of 16 bytes global -> shared through L2 only (.cg), tracked per thread in commit groups
__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
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
#include "sweep_split.cuh"

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
    int dk[3];             // data voxels per light voxel along each native axis (1; 2 for a half-resolution light volume)
    int dmin[3], dmax[3];  // min / max of (i0 - dk * c) per native axis
    int bmin[2], bmax[2];
    bool pairs_ok = true;  // i0(c+1) == i0(c) + 1 everywhere along each axis (the two-pixel forms share tap columns between neighbours)
    bool weights_lt_one = true;  // every trilinear weight is < 1 (needed by the exact empty-space skip)
};

static void build_tabs(const SweepUniforms& u, const LightPass& L, HostTabs& T) {
    for (int a = 0; a < 3; ++a) {
        const int n = u.ldims[a], nd = u.ddims[a];
        T.S[a].resize(n), T.f[a].resize(n), T.meta[a].resize(n);
        T.dmin[a] = 1 << 30, T.dmax[a] = -(1 << 30);
        const int dk = T.dk[a] = (n > 0 && nd % n == 0) ? nd / n : 1;
        for (int c = 0; c < n; ++c) {
            const float s = ((float) c + 0.5f) / (float) n + L.uvw_off[a];  // GetUVW + UVWOffset
            const float x = s * (float) nd - 0.5f;
            float fl = floorf(x);
            const float fr = x - fl;
            fl = fminf(fmaxf(fl, -4.0f), (float) nd + 4.0f);
            const int i0 = (int) fl;
            const float sat = fminf(fmaxf(s, 0.0f), 1.0f);
            T.S[a][c] = s, T.f[a][c] = fr, T.meta[a][c] = make_int2(i0, s == sat ? 1 : 0);
            if (!(fr >= 0.0f && fr < 1.0f)) T.weights_lt_one = false;
            T.dmin[a] = std::min(T.dmin[a], i0 - dk * c), T.dmax[a] = std::max(T.dmax[a], i0 - dk * c);
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

// Largest byte T the low cut-off rejects, evaluated with the shader's fp32 arithmetic (GetTransferFuncPosition +
// cut-off, WindowedSampling.usf:14-29). A trilinear value lies between its smallest and largest tap (each lerp
// fma(t, b-a, a) with 0 <= t < 1 stays inside [a,b]) and the window position is monotone in the value, so taps that
// are all <= T imply a rejected sample. mode 0: no skip; 1: T < 128; 2: T >= 128; `add` is the SWAR addend.
static void window_cut_byte(const Windowing& w, bool allow, unsigned int& mode, unsigned int& add, int* cut = nullptr) {
    int lo = -1;
    bool ok = allow && w.low > 0.0f;
    for (int b = 0; b < 256 && ok; ++b) {
        const float v = (float) b / 255.0f;
        const float pos = (v - w.center + (w.width / 2.0f)) / w.width;
        if (pos < 0.0f) {
            if (lo != b - 1) ok = false;  // the rejected bytes must be a prefix of the byte range
            lo = b;
        }
    }
    mode = 0, add = 0;
    if (cut) *cut = (ok && lo >= 0) ? lo : -1;
    if (!ok || lo < 0) return;
    if (lo < 128)
        mode = 1, add = (unsigned) (127 - lo) * 0x010101u;
    else
        mode = 2, add = (unsigned) (255 - lo) * 0x010101u;
}

// the per-pass sampler tables travel as ONE host-to-device copy: they are packed into a host staging vector first (eleven small copies per
// pass were ~0.1 ms of host time — nothing on one GPU, where the kernels run for milliseconds, but a sixth of the step on eight)
static void pack(std::vector<unsigned char>& staging, const tbrm_resources& r, const void* src, size_t bytes, const void** dptr) {
    const size_t off = (staging.size() + 15) & ~(size_t) 15;
    staging.resize(off + bytes);
    memcpy(staging.data() + off, src, bytes);
    *dptr = (const char*) r.tables + off;
}

static const void* tma_kernel_l8(int axis, bool clip, bool slab, int px, int th = 8) {  // G8 light volume
#define TBRM_K(A, PX) (slab ? (clip ? (const void*) sweep_tma_kernel<A, true, true, PX, true> : (const void*) sweep_tma_kernel<A, false, true, PX, true>) \
                            : (clip ? (const void*) sweep_tma_kernel<A, true, false, PX, true> : (const void*) sweep_tma_kernel<A, false, false, PX, true>))
#define TBRM_K7(A) (clip ? (const void*) sweep_tma_kernel<A, true, false, 2, true, 7> : (const void*) sweep_tma_kernel<A, false, false, 2, true, 7>)
    if (th == 7 && px == 2 && !slab) return axis == 0 ? TBRM_K7(0) : (axis == 1 ? TBRM_K7(1) : TBRM_K7(2));  // 7-row tiles: two-pixel form only
    if (px == 1) return axis == 0 ? TBRM_K(0, 1) : (axis == 1 ? TBRM_K(1, 1) : TBRM_K(2, 1));
    return axis == 0 ? TBRM_K(0, 2) : (axis == 1 ? TBRM_K(1, 2) : TBRM_K(2, 2));
#undef TBRM_K
#undef TBRM_K7
}

static const void* tma_kernel(int axis, bool clip, bool slab, int px, int th = 8) {
#define TBRM_K(A, PX) (slab ? (clip ? (const void*) sweep_tma_kernel<A, true, true, PX> : (const void*) sweep_tma_kernel<A, false, true, PX>) \
                            : (clip ? (const void*) sweep_tma_kernel<A, true, false, PX> : (const void*) sweep_tma_kernel<A, false, false, PX>))
#define TBRM_K7(A, PX) (slab ? (clip ? (const void*) sweep_tma_kernel<A, true, true, PX, false, 7> : (const void*) sweep_tma_kernel<A, false, true, PX, false, 7>) \
                             : (clip ? (const void*) sweep_tma_kernel<A, true, false, PX, false, 7> : (const void*) sweep_tma_kernel<A, false, false, PX, false, 7>))
    if (th == 7) {  // 7-row tiles
        if (px == 1) return axis == 0 ? TBRM_K7(0, 1) : (axis == 1 ? TBRM_K7(1, 1) : TBRM_K7(2, 1));
        return axis == 0 ? TBRM_K7(0, 2) : (axis == 1 ? TBRM_K7(1, 2) : TBRM_K7(2, 2));
    }
    if (px == 1) return axis == 0 ? TBRM_K(0, 1) : (axis == 1 ? TBRM_K(1, 1) : TBRM_K(2, 1));
    return axis == 0 ? TBRM_K(0, 2) : (axis == 1 ? TBRM_K(1, 2) : TBRM_K(2, 2));
#undef TBRM_K
#undef TBRM_K7
}

static const void* chain_kernel(int axis, bool slab, int px) {
#define TBRM_K(A, PX) (slab ? (const void*) sweep_chain_kernel<A, true, PX> : (const void*) sweep_chain_kernel<A, false, PX>)
    if (px == 1) return axis == 0 ? TBRM_K(0, 1) : (axis == 1 ? TBRM_K(1, 1) : TBRM_K(2, 1));
    return axis == 0 ? TBRM_K(0, 2) : (axis == 1 ? TBRM_K(1, 2) : TBRM_K(2, 2));
#undef TBRM_K
}
typedef void (*occlusion_fn)(const OccParams, const float4*);
static occlusion_fn occlusion_kernel_of(int axis, bool clip, int px) {
#define TBRM_K(A, PX) (clip ? (occlusion_fn) occlusion_kernel<A, true, PX> : (occlusion_fn) occlusion_kernel<A, false, PX>)
    if (px == 1) return axis == 0 ? TBRM_K(0, 1) : (axis == 1 ? TBRM_K(1, 1) : TBRM_K(2, 1));
    return axis == 0 ? TBRM_K(0, 2) : (axis == 1 ? TBRM_K(1, 2) : TBRM_K(2, 2));
#undef TBRM_K
}

static cudaError_t tma_launch(tbrm_resources& r, const void* kern, int threads, const CUtensorMap& lm, const CUtensorMap& dm, const CUtensorMap& sm,
                              const PushMaps& pm, const TmaParams& P, int ntiles, size_t smem) {
    const float4* tf = r.tf;
    void* args[] = {(void*) &lm, (void*) &dm, (void*) &sm, (void*) &pm, (void*) &P, (void*) &tf};
    return cudaLaunchCooperativeKernel(kern, dim3(ntiles), dim3(threads), args, smem, r.stream);
}

// ---- slab exchange arena: header (acks, error word) + 4 regions x (inbox + hand-off plane) of LL cells ----------------
// A launch reads the region (pass parity, band parity): neighbouring bands of one pass use different regions, and a region
// comes up again two passes later — which is what the ack protocol below protects (a neighbour that runs ahead writes
// pass seq + 1 into the other pair of regions while this GPU still reads pass seq).
static int arena_region(unsigned int seq, int band) { return (int) ((seq & 1u) << 1) | (band & 1); }
constexpr size_t kArenaHeader = 256;
constexpr int kInboxSlots = 8;  // rows of a neighbouring band a launch may read (reach_lo + reach_hi)
static size_t arena_inbox_cells(const int32_t d[3]) {
    const size_t a0 = (size_t) d[0] * d[1], a2 = (size_t) d[2] * d[0];  // axis 0: ns = X, tx = Y; axis 1: ns = Y, tx = X; axis 2
    return std::max(a0, a2) * kInboxSlots;
}
static size_t arena_region_bytes(const int32_t d[3]) { return (arena_inbox_cells(d) + (size_t) d[0] * d[1]) * sizeof(unsigned long long); }
size_t slab_arena_bytes(const int32_t ldims[3]) { return kArenaHeader + 4 * arena_region_bytes(ldims); }
static unsigned long long* arena_inbox(void* arena, const int32_t d[3], int region) {
    return (unsigned long long*) ((char*) arena + kArenaHeader + (size_t) region * arena_region_bytes(d));
}
static unsigned long long* arena_zplane(void* arena, const int32_t d[3], int region) { return arena_inbox(arena, d, region) + arena_inbox_cells(d); }
static unsigned int* arena_word(void* arena, int i) { return (unsigned int*) arena + i; }  // 0: ack from lo, 1: ack from hi, 2: error

cudaError_t slab_ensure_arena(tbrm_resources& r) {
    if (r.arena) return cudaSuccess;
    const size_t bytes = slab_arena_bytes(r.ldims);
    cudaError_t e = cudaMalloc(&r.arena, bytes);
    if (e != cudaSuccess) return e;
    r.arena_bytes = bytes;
    return cudaMemsetAsync(r.arena, 0, bytes, r.stream);
}

// flow control between the passes of neighbouring slabs: a region of a neighbour's arena is rewritten every second pass, so
// before exporting pass `seq` the neighbours must have finished pass seq - 2 (they post the last pass they completed)
__global__ void slab_wait_kernel(const unsigned int* ack_lo, const unsigned int* ack_hi, unsigned int need, unsigned int* error,
                                 unsigned long long timeout_ns) {
    const unsigned long long t0 = global_timer_ns();
    for (int i = 0; i < 2; ++i) {
        const unsigned int* a = i ? ack_hi : ack_lo;
        if (!a) continue;
        unsigned int v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
            if (v < need && global_timer_ns() - t0 > timeout_ns) {
                atomicExch(error, 1u);
                return;
            }
        } while (v < need);
    }
}
__global__ void slab_signal_kernel(unsigned int* to_lo, unsigned int* to_hi, unsigned int seq) {
    __threadfence_system();
    if (to_lo) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(to_lo), "r"(seq) : "memory");
    if (to_hi) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(to_hi), "r"(seq) : "memory");
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

// The same permutation for X % 16 == 0 and Y % 16 == 0 with 16-byte global accesses on both sides (the kernel above moves single bytes):
// a 256-thread CTA transposes a 64 (x) x 64 (y) tile of one z plane through shared memory. In: one 16-byte load per thread, rows of the
// tile stored as words (pitch 17: a warp's column reads below fall into one row, no bank conflict). Out: a thread gathers the 16 bytes
// y0 .. y0 + 15 of one column x and writes them with one 16-byte store.
__global__ void __launch_bounds__(256) permute_yzx_vec_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int X, int Y, int Z) {
    __shared__ uint32_t tile[64][17];
    const int t = threadIdx.x, z = blockIdx.z, bx = blockIdx.x * 64, by = blockIdx.y * 64;
    {
        const int row = t >> 2, vec = t & 3;
        const int x = bx + 16 * vec, y = by + row;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (x < X && y < Y) v = __ldg(reinterpret_cast<const uint4*>(src + (size_t) x + (size_t) X * ((size_t) y + (size_t) Y * z)));
        tile[row][4 * vec + 0] = v.x, tile[row][4 * vec + 1] = v.y, tile[row][4 * vec + 2] = v.z, tile[row][4 * vec + 3] = v.w;
    }
    __syncthreads();
    const int col = t & 63, seg = t >> 6;  // column x of the tile, 16 consecutive y
    const int x = bx + col, y = by + 16 * seg;
    if (x >= X || y >= Y) return;
    const int word = col >> 2, shift = 8 * (col & 3);
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc |= ((tile[16 * seg + 4 * q + i][word] >> shift) & 0xffu) << (8 * i);
        o[q] = acc;
    }
    *reinterpret_cast<uint4*>(dst + (size_t) y + (size_t) Y * ((size_t) z + (size_t) Z * x)) = make_uint4(o[0], o[1], o[2], o[3]);
}

// dst[i1 + D1 (i2 + D2 i0)] = src[i0 + D0 (i1 + D1 i2)]: the byte transpose behind the (y,z,x) replicas. Applied three times it is the identity,
// so two more applications (with the dimensions rotated) undo one.
static cudaError_t permute_bytes(tbrm_resources& r, const uint8_t* src, uint8_t* dst, int D0, int D1, int D2) {
    if ((D0 & 15) == 0 && (D1 & 15) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        const dim3 grid((D0 + 63) / 64, (D1 + 63) / 64, D2);
        permute_yzx_vec_kernel<<<grid, 256, 0, r.stream>>>(src, dst, D0, D1, D2);
    } else {
        const dim3 block(32, 8), grid((D0 + 31) / 32, (D1 + 31) / 32, D2);
        permute_yzx_kernel<<<grid, block, 0, r.stream>>>(src, dst, D0, D1, D2);
    }
    count_launch();
    return cudaGetLastError();
}

// the (y,z,x)-ordered replica of the data volume used by sweeps along X; rebuilt lazily after an upload
static cudaError_t ensure_replica(tbrm_resources& r) {
    if (r.data_yzx_valid) return cudaSuccess;
    const size_t bytes = r.data_voxels();
    cudaError_t e;
    if (!r.data_yzx && (e = cudaMalloc(&r.data_yzx, bytes)) != cudaSuccess) return e;
    if ((r.ddims[0] & 15) == 0 && (r.ddims[1] & 15) == 0 && ((reinterpret_cast<uintptr_t>(r.data) | reinterpret_cast<uintptr_t>(r.data_yzx)) & 15) == 0) {
        const dim3 grid((r.ddims[0] + 63) / 64, (r.ddims[1] + 63) / 64, r.ddims[2]);
        permute_yzx_vec_kernel<<<grid, 256, 0, r.stream>>>((const uint8_t*) r.data, (uint8_t*) r.data_yzx, r.ddims[0], r.ddims[1], r.ddims[2]);
    } else {
        const dim3 block(32, 8), grid((r.ddims[0] + 31) / 32, (r.ddims[1] + 31) / 32, r.ddims[2]);
        permute_yzx_kernel<<<grid, block, 0, r.stream>>>((const uint8_t*) r.data, (uint8_t*) r.data_yzx, r.ddims[0], r.ddims[1], r.ddims[2]);
    }
    count_launch();
    const cudaError_t le = cudaGetLastError();
    r.data_yzx_valid = le == cudaSuccess;
    return le;
}

cudaError_t build_replica_for_tests(tbrm_resources& r) { return ensure_replica(r); }

// the pass is left to another sweep implementation; remember why (shown when a sharded volume has no alternative)
static cudaError_t not_handled(const char* why) {
    set_last_error(std::string("TMA-staged sweep not applicable: ") + why);
    static const bool verbose = getenv("TBRM_DEBUG") != nullptr;
    if (verbose) fprintf(stderr, "[tbrm] TMA-staged sweep not applicable: %s\n", why);
    return cudaSuccess;
}

int slab_pass_order(const SweepUniforms& u) {
    if (u.axis == 2) return u.dirn > 0 ? 1 : -1;  // the slabs of a sweep along Z form a chain in sweep order
    HostTabs T;
    build_tabs(u, u.a, T);
    const int reach_lo = std::max(0, -T.bmin[1]), reach_hi = std::max(0, T.bmax[1] + 1);
    if (reach_lo > 0 && reach_hi > 0) return 2;
    return reach_hi > 0 ? -1 : 1;
}

static cudaError_t sweep_pass_tma_mode(tbrm_resources& r, const SweepUniforms& u, int mode, int* launches, bool* handled);

// resident blocks per SM of a kernel at a dynamic shared-memory size; the attribute and the query are made once per (kernel, size, device).
// *fits = false when the size cannot be set for the kernel.
static cudaError_t blocks_per_sm(const void* kern, int threads, size_t smem, int dev, int* per_sm, bool* fits) {
    struct Known {
        const void* kern;
        size_t smem;
        int dev, per_sm;
    };
    static thread_local std::vector<Known> known;
    *fits = true;
    for (const Known& kn : known)
        if (kn.kern == kern && kn.smem == smem && kn.dev == dev) {
            *per_sm = kn.per_sm;
            return cudaSuccess;
        }
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess) {
        cudaGetLastError();
        *fits = false, *per_sm = 0;
        return cudaSuccess;
    }
    int occ = 0;
    const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    if (e != cudaSuccess) return e;
    // another size may have been set for this kernel in between: drop stale entries of the kernel
    known.erase(std::remove_if(known.begin(), known.end(), [&](const Known& kn) { return kn.kern == kern && kn.dev == dev; }), known.end());
    known.push_back({kern, smem, dev, occ});
    *per_sm = occ;
    return cudaSuccess;
}

// One axis pass. ChangeDirLight (LightingShaders.cpp:168-326) runs as two launches that share the exchange machinery of the
// Add sweep: the removed light's sweep leaves its propagated light in a scratch volume, the added light's sweep combines
// LightVolume += added - removed under the reference's threshold on the difference — the same values, voxel by voxel, as
// the fused shader.
cudaError_t sweep_pass_tma(tbrm_resources& r, const SweepUniforms& u, bool change, int* launches, bool* handled) {
    if (!change) return sweep_pass_tma_mode(r, u, kModeAdd, launches, handled);
    *handled = false;
    cudaError_t e;
    if (!r.change_scratch && (e = cudaMalloc(&r.change_scratch, r.light_voxels() * sizeof(float))) != cudaSuccess) return e;
    SweepUniforms ur = u;
    ur.a = u.r;  // the removed light drives the first launch
    bool h1 = false, h2 = false;
    const unsigned int seq0 = r.pass_seq;
    if ((e = sweep_pass_tma_mode(r, ur, kModeStore, launches, &h1)) != cudaSuccess) return e;
    if (!h1) return cudaSuccess;
    if ((e = sweep_pass_tma_mode(r, u, kModeCombine, launches, &h2)) != cudaSuccess) return e;
    if (!h2) {
        // the first launch only wrote the scratch volume, so another implementation can still take the whole pass; a sharded
        // volume has none (and its neighbours have consumed a sequence number): report it
        if (r.slab.nranks > 1) {
            set_last_error("ChangeDirLight: the added light's sweep is not supported by the TMA-staged sweep although the removed light's is");
            return cudaErrorNotSupported;
        }
        (void) seq0;
        return cudaSuccess;
    }
    *handled = true;
    return cudaSuccess;
}

static cudaError_t sweep_pass_tma_mode(tbrm_resources& r, const SweepUniforms& u, int mode, int* launches, bool* handled) {
    *handled = false;
    if (r.data_fmt != TBRM_FMT_G8) return not_handled("r.data_fmt != TBRM_FMT_G8");
    // a G8 light volume (the reference's default format): byte bricks — AddDirLight, sweeps along Y / Z, unsharded (a byte brick of 4 slices
    // along X has 4-byte rows, below TMA's 16-byte minimum; ChangeDirLight keeps its removed light in an R32F scratch volume)
    const bool l8 = r.light_fmt == TBRM_FMT_G8;
    // the brick a G8 pass loads, updates and stores is bytes of the light volume — except for the removed light of a ChangeDirLight, whose
    // propagated light goes to the R32F scratch volume (its forwarded values are still quantised like the G8 propagation buffers)
    const bool byte_brick = l8 && mode != kModeStore;
    const int X = r.ddims[0], Y = r.ddims[1], Z = r.ddims[2];
    if (X % 16 != 0 || (u.axis == 0 && Y % 16 != 0)) return not_handled("X % 16 != 0 || (u.axis == 0 && Y % 16 != 0)");  // TMA global strides must be multiples of 16 bytes
    if (((uintptr_t) r.data & 15) || ((uintptr_t) r.light & 15)) return not_handled("((uintptr_t) r.data & 15) || ((uintptr_t) r.light & 15)");
    if (!get_encode()) return not_handled("!get_encode()");
    int dev = r.device, sms = 0, coop = 0;
    cudaError_t e;
    {   // device attributes are queried once per device (host time per pass matters once the kernels are short: 8-GPU slabs)
        static thread_local int cached_dev = -1, cached_sms = 0, cached_coop = 0;
        if (cached_dev != dev) {
            if ((e = cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
            if ((e = cudaDeviceGetAttribute(&cached_coop, cudaDevAttrCooperativeLaunch, dev)) != cudaSuccess) return e;
            cached_dev = dev;
        }
        sms = cached_sms, coop = cached_coop;
    }
    if (!coop) return not_handled("!coop");

    HostTabs T;
    build_tabs(u, u.a, T);
    // the two-pixel forms (and the second generation) share tap columns between neighbouring pixels: unit stride from pixel to data voxel.
    // A half-resolution light volume (two data voxels per light voxel) or irregular tap tables take the one-pixel form.
    const bool unit_stride = T.pairs_ok && T.dk[0] == 1 && T.dk[1] == 1 && T.dk[2] == 1;
    const int tx = u.td[0], ty = u.td[1];
    TmaParams P;
    memset(&P, 0, sizeof(P));
    P.U = u;
    // pixels per thread: two (tile 64 x 8), or one (tile 32 x 8) for launches that cannot fill the SMs — there a pass costs slices x one
    // tile's per-slice chain, which one pixel per thread shortens. Automatic by default (measured on a B200 in round 2: 256^3 reset of two
    // lights 1.52 -> 1.31 ms, the GPU suite green in both forms); TBRM_SWEEP_PX=1 / 2 (or bits 4-5 of reserved[0]) force a form.
    // kernel generation: 1 = sweep_tma_kernel.cuh (the default: still the fastest measured), 2 = occlusion kernel + propagation-chain kernel
    // (sweep_split.cuh; TBRM_SWEEP_GEN=2 or bit 6 of reserved[0]). Round 2 measured generation 2 at 4.9 ms against 4.67 ms for the cfg2
    // reset (bit-identical results): its chain kernel runs at ~2300 cycles per slice, half of it waiting.
    static const int env_gen = [] {
        const char* e = getenv("TBRM_SWEEP_GEN");
        return e ? atoi(e) : 1;
    }();
    // (the chain kernel of the second generation walks whole blocks of kSB slices: other slice counts take the first generation)
    const bool ws = (env_gen == 2 || (r.options.reserved[0] & 64)) && u.td[2] % kSB == 0 && r.light_fmt == TBRM_FMT_R32F && unit_stride;
    int px = 2;
    {
        static const int env_px = [] {  // 1 / 2 forced, 3 automatic (default)
            const char* e = getenv("TBRM_SWEEP_PX");
            return !e ? 3 : (e[0] == 'a' ? 3 : atoi(e));
        }();
        const int opt = (r.options.reserved[0] >> 4) & 3;
        const int want = opt ? opt : env_px;
        if (want == 1 || want == 2) px = want;
        if (want == 3) {
            const int rows = (r.slab.nranks > 1 && u.axis != 2) ? r.slab.z_end - r.slab.z_begin : ty;  // buffer rows this GPU sweeps
            const long long tiles64 = (long long) ((tx + 63) / 64) * ((rows + kTH - 1) / kTH);
            // round 1's numbers: a tile alone on an SM needs L = 3 400 cycles per slice, an SM with 4 resident tiles 4 875 (T = 1 220 per tile
            // and slice); one pixel per thread scales both by 446 / 611 (SASS of the slice loop) and doubles the tiles, so it wins while the
            // doubled tiles still fit three to an SM: max(0.73 L, ceil(2 t / SMs) * 0.73 T) < max(L, ceil(t / SMs) * T)
            px = (2 * tiles64 <= 3ll * sms) ? 1 : 2;
        }
        if (!unit_stride) {
            if (want == 2) return not_handled("two pixels per thread need unit stride from pixel to data voxel");
            px = 1;
        }
    }
    const int kTW = 32 * px, kFpW = kTW + 4;  // shadow the two-pixel constants below
    const bool clip = !clip_is_inactive(u, T);
    // transposed axis order (p,q,s) in native axes
    const int pa = u.axis == 0 ? 1 : 0, qa = u.axis == 2 ? 1 : 2, sa = u.axis;
    // data box extents: tile (+1 tap, + spread of i0 - c), inner extent rounded up to 16 bytes (+4 for the funnel read)
    // TMA needs the box origin 16-byte aligned along the innermost dimension: round the u8 box start down to 16 voxels
    // (tile origins are multiples of 64), and widen the box accordingly.
    const int dmin_p_al = (int) floorf((float) T.dmin[pa] / 16.0f) * 16;
    struct StageLayout {
        int dext[3], light_bytes, data_off, data_bytes, stage_bytes;
        size_t smem;
    };
    auto layout_of = [&](int h) {  // SMEM of the first-generation kernel for tiles of h rows
        StageLayout L;
        // first tap of the first pixel .. second tap of the last pixel: dk * (pixels - 1) + spread of (i0 - dk * c) + 2
        const int ext_p = T.dk[pa] * (kTW - 1) + (T.dmax[pa] - dmin_p_al) + 2, ext_q = T.dk[qa] * (h - 1) + (T.dmax[qa] - T.dmin[qa]) + 2,
                  ext_s = T.dk[sa] * (kSB - 1) + (T.dmax[sa] - T.dmin[sa]) + 2;
        L.dext[0] = (ext_p + 4 + 15) / 16 * 16, L.dext[1] = ext_q, L.dext[2] = ext_s;
        L.light_bytes = kTW * h * kSB * (byte_brick ? 1 : 4);
        L.data_off = L.light_bytes + (mode == kModeCombine ? kTW * h * kSB * 4 : 0);  // combining: the removed light's R32F brick follows
        L.data_bytes = L.dext[0] * L.dext[1] * L.dext[2];
        L.stage_bytes = (L.data_off + L.data_bytes + 16 + 127) / 128 * 128;
        L.smem = (size_t) kStages * L.stage_bytes + (2 * kFpW * kFpH + 256) * sizeof(float) + kStages * sizeof(uint64_t) + 16;
        return L;
    };
    // ---- which part of the pass this GPU runs, and in how many co-resident waves (bands of tile rows) ----
    const tbrm_slab& sl = r.slab;
    const bool sharded = sl.nranks > 1;
    const bool shard_q = sharded && u.axis != 2, shard_s = sharded && u.axis == 2;  // q = z for sweeps along X and Y
    const int ns = u.td[2];
    int q_begin = 0, q_end = ty, k_begin = 0, k_end = ns;
    if (shard_q) q_begin = sl.z_begin, q_end = sl.z_end;
    if (shard_s) {
        k_begin = u.dirn > 0 ? sl.z_begin : ns - sl.z_end;
        k_end = u.dirn > 0 ? sl.z_end : ns - sl.z_begin;
    }
    const int reach_lo = std::max(0, -T.bmin[1]), reach_hi = std::max(0, T.bmax[1] + 1);
    // Tile rows. All tiles of a wave are co-resident and advance in lock step (each waits for its upstream neighbours every slice), so a wave
    // runs at the pace of the fullest SM: 512^2 pixels in 64 x 8 tiles are 512 tiles on 148 SMs — 4 on most, 3 on the rest. 64 x 7 tiles are
    // 592 = 4 x 148: every SM holds four 7-warp blocks, an eighth less work on the SMs that set the pace (measured: 5.02 -> 4.65 ms for the cfg2
    // reset; 6 rows: 5.03 ms). The same holds for the waves (bands) of a larger plane: 1024^2 is 4 bands of 512 tiles or of 592. Passes whose
    // tile rows are cut by slab boundaries (sweeps along X / Y of a sharded volume: the exchange is laid out in 8-row units) keep 8 rows; so do
    // G8 slabs and the one-pixel G8 form. TBRM_SWEEP_TH=7|8 or bits 8-9 of reserved[0] (2 / 3) ask for a height.
    int th = ::tbrm::kTH;
    const bool slab_launch = sharded || r.options.reserved[2] > 0;  // (more bands than one are found below)
    // the slabs of a sweep along X / Y must tile alike on every rank (a rank addresses its neighbours' bands): equal slabs only
    const bool equal_slabs = !shard_q || (u.ldims[2] % (8 * sl.nranks) == 0);
    const int span = q_end - q_begin;  // buffer rows this GPU sweeps
    if (!ws && !(l8 && (px == 1 || slab_launch)) && equal_slabs) {
        static const int env_th = [] {
            const char* e = getenv("TBRM_SWEEP_TH");
            return e ? atoi(e) : 0;
        }();
        const int opt_th = (r.options.reserved[0] >> 8) & 3;
        const int want = opt_th >= 2 ? 5 + opt_th : env_th;
        const long long ntx_ = (tx + kTW - 1) / kTW, slots = 4ll * sms;  // launch bounds: four blocks per SM
        auto tiles_of = [&](int h) { return ntx_ * ((span + h - 1) / h); };
        // cost of the pass in (rows of the fullest SM) x bands: bands of at most `slots` tiles, split evenly
        auto cost_of = [&](int h) {
            const long long nty_ = (span + h - 1) / h, cap = std::max(1ll, slots / ntx_), bands = (nty_ + cap - 1) / cap, rows = (nty_ + bands - 1) / bands;
            return bands * ((rows * ntx_ + sms - 1) / sms) * h;
        };
        const bool pays = tiles_of(8) > sms && cost_of(7) < cost_of(8) && !l8;
        const bool pays_l8 = l8 && tiles_of(8) > sms && tiles_of(7) <= slots && cost_of(7) < cost_of(8);  // G8: single wave only
        if (want == 7 || (want != 8 && (pays || pays_l8))) {  // ... if four 7-row blocks are resident per SM (or one wave holds the plane anyway)
            int occ = 0, occ_s = 4;
            bool fits = false, fits_s = true;
            const void* k7 = l8 ? tma_kernel_l8(u.axis, clip, false, px, 7) : tma_kernel(u.axis, clip, false, px, 7);
            if ((e = blocks_per_sm(k7, 32 * 7, layout_of(7).smem, dev, &occ, &fits)) != cudaSuccess) return e;
            if (!l8 && (e = blocks_per_sm(tma_kernel(u.axis, clip, true, px, 7), 32 * 7, layout_of(7).smem, dev, &occ_s, &fits_s)) != cudaSuccess) return e;
            const long long per = std::min(occ, occ_s);
            // the rule of the banded launches below, for 7-row tiles: a footprint must not reach past the adjacent band (the last one may be thin)
            const long long reach = std::max(reach_lo, reach_hi), nty7 = (span + 6) / 7;
            long long cap = std::max(1ll, per * sms / ntx_);
            if (r.options.reserved[2] > 0) cap = std::min<long long>(cap, r.options.reserved[2]);
            const long long nb = (nty7 + cap - 1) / cap, rpb = (nty7 + nb - 1) / nb;
            const bool bands_ok = (nb == 1 && !sharded) || reach <= std::min<long long>(7, span - (nb - 1) * rpb * 7);
            if (fits && fits_s && bands_ok && (per * sms >= tiles_of(7) || (!l8 && per >= 4))) th = 7;  // (G8: one wave of the plain kernel must hold the plane)
        }
    }
    // Tile rows are anchored at the first row this GPU sweeps (row j covers [q_org + j th, q_org + (j + 1) th), q_org in (-th, 0]): the slab
    // of a sweep along X / Y starts on a tile row whatever th; its last tile row may end past the slab — those pixels are masked and the
    // light maps are clipped at the slab's end (below).
    const int j_first = (q_begin + th - 1) / th, q_org = q_begin - j_first * th;
    P.q_org = q_org;
    const int kTH = th;  // shadow the 8-row constant below
    P.ntx = (tx + kTW - 1) / kTW, P.nty = (ty - q_org + kTH - 1) / kTH;
    const int ntiles = P.ntx * P.nty;
    const int nat[3] = {pa, qa, sa};
    for (int t = 0; t < 3; ++t) {
        P.dmin[t] = T.dmin[nat[t]], P.dk[t] = T.dk[nat[t]];
        P.data_dims_t[t] = r.ddims[nat[t]];
    }
    P.dmin[0] = dmin_p_al;
    const StageLayout L = layout_of(th);
    for (int t = 0; t < 3; ++t) P.dext[t] = L.dext[t];
    if (P.dext[0] > 256 || P.dext[1] > 256 || P.dext[2] > 256) return not_handled("P.dext[0] > 256 || P.dext[1] > 256 || P.dext[2] > 256");
    for (int d = 0; d < 2; ++d) P.bmin[d] = T.bmin[d], P.bext[d] = (d ? kTH : kTW) + (T.bmax[d] - T.bmin[d]) + 1;
    if (P.bext[0] > kFpW || P.bext[1] > kFpH) return not_handled("P.bext[0] > kFpW || P.bext[1] > kFpH");
    if (u.td[2] >= 65000) return not_handled("u.td[2] >= 65000");  // slice index + 1 must fit the 16-bit tag field
    if ((long long) tx * ty * kRingDepth >= (1ll << 31)) return not_handled("ring cells do not fit 32-bit indices");
    // dependency lists must fit
    {
        const long long nx = (long long) (P.bext[0] + kTW - 1) / kTW + 1, ny = (long long) (P.bext[1] + kTH - 1) / kTH + 1;
        if (nx * ny - 1 > kFusedMaxDeps) return not_handled("nx * ny - 1 > kFusedMaxDeps");
    }
    // tensor maps. Light: native (X,Y,Z) fp32, box = tile x SB along the sweep axis.
    int lbox[3], ldims[3] = {r.ldims[0], r.ldims[1], r.ldims[2]};
    lbox[pa] = kTW, lbox[qa] = kTH, lbox[sa] = kSB;
    // a slab that ends inside its last tile row: the light, scratch and push maps end with the slab (q = z is the outermost dimension, so
    // the strides stay), loads beyond it read zeros nobody uses and stores beyond it are dropped
    if (shard_q && (q_end - q_org) % kTH != 0) ldims[2] = q_end;
    CUtensorMap lm, dm, sm;
    // the brick that is loaded, updated and stored: the light volume, or the scratch volume when the light is only stored
    const bool l8_x = byte_brick && u.axis == 0;  // a G8 volume's sweep along X works on a (y,z,x)-ordered copy of the light volume (see the kernel)
    if (l8_x) {
        const size_t lbytes = r.light_voxels();
        if (r.light_perm_bytes < lbytes) {
            if (r.light_perm[0]) cudaStreamSynchronize(r.stream), cudaFree(r.light_perm[0]), cudaFree(r.light_perm[1]);
            r.light_perm[0] = r.light_perm[1] = nullptr, r.light_perm_bytes = 0;
            if ((e = cudaMalloc(&r.light_perm[0], lbytes)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&r.light_perm[1], lbytes)) != cudaSuccess) return e;
            r.light_perm_bytes = lbytes;
        }
        const int pd[3] = {r.ldims[1], r.ldims[2], r.ldims[0]}, pb[3] = {kTW, kTH, kSB};  // (p,q,s) = (y,z,x)
        if (!make_map3(&lm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, r.light_perm[0], pd, pb)) return not_handled("tensor map of the permuted G8 light volume");
    } else if (byte_brick ? !make_map3(&lm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, r.light, ldims, lbox)
           : !make_map3(&lm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, mode == kModeStore ? r.change_scratch : r.light, ldims, lbox))
        return not_handled("tensor map of the light volume");
    if (l8 && mode != kModeCombine)
        sm = lm;  // never read
    else if (!make_map3(&sm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, mode == kModeCombine ? r.change_scratch : r.light, ldims, lbox))
        return not_handled("tensor map of the scratch volume");
    P.mode = mode;
    // push-gather: the same box over every other rank's light volume (the brick that updates the LIGHT volume: add / combine launches)
    static thread_local PushMaps pm;
    P.n_push = 0;
    if (r.push_this_pass && mode != kModeStore && r.slab.nranks > 1 && l8) {
        set_last_error("push-gather: R32F light volumes only (the pushed bricks are float boxes)");
        return cudaErrorNotSupported;
    }
    if (r.push_this_pass && mode != kModeStore && r.slab.nranks > 1) {
        for (int pr = 0; pr < r.slab.nranks; ++pr) {
            if (pr == r.slab.rank) continue;
            if (!r.peer_light[pr] || P.n_push >= kMaxPush) {
                set_last_error("push-gather: the light volumes of all ranks must be connected (tbrm_slab_open_peer_light / tbrm_slab_set_peer_light)");
                return cudaErrorNotSupported;
            }
            if (!make_map3(&pm.m[P.n_push], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, r.peer_light[pr], ldims, lbox)) return not_handled("tensor map of a peer light volume");
            ++P.n_push;
        }
    }
    // light SMEM strides: box is stored with native x fastest, then y, then z
    {
        const int str[3] = {1, lbox[0], lbox[0] * lbox[1]};
        P.ls_p = str[pa], P.ls_q = str[qa], P.ls_s = str[sa];
        P.ss_p = P.ls_p, P.ss_q = P.ls_q, P.ss_s = P.ls_s;  // the scratch brick (combining) is always in native order
        if (l8_x) P.ls_p = 1, P.ls_q = kTW, P.ls_s = kTW * kTH;
    }
    if (u.axis == 0) {
        if ((e = ensure_replica(r)) != cudaSuccess) return e;
        const int dd[3] = {Y, Z, X}, db[3] = {P.dext[0], P.dext[1], P.dext[2]};
        if (!make_map3(&dm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, r.data_yzx, dd, db)) return not_handled("!make_map3(&dm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, r.data_yzx, dd, db)");
        P.ds_q = P.dext[0], P.ds_s = P.dext[0] * P.dext[1];
    } else {
        int dd[3] = {X, Y, Z}, db[3];
        db[pa] = P.dext[0], db[qa] = P.dext[1], db[sa] = P.dext[2];
        if (!make_map3(&dm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, r.data, dd, db)) return not_handled("!make_map3(&dm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, r.data, dd, db)");
        const int str[3] = {1, db[0], db[0] * db[1]};
        P.ds_q = str[qa], P.ds_s = str[sa];
    }
    P.light_bytes = L.light_bytes, P.data_off = L.data_off, P.data_bytes = L.data_bytes, P.stage_bytes = L.stage_bytes;
    const int threads = ws ? kChThreads : 32 * th;
    size_t smem = L.smem;
    if (ws) {  // T bricks, light bricks, footprints, mbarriers
        smem = (size_t) (kChTStages + kChLStages) * P.light_bytes + (size_t) 2 * kFpW * kFpH * sizeof(float) + (kChTStages + kChLStages) * sizeof(uint64_t) + 16;
        if (P.bext[0] * P.bext[1] - kChThreads > kChHaloOverflow) return not_handled("halo list");
    }
    P.scratch = (const float*) r.change_scratch;

    // ---- in how many co-resident waves (bands of tile rows) this GPU's part of the pass runs ----
    int per_sm = 0;
    const void* kern_plain = ws ? chain_kernel(u.axis, false, px) : (l8 ? tma_kernel_l8(u.axis, clip, false, px, th) : tma_kernel(u.axis, clip, false, px, th));
    const void* kern_slab = ws ? chain_kernel(u.axis, true, px) : (l8 ? tma_kernel_l8(u.axis, clip, true, px) : tma_kernel(u.axis, clip, true, px, th));
    {
        int occ_plain = 0, occ_slab = 0;
        bool fits = false;
        if ((e = blocks_per_sm(kern_plain, threads, smem, dev, &occ_plain, &fits)) != cudaSuccess) return e;
        if (!fits) return not_handled("shared memory per block");
        if ((e = blocks_per_sm(kern_slab, threads, smem, dev, &occ_slab, &fits)) != cudaSuccess) return e;
        if (!fits) return not_handled("shared memory per block");
        // (an unsharded pass in one wave runs the plain kernel; the G8 slab kernels have 8-row tiles only)
        per_sm = (th == 7 && !sharded) ? (l8 ? occ_plain : std::min(occ_plain, occ_slab)) : occ_slab;
    }
    int cap_rows = (int) std::min<long long>((long long) sms * per_sm / P.ntx, 1 << 20);  // tile rows of one co-resident wave
    if (r.options.reserved[2] > 0) cap_rows = std::min(cap_rows, r.options.reserved[2]);      // test hook: force banding on small planes
    if (cap_rows < 1) return not_handled("cap_rows < 1");
    const int tr0 = j_first, tr1 = (q_end - q_org + kTH - 1) / kTH;
    auto bands_of = [&](int qb, int qe) { return ((qe - qb + kTH - 1) / kTH + cap_rows - 1) / cap_rows; };  // (every slab is anchored at its own first row)
    const int nbands = bands_of(q_begin, q_end);
    const int rows_per_band = (tr1 - tr0 + nbands - 1) / nbands;
    const bool use_slab = sharded || nbands > 1;
    r.last_geom[0] = th, r.last_geom[1] = px, r.last_geom[2] = P.ntx * (tr1 - tr0), r.last_geom[3] = nbands;
    if (use_slab) {
        if (reach_lo + reach_hi > kInboxSlots) return not_handled("the footprints reach more than 8 rows into the neighbouring bands");
        // a footprint must not reach past the adjacent band (bands and slabs are at least 8 rows)
        if (std::max(reach_lo, reach_hi) > std::min(kTH, q_end - (q_org + (tr0 + (nbands - 1) * rows_per_band) * kTH)))
            return not_handled("the footprints reach past the adjacent band");
        if (nbands > 1 && reach_lo > 0 && reach_hi > 0) return not_handled("nbands > 1 && reach_lo > 0 && reach_hi > 0");  // bands run one after the other: one-way dependencies only
        // a slab that ends inside its last tile row needs the clipped light maps: 7-row R32F tiles (the G8 slab kernels have 8 rows, where
        // slabs of multiples of 8 slices end on a tile row)
        if ((q_end - q_org) % kTH != 0 && q_end != ty && (l8 || th != 7)) return not_handled("(q_end - q_org) % kTH != 0 && q_end != ty");
        if (k_begin % kSB != 0 || (k_end % kSB != 0 && k_end != ns) || (shard_s && ns % kSB != 0)) return not_handled("k_begin % kSB != 0 || (k_end % kSB != 0 && k_end != ns) || (shard_s && ns % kSB != 0)");
        if ((e = slab_ensure_arena(r)) != cudaSuccess) return e;
    }

    // scratch: ring, flags, tables
    const size_t ring_bytes = (size_t) kRingDepth * tx * ty * sizeof(unsigned long long);
    if (r.ring_bytes < ring_bytes) {
        if (r.ring) cudaStreamSynchronize(r.stream), cudaFree(r.ring);
        r.ring = nullptr, r.ring_bytes = 0;
        if ((e = cudaMalloc(&r.ring, ring_bytes)) != cudaSuccess) return e;
        r.ring_bytes = ring_bytes;
        r.ring_epoch = 0;
    }
    // Cell tags are (pass_seq << 16 | slice + 1): a fresh sequence number per pass means cells left by earlier passes never
    // match. pass_seq counts the TMA passes of this resource set; the slabs of a sharded volume execute the same passes, so
    // their numbers agree. The ring is cleared when it is new or the generic fused kernel (plain floats) used it since
    // (ring_epoch == 0); when the 16-bit sequence space is used up everything holding tags is cleared — locally in stream
    // order, for a sharded volume by tbrm_slab_reset_comm on every rank (the exchange arenas are written by the neighbours).
    if (r.pass_seq >= 0xfff0u) {
        if (sharded) {
            set_last_error("slab exchange: the tag sequence is used up, call tbrm_slab_reset_comm on every rank (between barriers)");
            return cudaErrorNotSupported;
        }
        r.pass_seq = 0, r.ring_epoch = 0;
        if (r.arena && (e = cudaMemsetAsync(r.arena, 0, r.arena_bytes, r.stream)) != cudaSuccess) return e;
    }
    if (r.ring_epoch == 0) {
        if ((e = cudaMemsetAsync(r.ring, 0, r.ring_bytes, r.stream)) != cudaSuccess) return e;
        r.ring_epoch = 1;
    }
    // the sequence number is committed only when nothing can fail before the launch any more: a rank that gave up on an allocation or a
    // missing peer below must not have consumed a number its neighbours did not
    P.epoch = r.pass_seq + 1;
    P.exp_flags = r.options.reserved[0];
    P.cut_lo_mode = 0, P.cut_lo_add = 0;
    int cut = -1;
    if (!(P.exp_flags & 1)) window_cut_byte(u.win, T.weights_lt_one, P.cut_lo_mode, P.cut_lo_add, &cut);  // reserved[0] bit 0 disables the skip
    const size_t flag_words = (size_t) ntiles * kFlagStride;  // one progress word (its own 32-byte sector) per tile
    if (r.flags_count < flag_words) {
        if (r.flags) cudaStreamSynchronize(r.stream), cudaFree(r.flags);
        r.flags = nullptr, r.flags_count = 0;
        if ((e = cudaMalloc((void**) &r.flags, flag_words * sizeof(unsigned int))) != cudaSuccess) return e;
        r.flags_count = flag_words;
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
    {
        static thread_local std::vector<unsigned char> staging;
        staging.clear();
        for (int a = 0; a < 3; ++a) {
            pack(staging, r, T.S[a].data(), T.S[a].size() * 4, (const void**) &P.A.ax[a].S);
            pack(staging, r, T.f[a].data(), T.f[a].size() * 4, (const void**) &P.A.ax[a].f);
            pack(staging, r, T.meta[a].data(), T.meta[a].size() * 8, (const void**) &P.A.ax[a].meta);
        }
        pack(staging, r, T.bx.data(), T.bx.size() * 8, (const void**) &P.A.bx);
        pack(staging, r, T.by.data(), T.by.size() * 8, (const void**) &P.A.by);
        if (staging.size() > r.tables_bytes) return cudaErrorInvalidValue;
        // a pageable-memory async copy returns once the source has been staged, so the vector may be reused by the next pass
        if ((e = cudaMemcpyAsync(r.tables, staging.data(), staging.size(), cudaMemcpyHostToDevice, r.stream)) != cudaSuccess) return e;
    }
    P.ring = (float*) r.ring;
    P.flags = r.flags;
    memset(&P.S, 0, sizeof(P.S));
    // every wait on another tile or launch gives up after the timeout and raises the error word (tbrm_slab_check reads it)
    if (!r.sweep_err) {
        if ((e = cudaMalloc((void**) &r.sweep_err, sizeof(unsigned int))) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(r.sweep_err, 0, sizeof(unsigned int), r.stream)) != cudaSuccess) return e;
    }
    P.timeout_ns = (unsigned long long) (r.slab_timeout_ms > 0 ? r.slab_timeout_ms : 4000) * 1000000ull;
    P.error = use_slab ? arena_word(r.arena, 2) : r.sweep_err;
    P.poll_limit = (unsigned int) std::min<unsigned long long>(P.timeout_ns / 256ull, 0x7fffffffull);
    if (ws) {
        // ---- occlusion_kernel: T = 1 - occlusion for this GPU's share of the pass (an ordinary launch: no inter-block dependency) ----
        const int nblocks = (ns + kSB - 1) / kSB;
        const size_t bricks = (size_t) ntiles * nblocks, tbytes = bricks * (size_t) P.light_bytes;
        if (r.tvol_bytes < tbytes) {
            if (r.tvol) cudaStreamSynchronize(r.stream), cudaFree(r.tvol);
            r.tvol = nullptr, r.tvol_bytes = 0;
            if ((e = cudaMalloc(&r.tvol, tbytes)) != cudaSuccess) return e;
            r.tvol_bytes = tbytes;
        }
        if (r.tones_bytes < bricks) {
            if (r.tones) cudaStreamSynchronize(r.stream), cudaFree(r.tones);
            r.tones = nullptr, r.tones_bytes = 0;
            if ((e = cudaMalloc(&r.tones, bricks)) != cudaSuccess) return e;
            r.tones_bytes = bricks;
        }
        P.tvol = (const float*) r.tvol, P.tones = (const unsigned char*) r.tones;
#ifdef TBRM_CHAIN_TIMERS
        if (!r.dbg && (e = cudaMalloc(&r.dbg, (size_t) 4096 * 4 * 8 * sizeof(long long))) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(r.dbg, 0, (size_t) 4096 * 4 * 8 * sizeof(long long), r.stream)) != cudaSuccess) return e;
        P.dbg = ntiles <= 4096 ? (long long*) r.dbg : nullptr;
#endif
        OccParams O;
        memset(&O, 0, sizeof(O));
        O.U = u, O.A = P.A;
        const long long sx = 1, sy = X, sz = (long long) X * Y;
        if (u.axis == 2) O.data = (const uint8_t*) r.data, O.dsq = sy, O.dss = sz;
        else if (u.axis == 1) O.data = (const uint8_t*) r.data, O.dsq = sz, O.dss = sy;
        else O.data = (const uint8_t*) r.data_yzx, O.dsq = Y, O.dss = (long long) Y * Z;  // the (y,z,x) replica: y fastest, then z, then x
        (void) sx;
        for (int t = 0; t < 3; ++t) O.data_dims_t[t] = P.data_dims_t[t];
        O.tvol = (float*) r.tvol, O.tones = (unsigned char*) r.tones;
        O.ntx = P.ntx, O.nblocks = nblocks;
        O.tile_row0 = tr0, O.tile_rows = tr1 - tr0;
        // native blocks of the slices this GPU walks (a sub-range only for the Z-slab of a sweep along Z)
        O.nb_begin = (u.dirn > 0 ? k_begin : ns - k_end) / kSB;
        O.nb_end = ((u.dirn > 0 ? k_end : ns - k_begin) + kSB - 1) / kSB;
        O.cut_lo_mode = P.cut_lo_mode, O.cut_lo_add = P.cut_lo_add;
        O.cut_byte = P.cut_lo_mode ? cut : -1;
        O.bricks = nullptr;
        if (O.cut_byte >= 0) {
            if ((e = ensure_bricks(r)) != cudaSuccess) return e;
            O.bricks = (const uint8_t*) r.bricks;
            for (int a = 0; a < 3; ++a) O.bdims[a] = (r.ddims[a] + 7) / 8;
            for (int t = 0; t < 3; ++t) O.tap_lo[t] = T.dmin[nat[t]], O.tap_hi[t] = T.dmax[nat[t]] + 1;
        }
        const int chunks = (O.nb_end - O.nb_begin + kOccChunk - 1) / kOccChunk;
        occlusion_fn occ = occlusion_kernel_of(u.axis, clip, px);
        occ<<<dim3((unsigned) (P.ntx * O.tile_rows * chunks)), dim3(256), 0, r.stream>>>(O, r.tf);
        count_launch();
        *launches += 1;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (!use_slab) {
        r.pass_seq = P.epoch;
        if ((e = cudaMemsetAsync(r.flags, 0, flag_words * sizeof(unsigned int), r.stream)) != cudaSuccess) return e;
        if (l8_x && (e = permute_bytes(r, (const uint8_t*) r.light, (uint8_t*) r.light_perm[0], r.ldims[0], r.ldims[1], r.ldims[2])) != cudaSuccess) return e;
        if ((e = tma_launch(r, kern_plain, threads, lm, dm, sm, pm, P, ntiles, smem)) != cudaSuccess) return e;
        count_launch();
        *launches += 1;
        if (l8_x) {  // back to (x,y,z): two more applications of the same transpose
            if ((e = permute_bytes(r, (const uint8_t*) r.light_perm[0], (uint8_t*) r.light_perm[1], r.ldims[1], r.ldims[2], r.ldims[0])) != cudaSuccess) return e;
            if ((e = permute_bytes(r, (const uint8_t*) r.light_perm[1], (uint8_t*) r.light, r.ldims[2], r.ldims[0], r.ldims[1])) != cudaSuccess) return e;
            *launches += 3;
        }
        *handled = true;
        return cudaSuccess;
    }

    // ---- partial launches: the bands of this GPU's share of the pass, upstream band first ----
    const unsigned int seq = P.epoch;
    void* peer_lo = shard_q ? r.peer_arena[0] : nullptr;  // neighbours along q exist only when q is the sharded axis
    void* peer_hi = shard_q ? r.peer_arena[1] : nullptr;
    if (shard_q && ((sl.rank > 0 && !peer_lo) || (sl.rank + 1 < sl.nranks && !peer_hi))) {
        set_last_error("slab exchange: neighbour arenas are not connected (tbrm_slab_open_peer / tbrm_slab_set_peer)");
        return cudaErrorNotSupported;
    }
    void* peer_up = nullptr;    // upstream / downstream slab of a sweep along Z
    void* peer_down = nullptr;
    if (shard_s) {
        peer_up = u.dirn > 0 ? r.peer_arena[0] : r.peer_arena[1];
        peer_down = u.dirn > 0 ? r.peer_arena[1] : r.peer_arena[0];
        if ((k_begin > 0 && !peer_up) || (k_end < ns && !peer_down)) {
            set_last_error("slab exchange: neighbour arenas are not connected (tbrm_slab_open_peer / tbrm_slab_set_peer)");
            return cudaErrorNotSupported;
        }
    }
    const unsigned long long timeout_ns = (unsigned long long) (r.slab_timeout_ms > 0 ? r.slab_timeout_ms : 4000) * 1000000ull;
    unsigned int* err_word = arena_word(r.arena, 2);
    r.pass_seq = P.epoch;
    if (sharded && seq > 2) {  // the neighbours must be done with the arena regions this pass overwrites
        const unsigned int* a_lo = r.peer_arena[0] ? arena_word(r.arena, 0) : nullptr;
        const unsigned int* a_hi = r.peer_arena[1] ? arena_word(r.arena, 1) : nullptr;
        slab_wait_kernel<<<1, 1, 0, r.stream>>>(a_lo, a_hi, seq - 2, err_word, timeout_ns);
        count_launch();
        *launches += 1;
    }
    // a G8 volume's sweep along X works on the (y,z,x)-ordered copy of the light volume (whole volume: only this GPU's slab matters)
    if (l8_x && (e = permute_bytes(r, (const uint8_t*) r.light, (uint8_t*) r.light_perm[0], r.ldims[0], r.ldims[1], r.ldims[2])) != cudaSuccess) return e;
    int nb_lo = nbands;  // bands of the lower neighbour's slab (same partition rule on every rank: tbrm_slab_partition)
    if (shard_q && sl.rank > 0) {
        int32_t zb, ze;
        slab_partition(u.ldims[2], sl.nranks, sl.rank - 1, &zb, &ze);
        nb_lo = bands_of(zb, ze);
    }
    const bool ascending = reach_lo > 0 || reach_hi == 0;  // footprints reach toward lower rows: lower bands are upstream
    for (int i = 0; i < nbands; ++i) {
        const int bi = ascending ? i : nbands - 1 - i;
        SlabParams& S = P.S;
        S.tile_row0 = tr0 + bi * rows_per_band;
        S.tile_rows = std::min(rows_per_band, tr1 - S.tile_row0);
        if (S.tile_rows <= 0) continue;
        S.q_lo = q_org + S.tile_row0 * kTH, S.q_hi = std::min(q_end, q_org + (S.tile_row0 + S.tile_rows) * kTH);
        S.k_begin = k_begin, S.k_end = k_end;
        S.reach_lo = reach_lo, S.reach_hi = reach_hi;
        S.inbox = arena_inbox(r.arena, r.ldims, arena_region(seq, bi));
        S.out_lo = S.out_hi = nullptr;
        if (reach_hi > 0) {  // the band below reads our first rows
            if (bi > 0)
                S.out_lo = arena_inbox(r.arena, r.ldims, arena_region(seq, bi - 1));
            else if (peer_lo)
                S.out_lo = arena_inbox(peer_lo, r.ldims, arena_region(seq, nb_lo - 1));  // the neighbour's last band
        }
        if (reach_lo > 0) {  // the band above reads our last rows
            if (bi < nbands - 1)
                S.out_hi = arena_inbox(r.arena, r.ldims, arena_region(seq, bi + 1));
            else if (peer_hi)
                S.out_hi = arena_inbox(peer_hi, r.ldims, arena_region(seq, 0));  // the neighbour's first band
        }
        S.zin = (shard_s && k_begin > 0) ? arena_zplane(r.arena, r.ldims, arena_region(seq, 0)) : nullptr;
        S.zout = (shard_s && k_end < ns) ? arena_zplane(peer_down, r.ldims, arena_region(seq, 0)) : nullptr;
        S.error = err_word;
        S.timeout_ns = timeout_ns;
        if ((e = cudaMemsetAsync(r.flags, 0, (size_t) ntiles * kFlagStride * sizeof(unsigned int), r.stream)) != cudaSuccess) return e;
        if ((e = tma_launch(r, kern_slab, threads, lm, dm, sm, pm, P, S.tile_rows * P.ntx, smem)) != cudaSuccess) return e;
        count_launch();
        *launches += 1;
    }
    if (l8_x) {  // back to (x,y,z): two more applications of the same transpose
        if ((e = permute_bytes(r, (const uint8_t*) r.light_perm[0], (uint8_t*) r.light_perm[1], r.ldims[1], r.ldims[2], r.ldims[0])) != cudaSuccess) return e;
        if ((e = permute_bytes(r, (const uint8_t*) r.light_perm[1], (uint8_t*) r.light, r.ldims[2], r.ldims[0], r.ldims[1])) != cudaSuccess) return e;
        *launches += 3;
    }
    if (sharded) {  // tell the neighbours this slab is done with pass `seq` (acks live in THEIR arenas: word 1 = "from hi" for our lower neighbour)
        unsigned int* to_lo = r.peer_arena[0] ? arena_word(r.peer_arena[0], 1) : nullptr;
        unsigned int* to_hi = r.peer_arena[1] ? arena_word(r.peer_arena[1], 0) : nullptr;
        slab_signal_kernel<<<1, 1, 0, r.stream>>>(to_lo, to_hi, seq);
        count_launch();
        *launches += 1;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    *handled = true;
    return cudaSuccess;
}

}  // namespace tbrm
