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
constexpr int kRingDepth = 8;

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

// correctly rounded v/255 for v in 0..255 without a division (Markstein: q = v*r; q += (v - q*255)*r)
__device__ __forceinline__ float decode_u8(uint32_t v) {
    const float x = (float) v;
    const float r = 0.003921568859368563f;  // RN(1/255)
    const float q = x * r;
    const float e = __fmaf_rn(-q, 255.0f, x);
    return __fmaf_rn(e, r, q);
}
// correctly rounded x / w given rw = RN(1/w)
__device__ __forceinline__ float div_markstein(float x, float w, float rw) {
    const float q = x * rw;
    const float e = __fmaf_rn(-q, w, x);
    return __fmaf_rn(e, rw, q);
}

// opacity toward the light of one voxel given its trilinear data value (WindowedSampling.usf:20-37), alpha only
__device__ __forceinline__ float opacity_from_value(float v, const Windowing& win, float rwidth, const float* s_alpha, float step) {
    const float pos = div_markstein(v - win.center + (win.width / 2.0f), win.width, rwidth);
    if ((pos < 0.0f && win.low > 0.0f) || (pos > 1.0f && win.high > 0.0f)) return 0.0f;
    int i0, i1;
    float f;
    tf_taps(pos, i0, i1, f);
    const float a = lerpf(s_alpha[i0], s_alpha[i1], f);
    return step_opacity(a, step);
}

// AXIS = native sweep axis. Transposed coordinates: AXIS 2 -> (p,q,s) = (x,y,z); 1 -> (x,z,y); 0 -> (y,z,x).
template <int AXIS, bool CLIP>
__global__ void __launch_bounds__(kTmaThreads, 4)
    sweep_tma_kernel(const __grid_constant__ CUtensorMap light_map, const __grid_constant__ CUtensorMap data_map, const TmaParams P,
                     const float4* __restrict__ tf) {
    constexpr int PA = (AXIS == 0) ? 1 : 0;                 // native axis of p
    constexpr int QA = (AXIS == 2) ? 1 : 2;                 // native axis of q
    constexpr int SA = AXIS;                                // native axis of s
    const SweepUniforms& U = P.U;
    const int tx = U.td[0], ty = U.td[1], ns = U.td[2];
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int tix = tile % P.ntx, tiy = tile / P.ntx;
    const int x0 = tix * kTW, y0 = tiy * kTH;
    const size_t plane = (size_t) tx * ty;

    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* stage_base = smem;
    float* s_fp = (float*) (smem + (size_t) kStages * P.stage_bytes);  // footprint of the previous slice
    float* s_alpha = s_fp + (kTW + 4) * (kTH + 4);
    uint64_t* s_bar = (uint64_t*) (s_alpha + 256);
    __shared__ int s_up[kFusedMaxDeps], s_down[kFusedMaxDeps];
    __shared__ int s_nup, s_ndown;

    const int fx0 = x0 + P.bmin[0], fy0 = y0 + P.bmin[1], FW = P.bext[0], FH = P.bext[1];
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) mbar_init(&s_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // tiles whose published slice we read (footprint) and tiles that read ours
        int nup = 0, ndown = 0;
        const int ax = max(fx0, 0) / kTW, bx = min(fx0 + FW - 1, tx - 1) / kTW;
        const int ay = max(fy0, 0) / kTH, by = min(fy0 + FH - 1, ty - 1) / kTH;
        if (fx0 + FW > 0 && fx0 < tx && fy0 + FH > 0 && fy0 < ty)
            for (int j = ay; j <= by; ++j)
                for (int i = ax; i <= bx; ++i)
                    if ((i != tix || j != tiy) && nup < kFusedMaxDeps) s_up[nup++] = j * P.ntx + i;
        // tile (i,j) reads [i*TW + bmin, i*TW + bmin + FW) : it touches us iff that interval meets [x0, x0 + TW)
        for (int j = 0; j < P.nty; ++j) {
            const int gy = j * kTH + P.bmin[1];
            if (gy + FH <= y0 || gy >= y0 + kTH) continue;
            for (int i = 0; i < P.ntx; ++i) {
                const int gx = i * kTW + P.bmin[0];
                if (gx + FW <= x0 || gx >= x0 + kTW) continue;
                if ((i != tix || j != tiy) && ndown < kFusedMaxDeps) s_down[ndown++] = j * P.ntx + i;
            }
        }
        s_nup = nup, s_ndown = ndown;
    }
    s_alpha[tid] = __ldg(&tf[tid]).w;

    // ---- per-thread invariants: 2 adjacent pixels (px, px+1) of row py --------------------------------------
    const int lx = (tid & 31) * 2, ly = tid >> 5;
    const int px = x0 + lx, py = y0 + ly;
    const bool v0 = px < tx && py < ty, v1 = px + 1 < tx && py < ty;
    const int pxc = min(px, tx - 1), px1c = min(px + 1, tx - 1), pyc = min(py, ty - 1);
    // data taps along p (3 columns shared by the two pixels) and q (2 rows)
    const int2 mp0 = __ldg(&P.A.ax[PA].meta[pxc]), mp1 = __ldg(&P.A.ax[PA].meta[px1c]), mq = __ldg(&P.A.ax[QA].meta[pyc]);
    const float fp0 = __ldg(&P.A.ax[PA].f[pxc]), fp1 = __ldg(&P.A.ax[PA].f[px1c]), fq = __ldg(&P.A.ax[QA].f[pyc]);
    const int dN_p = P.data_dims_t[0], dN_q = P.data_dims_t[1], dN_s = P.data_dims_t[2];
    const int col = mp0.x - (x0 + P.dmin[0]);  // column of the first tap inside the data box
    const int rowq = mq.x - (y0 + P.dmin[1]);
    const bool inP0 = (unsigned) mp0.x < (unsigned) dN_p, inP1 = (unsigned) (mp0.x + 1) < (unsigned) dN_p,
               inP2 = (unsigned) (mp0.x + 2) < (unsigned) dN_p;
    const bool inQ0 = (unsigned) mq.x < (unsigned) dN_q, inQ1 = (unsigned) (mq.x + 1) < (unsigned) dN_q;
    const bool all_pq = inP0 && inP1 && inP2 && inQ0 && inQ1;
    const bool inside_pq0 = mp0.y && mq.y, inside_pq1 = mp1.y && mq.y;
    float Sp0 = 0.f, Sp1 = 0.f, Sq = 0.f;
    if (CLIP) {
        Sp0 = __ldg(&P.A.ax[PA].S[pxc]), Sp1 = __ldg(&P.A.ax[PA].S[px1c]), Sq = __ldg(&P.A.ax[QA].S[pyc]);
    }
    // read-buffer bilinear: 3 columns x 2 rows of the footprint
    const int2 bxa = __ldg(&P.A.bx[pxc]), bxb = __ldg(&P.A.bx[px1c]), bya = __ldg(&P.A.by[pyc]);
    const float bfx0 = __int_as_float(bxa.y), bfx1 = __int_as_float(bxb.y), bfy = __int_as_float(bya.y);
    const int fcol = bxa.x - fx0, frow = bya.x - fy0;
    const float rwidth = 1.0f / U.win.width;
    const float step = U.a.step;
    const int shift = (col & 3) * 8;
    const int col4 = col & ~3;
    __syncthreads();
    const int nup = s_nup, ndown = s_ndown;

    const int nblocks = (ns + kSB - 1) / kSB;
    // native coordinates of block b's light box origin along s
    // blocks are aligned to multiples of kSB in native coordinates (TMA: 16-byte aligned inner coordinate when the
    // sweep axis is x); a descending sweep visits them last-to-first
    auto block_s0 = [&](int b) { return (U.dirn > 0 ? b : nblocks - 1 - b) * kSB; };
    auto issue_load = [&](int b) {
        const int st = b % kStages;
        unsigned char* sb = stage_base + (size_t) st * P.stage_bytes;
        mbar_expect_tx(&s_bar[st], (uint32_t) (P.light_bytes + P.data_bytes));
        const int s0 = block_s0(b);
        int lc[3], dc[3];
        lc[PA] = x0, lc[QA] = y0, lc[SA] = s0;  // light map is over native (x,y,z)
        tma_load_3d(sb, &light_map, lc[0], lc[1], lc[2], &s_bar[st]);
        // data map: native dims for Z / Y sweeps, the (y,z,x) replica for X sweeps, i.e. always (p,q,s)-ordered for X
        if (AXIS == 0) {
            tma_load_3d(sb + P.light_bytes, &data_map, x0 + P.dmin[0], y0 + P.dmin[1], s0 + P.dmin[2], &s_bar[st]);
        } else {
            dc[PA] = x0 + P.dmin[0], dc[QA] = y0 + P.dmin[1], dc[SA] = s0 + P.dmin[2];
            tma_load_3d(sb + P.light_bytes, &data_map, dc[0], dc[1], dc[2], &s_bar[st]);
        }
    };
    if (tid == 0) issue_load(0);

    for (int b = 0; b < nblocks; ++b) {
        const int st = b % kStages;
        if (tid == 0 && b + 1 < nblocks) {
            tma_wait_read<1>();  // the store that last read stage (b+1)%3 (block b-2) has finished reading SMEM
            issue_load(b + 1);
        }
        mbar_wait(&s_bar[st], (uint32_t) ((b / kStages) & 1));
        unsigned char* sb = stage_base + (size_t) st * P.stage_bytes;
        float* s_light = (float*) sb;
        const unsigned char* s_data = sb + P.light_bytes;
        const int s0 = block_s0(b);

#pragma unroll 1
        for (int sl = 0; sl < kSB; ++sl) {
            const int loop = s0 + (U.dirn > 0 ? sl : kSB - 1 - sl);
            if (loop >= ns) continue;
            const int k = U.dirn > 0 ? loop : ns - 1 - loop;  // position in sweep order
            // ---- A: opacity toward the light for this thread's two voxels (independent of the previous slice) ----
            const int2 ms = __ldg(&P.A.ax[SA].meta[loop]);
            const float fs = __ldg(&P.A.ax[SA].f[loop]);
            const int rows = ms.x - (s0 + P.dmin[2]);
            const bool inS0 = (unsigned) ms.x < (unsigned) dN_s, inS1 = (unsigned) (ms.x + 1) < (unsigned) dN_s;
            float w0 = 1.0f, w1 = 1.0f;
            if (CLIP) {
                const float Ss = __ldg(&P.A.ax[SA].S[loop]);
                float S0[3], S1[3];
                S0[PA] = Sp0, S0[QA] = Sq, S0[SA] = Ss;
                S1[PA] = Sp1, S1[QA] = Sq, S1[SA] = Ss;
                const float rx = (float) U.ldims[0], ry = (float) U.ldims[1], rz = (float) U.ldims[2];
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const float* S = v ? S1 : S0;
                    const float dist = dot3(S[0] - U.clip_center[0], S[1] - U.clip_center[1], S[2] - U.clip_center[2], U.clip_dir[0],
                                            U.clip_dir[1], U.clip_dir[2]);
                    const float ox = S[0] - (S[0] + U.clip_dir[0] * dist), oy = S[1] - (S[1] + U.clip_dir[1] * dist),
                                oz = S[2] - (S[2] + U.clip_dir[2] * dist);
                    const float vx = ox * rx, vy = oy * ry, vz = oz * rz;
                    const float vdist = sqrtf(dot3(vx, vy, vz, vx, vy, vz));
                    const float sgn = dist > 0.0f ? 1.0f : (dist < 0.0f ? -1.0f : 0.0f);
                    const float w = fminf(fmaxf(0.5f + (0.57735026919f * vdist * sgn), 0.0f), 1.0f);
                    if (v) w1 = w; else w0 = w;
                }
            }
            const bool g0 = v0 && w0 > 0.0f && inside_pq0 && ms.y, g1 = v1 && w1 > 0.0f && inside_pq1 && ms.y;
            float cs0 = 0.0f, cs1 = 0.0f;
            if (g0 || g1) {
                // 4 rows (q, s) x 3 columns of taps; two aligned 32-bit loads + a funnel shift per row
                float t[3][2][2];
                const bool all_in = all_pq && inS0 && inS1;
#pragma unroll
                for (int js = 0; js < 2; ++js)
#pragma unroll
                    for (int jq = 0; jq < 2; ++jq) {
                        const unsigned char* rowp = s_data + (size_t) (rowq + jq) * P.ds_q + (size_t) (rows + js) * P.ds_s + col4;
                        const uint32_t a = *(const uint32_t*) rowp, bb = *(const uint32_t*) (rowp + 4);
                        const uint32_t w = __funnelshift_r(a, bb, shift);
                        t[0][jq][js] = decode_u8(w & 0xffu);
                        t[1][jq][js] = decode_u8((w >> 8) & 0xffu);
                        t[2][jq][js] = decode_u8((w >> 16) & 0xffu);
                    }
                if (!all_in) {
                    const bool ip[3] = {inP0, inP1, inP2}, iq[2] = {inQ0, inQ1}, is[2] = {inS0, inS1};
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int jq = 0; jq < 2; ++jq)
#pragma unroll
                            for (int js = 0; js < 2; ++js)
                                if (!(ip[c] && iq[jq] && is[js])) t[c][jq][js] = U.data_border;
                }
                float val0, val1;
                if (AXIS == 2) {  // x = p, y = q, z = s
                    const float a00 = lerpf(t[0][0][0], t[1][0][0], fp0), a10 = lerpf(t[0][1][0], t[1][1][0], fp0);
                    const float a01 = lerpf(t[0][0][1], t[1][0][1], fp0), a11 = lerpf(t[0][1][1], t[1][1][1], fp0);
                    val0 = lerpf(lerpf(a00, a10, fq), lerpf(a01, a11, fq), fs);
                    const float b00 = lerpf(t[1][0][0], t[2][0][0], fp1), b10 = lerpf(t[1][1][0], t[2][1][0], fp1);
                    const float b01 = lerpf(t[1][0][1], t[2][0][1], fp1), b11 = lerpf(t[1][1][1], t[2][1][1], fp1);
                    val1 = lerpf(lerpf(b00, b10, fq), lerpf(b01, b11, fq), fs);
                } else if (AXIS == 1) {  // x = p, y = s, z = q
                    const float a00 = lerpf(t[0][0][0], t[1][0][0], fp0), a10 = lerpf(t[0][1][0], t[1][1][0], fp0);
                    const float a01 = lerpf(t[0][0][1], t[1][0][1], fp0), a11 = lerpf(t[0][1][1], t[1][1][1], fp0);
                    val0 = lerpf(lerpf(a00, a01, fs), lerpf(a10, a11, fs), fq);
                    const float b00 = lerpf(t[1][0][0], t[2][0][0], fp1), b10 = lerpf(t[1][1][0], t[2][1][0], fp1);
                    const float b01 = lerpf(t[1][0][1], t[2][0][1], fp1), b11 = lerpf(t[1][1][1], t[2][1][1], fp1);
                    val1 = lerpf(lerpf(b00, b01, fs), lerpf(b10, b11, fs), fq);
                } else {  // x = s, y = p, z = q
                    const float d00 = lerpf(t[0][0][0], t[0][0][1], fs), d10 = lerpf(t[1][0][0], t[1][0][1], fs),
                                d20 = lerpf(t[2][0][0], t[2][0][1], fs);
                    const float d01 = lerpf(t[0][1][0], t[0][1][1], fs), d11 = lerpf(t[1][1][0], t[1][1][1], fs),
                                d21 = lerpf(t[2][1][0], t[2][1][1], fs);
                    val0 = lerpf(lerpf(d00, d10, fp0), lerpf(d01, d11, fp0), fq);
                    val1 = lerpf(lerpf(d10, d20, fp1), lerpf(d11, d21, fp1), fq);
                }
                if (g0) cs0 = opacity_from_value(val0, U.win, rwidth, s_alpha, step) * w0;
                if (g1) cs1 = opacity_from_value(val1, U.win, rwidth, s_alpha, step) * w1;
            }

            // ---- B: wait for the tiles we read (slice k-1 published) and the tiles that read the slot we overwrite ----
            if (k > 0) {
                if (tid < nup) {
                    const unsigned int* f = P.flags + (size_t) s_up[tid] * kFlagStride;
                    while (ld_acquire(f) < (unsigned) k) {
                    }
                } else if (tid >= 32 && tid - 32 < ndown && k >= kRingDepth) {
                    const unsigned int* f = P.flags + (size_t) s_down[tid - 32] * kFlagStride;
                    while (ld_acquire(f) < (unsigned) (k - kRingDepth + 2)) {
                    }
                }
            }
            __syncthreads();
            // ---- C: footprint of slice k-1 (L2 ring; slice -1 is the cleared buffer = LightAlpha) ----
            {
                const float* rd = P.ring + (size_t) ((k + kRingDepth - 1) % kRingDepth) * plane;
                for (int c = tid; c < FW * FH; c += kTmaThreads) {
                    const int gx = fx0 + c % FW, gy = fy0 + c / FW;
                    const bool in = (unsigned) gx < (unsigned) tx && (unsigned) gy < (unsigned) ty;
                    s_fp[c] = in ? (k > 0 ? __ldcg(rd + (size_t) gx + (size_t) tx * gy) : U.a.light_alpha) : U.a.border;
                }
            }
            __syncthreads();
            // ---- D: propagate, publish to the ring, accumulate into the light brick ----
            {
                const float* r0 = s_fp + frow * FW + fcol;
                const float* r1 = r0 + FW;
                const float t00 = r0[0], t10 = r0[1], t20 = r0[2], t01 = r1[0], t11 = r1[1], t21 = r1[2];
                const float prev0 = lerpf(lerpf(t00, t10, bfx0), lerpf(t01, t11, bfx0), bfy);
                const float prev1 = lerpf(lerpf(t10, t20, bfx1), lerpf(t11, t21, bfx1), bfy);
                const float cur0 = prev0 * (1.0f - cs0), cur1 = prev1 * (1.0f - cs1);
                float* wr = P.ring + (size_t) (k % kRingDepth) * plane + (size_t) px + (size_t) tx * py;
                if (v1)
                    __stcg((float2*) wr, make_float2(cur0, cur1));
                else if (v0)
                    __stcg(wr, cur0);
                const int sli = loop - s0;
                float* lp = s_light + lx * P.ls_p + ly * P.ls_q + sli * P.ls_s;
                if (v0 && fabsf(cur0) > 1e-3f) lp[0] = lp[0] + (cur0 * U.sign);
                if (v1 && fabsf(cur1) > 1e-3f) lp[P.ls_p] = lp[P.ls_p] + (cur1 * U.sign);
            }
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                st_release(P.flags + (size_t) tile * kFlagStride, (unsigned) (k + 1));
            }
        }
        // ---- write the updated light brick back ----
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            int lc[3];
            lc[PA] = x0, lc[QA] = y0, lc[SA] = s0;
            tma_store_3d(&light_map, lc[0], lc[1], lc[2], s_light);
            tma_commit();
        }
    }
    if (tid == 0) tma_wait_all<0>();
}

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
    if (P.bext[0] > kTW + 4 || P.bext[1] > kTH + 4) return cudaSuccess;
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
    const size_t smem = (size_t) kStages * P.stage_bytes + ((kTW + 4) * (kTH + 4) + 256) * sizeof(float) + kStages * sizeof(uint64_t) + 16;

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
    const size_t ring_bytes = (size_t) kRingDepth * tx * ty * sizeof(float);
    if (r.ring_bytes < ring_bytes) {
        if (r.ring) cudaStreamSynchronize(r.stream), cudaFree(r.ring);
        r.ring = nullptr, r.ring_bytes = 0;
        if ((e = cudaMalloc(&r.ring, ring_bytes)) != cudaSuccess) return e;
        r.ring_bytes = ring_bytes;
    }
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
