// sweep_split.cuh — second generation of the TMA-staged fused plane-sweep (included by sweep_tma.cuh): the pass as TWO kernels.
//
// AddDirLightShader.usf:104-127 has two parts with different dependencies:
//   * the occlusion term of a voxel (clip weight, trilinear data sample, window, transfer function, pow: :104-114) depends on nothing the
//     sweep produces — it is the bulk of the arithmetic (~75 fp32 operations per non-empty voxel) and embarrassingly parallel;
//   * the propagation  cur = bilinear(previous slice) * (1 - occlusion)  and the light-volume update (:116-127) form the serial chain along
//     the sweep axis, with a cheap body.
// sweep_tma_kernel (first generation, kept for A/B: TBRM_SWEEP_GEN=1) runs both in every thread slice by slice, so the pass costs
// slices x the whole per-slice instruction chain of its fullest tile. Here
//   1. occlusion_kernel — an ordinary grid launch, no inter-block dependency — writes T = 1 - occlusion for every voxel of the pass as
//      BRICKS of kSB slices x one tile, brick-major (a brick = 8 KiB contiguous), plus one flag per brick "every T is exactly 1" (empty
//      space under the low cut-off: such bricks are neither written nor read). A thread walks its two pixels along the sweep axis and
//      carries the decoded / partially interpolated tap plane from slice to slice (a data plane is decoded once, not twice); the exact
//      empty-space test runs on the raw tap bytes (SWAR) and skips decode, interpolation, window, TF and pow.
//   2. sweep_chain_kernel — the cooperative launch with the L2 ring / progress words / slab inboxes of the first generation — only
//      propagates: T bricks arrive by 1-D bulk TMA (3 stages), light bricks by tensor TMA (2 stages); per slice a thread issues its halo
//      load, multiplies the bilinear of the previous slice (SMEM) with T, forwards / exports the result and keeps it in a register; the
//      light brick is updated once per block with 16-byte SMEM accesses and written back by a TMA store.
// A fused, warp-specialised form (producer warps computing T inside the chain kernel) was built and measured first (round 2): it was
// paced by the producer warps of the fullest tiles and spent ~30 % of its issue slots polling; splitting the kernels lets the
// occlusion run at full-chip parallelism. Every per-voxel fp32 operation is that of the first generation, so results are bit-identical
// to it, to the per-slice schedule and to the oracle.
#pragma once

namespace tbrm {

#ifndef TBRM_POLL_SLEEP_NS
#define TBRM_POLL_SLEEP_NS 100
#endif
constexpr int kChThreads = 128;  // 4 warps: CPX = 4 adjacent pixels per thread on a 64-wide tile, 2 on a 32-wide one. The chain's cost is
                                 // per-thread bookkeeping per slice, not arithmetic: fewer, fatter threads with registers to spare (no
                                 // rematerialisation at 4 CTAs x 128 threads per SM) beat 256 thin ones (measured: 250 instructions per
                                 // thread and slice for 2 pixels with 256 threads)
constexpr int kChTStages = 3, kChLStages = 2;
constexpr int kChHaloDepth = 3;  // staging slots of a thread's halo cell: requested two slices before it is read
constexpr int kChService = 96;  // the thread that issues the TMA copies and posts the progress word: lane 0 of the last warp, whose threads
                                // hold no halo cell on ordinary footprints (<= 96 cells) and reach the barrier first
constexpr int kChHaloOverflow = kFpW * kFpH - kChThreads;  // halo cells beyond the one a thread keeps in a register
constexpr int kOccChunk = 8;                               // blocks of kSB slices an occlusion CTA walks (plane carry across the chunk)

struct OccParams {
    SweepUniforms U;
    LightTabs A;
    const uint8_t* data;      // R8 data volume (the (y,z,x) replica for sweeps along X): p is the fastest axis
    long long dsq, dss;       // byte strides along q and s
    int data_dims_t[3];       // data dims in transposed (p,q,s) order
    float* tvol;              // T bricks: [tile][native block][slot][row][p]
    unsigned char* tones;     // per brick: 1 = every T of the brick is exactly 1 (the brick is not written)
    int ntx, nblocks;         // tiles per plane row, blocks per pass
    int tile_row0, tile_rows; // tile rows to compute (a slab / band of the plane)
    int nb_begin, nb_end;     // native blocks to compute
    unsigned int cut_lo_mode, cut_lo_add;
    // coarse empty-space skipping: the brick max-grid of the data volume (1 byte per 8^3 brick = max over [8b, 8b + 8] per native axis,
    // raymarch.cu) lets a CTA drop a whole tile x block whose taps all lie at or below `cut_byte` without touching the voxels
    const uint8_t* bricks;    // or null
    int bdims[3];             // bricks per native axis
    int cut_byte;             // largest byte the low cut-off rejects (-1: no cut-off)
    int tap_lo[3], tap_hi[3]; // taps of coordinate c lie in [c + tap_lo, c + tap_hi], per transposed axis (p,q,s)
};

// A tile that has caught up with its upstream neighbour waits for every slice; hundreds of warps re-reading their cell as fast as L2
// answers take its request bandwidth away from the tiles that do the work (measured: the pass ran at twice the per-slice work of any
// tile). A waiter therefore sleeps between two looks.
__device__ __forceinline__ void poll_backoff() {
#ifndef TBRM_HOST_EMULATION
    __nanosleep(TBRM_POLL_SLEEP_NS);
#endif
}
// Slow paths of the waits on other tiles / launches, out of line (the hot loop stays small).
// Polls an LL cell until it carries `want_tag` (or the block gave up / the timeout expires); returns the last value read.
__device__ __noinline__ unsigned long long chain_poll_cell(const unsigned long long* cell, unsigned int want_tag, int sys, unsigned long long timeout_ns,
                                                           volatile int* abort_flag, unsigned int* error) {
    unsigned long long v = sys ? ld_relaxed_sys_u64(cell) : ld_relaxed_u64(cell);
    unsigned long long t0 = 0;
    unsigned int polls = 0;
    while ((unsigned int) (v >> 32) != want_tag && !*abort_flag) {
        poll_backoff();
        v = sys ? ld_relaxed_sys_u64(cell) : ld_relaxed_u64(cell);
        if ((++polls & 255u) == 0) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > timeout_ns) {
                *abort_flag = 1;
                atomicExch(error, 1u);
                break;
            }
        }
    }
    return v;
}
// Polls a progress word until it reaches `need`.
__device__ __noinline__ void chain_poll_flag(const unsigned int* flag, unsigned int need, unsigned long long timeout_ns, volatile int* abort_flag,
                                             unsigned int* error) {
    unsigned long long t0 = 0;
    unsigned int polls = 0;
    while (ld_relaxed_u32(flag) < need && !*abort_flag) {
        poll_backoff();
        if ((++polls & 255u) == 0) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > timeout_ns) {
                *abort_flag = 1;
                atomicExch(error, 1u);
                break;
            }
        }
    }
}

// ====================================================================================================================================
// 1. occlusion: T = 1 - opacity toward the light * clip weight, for every voxel of the pass
// ====================================================================================================================================
// One CTA = one tile (TW x kTH pixels, a warp per row, PX adjacent pixels per lane) x kOccChunk blocks of kSB slices, walked in native
// order. AXIS = native sweep axis; transposed coordinates: AXIS 2 -> (p,q,s) = (x,y,z); 1 -> (x,z,y); 0 -> (y,z,x).
template <int AXIS, bool CLIP, int PX>
__global__ void __launch_bounds__(256) occlusion_kernel(const OccParams P, const float4* __restrict__ tf) {
    constexpr int PA = (AXIS == 0) ? 1 : 0, QA = (AXIS == 2) ? 1 : 2, SA = AXIS;
    constexpr int TW = 32 * PX;
    constexpr int kBrick = kSB * kTH * TW;
    const SweepUniforms& U = P.U;
    const int tx = U.td[0], ty = U.td[1], ns = U.td[2];
    const int tid = threadIdx.x, row = tid >> 5, lane = tid & 31;
    const int tiles = P.ntx * P.tile_rows;
    const int tl = (int) blockIdx.x % tiles, chunk = (int) blockIdx.x / tiles;
    const int tix = tl % P.ntx, tiy = P.tile_row0 + tl / P.ntx;
    const int x0 = tix * TW, y0 = tiy * kTH;
    __shared__ float s_alpha[256];
    __shared__ int s_skip[kOccChunk];
    s_alpha[tid] = __ldg(&tf[tid]).w;
    {
        // warp w classifies block w of the chunk: every tap of the tile x block inside the volume and under bricks whose maximum the low
        // cut-off rejects => every sample returns exactly 0 (WindowedSampling.usf:28) => T == 1 throughout
        const int nb = P.nb_begin + chunk * kOccChunk + row;
        bool skip = false;
        if (P.bricks != nullptr && P.cut_byte >= 0 && row < kOccChunk && nb < P.nb_end) {
            int lo[3], hi[3];  // tap ranges, transposed (p,q,s)
            lo[0] = x0 + P.tap_lo[0], hi[0] = min(x0 + TW, tx) - 1 + P.tap_hi[0];
            lo[1] = y0 + P.tap_lo[1], hi[1] = min(y0 + kTH, ty) - 1 + P.tap_hi[1];
            lo[2] = nb * kSB + P.tap_lo[2], hi[2] = min(nb * kSB + kSB, ns) - 1 + P.tap_hi[2];
            bool inside = true;
#pragma unroll
            for (int t = 0; t < 3; ++t) inside = inside && lo[t] >= 0 && hi[t] < P.data_dims_t[t];
            if (inside) {
                int blo[3], bn[3];  // brick ranges in native axes
                blo[PA] = lo[0] >> 3, bn[PA] = (hi[0] >> 3) - blo[PA] + 1;
                blo[QA] = lo[1] >> 3, bn[QA] = (hi[1] >> 3) - blo[QA] + 1;
                blo[SA] = lo[2] >> 3, bn[SA] = (hi[2] >> 3) - blo[SA] + 1;
                unsigned int mx = 0;
                for (int i = lane; i < bn[0] * bn[1] * bn[2]; i += 32) {
                    const int bx = blo[0] + i % bn[0], by = blo[1] + (i / bn[0]) % bn[1], bz = blo[2] + i / (bn[0] * bn[1]);
                    mx = max(mx, (unsigned int) __ldg(P.bricks + ((size_t) bx + (size_t) P.bdims[0] * ((size_t) by + (size_t) P.bdims[1] * bz))));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                skip = (int) mx <= P.cut_byte;
            }
        }
        if (row < kOccChunk && lane == 0) s_skip[row] = skip ? 1 : 0;
    }
    __syncthreads();

    const int px = x0 + lane * PX, py = y0 + row;
    const int pxc = min(px, tx - 1), px1c = min(px + 1, tx - 1), pyc = min(py, ty - 1);
    const int2 mp0 = __ldg(&P.A.ax[PA].meta[pxc]), mp1 = __ldg(&P.A.ax[PA].meta[px1c]), mq = __ldg(&P.A.ax[QA].meta[pyc]);
    const float fp0 = __ldg(&P.A.ax[PA].f[pxc]), fp1 = __ldg(&P.A.ax[PA].f[px1c]), fq = __ldg(&P.A.ax[QA].f[pyc]);
    const int dN_p = P.data_dims_t[0], dN_q = P.data_dims_t[1], dN_s = P.data_dims_t[2];
    const bool inP = (unsigned) mp0.x < (unsigned) dN_p && (unsigned) (mp0.x + 1) < (unsigned) dN_p && (PX == 1 || (unsigned) (mp0.x + 2) < (unsigned) dN_p);
    const bool inQ = (unsigned) mq.x < (unsigned) dN_q && (unsigned) (mq.x + 1) < (unsigned) dN_q;
    const bool all_pq = inP && inQ;
    const bool gate = U.gate_saturate != 0;
    // AddDirLight samples only where GetUVW + UVWOffset is inside [0,1]^3 (AddDirLightShader.usf:110); ChangeDirLight has no such gate
    const bool gq = py < ty && (!gate || mq.y);
    const bool g0 = gq && px < tx && (!gate || mp0.y), g1 = PX == 2 && gq && px + 1 < tx && (!gate || mp1.y);
    float Sp0 = 0.f, Sp1 = 0.f, Sq = 0.f;
    if (CLIP) Sp0 = __ldg(&P.A.ax[PA].S[pxc]), Sp1 = __ldg(&P.A.ax[PA].S[px1c]), Sq = __ldg(&P.A.ax[QA].S[pyc]);
    const float rwidth = 1.0f / U.win.width, step = U.a.step;
    // the two aligned words that hold a row's 3 taps (rows are multiples of 16 bytes, so a word is inside or outside as a whole; outside
    // reads as 0 like a TMA box would, and the cold path below replaces taps outside the volume by the border colour)
    const int wa = mp0.x & ~3, shift = (mp0.x & 3) * 8;
    const bool in_a = (unsigned) wa < (unsigned) dN_p, in_b = (unsigned) (wa + 4) < (unsigned) dN_p;
    const int q0 = min(max(mq.x, 0), dN_q - 1), q1 = min(max(mq.x + 1, 0), dN_q - 1);  // rows outside are patched as a whole: any address will do
    const uint8_t* const row0 = P.data + (long long) q0 * P.dsq + wa;
    const uint8_t* const row1 = P.data + (long long) q1 * P.dsq + wa;
    constexpr uint32_t kTapFlags = PX == 2 ? 0x00808080u : 0x00008080u;  // the flag bits of the tap bytes in use
    auto tap_words = [&](int plane, uint32_t& w0, uint32_t& w1) {  // raw taps of one data plane: bytes 0..2 = the taps along p of the two rows
        const long long off = (long long) min(max(plane, 0), dN_s - 1) * P.dss;
        const uint32_t a0 = in_a ? __ldg((const uint32_t*) (row0 + off)) : 0u, b0 = in_b ? __ldg((const uint32_t*) (row0 + off + 4)) : 0u;
        const uint32_t a1 = in_a ? __ldg((const uint32_t*) (row1 + off)) : 0u, b1 = in_b ? __ldg((const uint32_t*) (row1 + off + 4)) : 0u;
        w0 = __funnelshift_r(a0, b0, shift), w1 = __funnelshift_r(a1, b1, shift);
    };
    // Exact empty-space test (SWAR "some byte > T", T = largest byte the low cut-off rejects): see sweep_tma_kernel.cuh
    auto any_above = [&](uint32_t w0, uint32_t w1) -> bool {
        if (P.cut_lo_mode == 1) return ((((w0 + P.cut_lo_add) | w0) | ((w1 + P.cut_lo_add) | w1)) & kTapFlags) != 0u;
        if (P.cut_lo_mode == 2) return (((((w0 & 0x7f7f7f7fu) + P.cut_lo_add) & w0) | (((w1 & 0x7f7f7f7fu) + P.cut_lo_add) & w1)) & kTapFlags) != 0u;
        return true;
    };
    // a decoded plane reduced as far as the interpolation order allows — AXIS 2 (x = p, y = q, z = s): a[v][0] = the plane's bilinear value;
    // AXIS 1 (x = p, y = s, z = q): a[v][jq] = its p-lerps; AXIS 0 (x = s first): t[c][jq] = the decoded taps
    struct Plane {
        float a[PX][2];
        float t[PX + 1][2];
    };
    auto reduce_plane = [&](int plane, uint32_t w0, uint32_t w1, Plane& o) {
        float t[PX + 1][2];
        t[0][0] = decode_u8(w0 & 0xffu), t[1][0] = decode_u8((w0 >> 8) & 0xffu);
        t[0][1] = decode_u8(w1 & 0xffu), t[1][1] = decode_u8((w1 >> 8) & 0xffu);
        if (PX == 2) t[PX][0] = decode_u8((w0 >> 16) & 0xffu), t[PX][1] = decode_u8((w1 >> 16) & 0xffu);
        const bool is = (unsigned) plane < (unsigned) dN_s;
        if (!(all_pq && is)) {  // cold (volume faces only)
            const bool ip[3] = {(unsigned) mp0.x < (unsigned) dN_p, (unsigned) (mp0.x + 1) < (unsigned) dN_p, (unsigned) (mp0.x + 2) < (unsigned) dN_p};
            const bool iq[2] = {(unsigned) mq.x < (unsigned) dN_q, (unsigned) (mq.x + 1) < (unsigned) dN_q};
#pragma unroll
            for (int c = 0; c <= PX; ++c)
#pragma unroll
                for (int jq = 0; jq < 2; ++jq)
                    if (!(ip[c] && iq[jq] && is)) t[c][jq] = U.data_border;
        }
        if (AXIS == 2) {
            o.a[0][0] = lerpf(lerpf(t[0][0], t[1][0], fp0), lerpf(t[0][1], t[1][1], fp0), fq);
            if (PX == 2) o.a[PX - 1][0] = lerpf(lerpf(t[1][0], t[PX][0], fp1), lerpf(t[1][1], t[PX][1], fp1), fq);
        } else if (AXIS == 1) {
            o.a[0][0] = lerpf(t[0][0], t[1][0], fp0), o.a[0][1] = lerpf(t[0][1], t[1][1], fp0);
            if (PX == 2) o.a[PX - 1][0] = lerpf(t[1][0], t[PX][0], fp1), o.a[PX - 1][1] = lerpf(t[1][1], t[PX][1], fp1);
        } else {
#pragma unroll
            for (int c = 0; c <= PX; ++c) o.t[c][0] = t[c][0], o.t[c][1] = t[c][1];
        }
    };

    const int nb0 = P.nb_begin + chunk * kOccChunk, nb1 = min(P.nb_end, nb0 + kOccChunk);
    if (nb0 >= nb1) return;
    // first tap plane of the chunk's first slice; the others follow (uniform tap pairs: the host checks i0(c + 1) == i0(c) + 1)
    int plane = __ldg(&P.A.ax[SA].meta[min(nb0 * kSB, ns - 1)]).x;
    uint32_t wp0 = 0, wp1 = 0;  // raw words of the plane the previous slice ended with (= the lower plane of the current slice)
    bool raw_valid = false;     // ... are loaded
    bool gt_prev = false;
    bool have_prev = false;  // `prev` holds the reduced form of that plane
    Plane prev;
#pragma unroll
    for (int v = 0; v < PX; ++v) prev.a[v][0] = prev.a[v][1] = 0.0f;
#pragma unroll
    for (int c = 0; c <= PX; ++c) prev.t[c][0] = prev.t[c][1] = 0.0f;

#pragma unroll 1
    for (int nb = nb0; nb < nb1; ++nb) {
        if (s_skip[nb - nb0]) {  // the whole tile x block is empty space
            if (tid == 0) P.tones[(size_t) (tiy * P.ntx + tix) * P.nblocks + nb] = 1;
            plane += kSB, raw_valid = false, have_prev = false;
            continue;
        }
        if (!raw_valid) {
            tap_words(plane, wp0, wp1);
            gt_prev = any_above(wp0, wp1);
            raw_valid = true;
        }
        float T[kSB][PX];
        bool ones = true;
#pragma unroll
        for (int sl = 0; sl < kSB; ++sl) {
            const int loop = nb * kSB + sl;
#pragma unroll
            for (int v = 0; v < PX; ++v) T[sl][v] = 1.0f;
            if (loop >= ns) continue;
            uint32_t wc0, wc1;  // the slice's upper plane
            tap_words(plane + 1, wc0, wc1);
            const bool gt_cur = any_above(wc0, wc1);
            const bool both_in = (unsigned) plane < (unsigned) (dN_s - 1);
            // both samples rejected by the low cut-off (they return exactly 0) when every tap is inside the volume and no byte exceeds T
            const bool empty = all_pq && both_in && !gt_prev && !gt_cur;
            const bool gs = !gate || __ldg(&P.A.ax[SA].meta[loop]).y;
            Plane cur;
#pragma unroll
            for (int v = 0; v < PX; ++v) cur.a[v][0] = cur.a[v][1] = 0.0f;
#pragma unroll
            for (int c = 0; c <= PX; ++c) cur.t[c][0] = cur.t[c][1] = 0.0f;
            bool have_cur = false;
            if (gs && !empty && (g0 || g1)) {
                const float fs = __ldg(&P.A.ax[SA].f[loop]);
                float w0 = 1.0f, w1 = 1.0f;
                if (CLIP) {  // clip weights of the slice's voxels (AddDirLightShader.usf:93-103)
                    const float Ss = __ldg(&P.A.ax[SA].S[loop]);
                    const float rx = (float) U.ldims[0], ry = (float) U.ldims[1], rz = (float) U.ldims[2];
#pragma unroll
                    for (int v = 0; v < PX; ++v) {
                        float S[3];
                        S[PA] = v ? Sp1 : Sp0, S[QA] = Sq, S[SA] = Ss;
                        const float dist = dot3(S[0] - U.clip_center[0], S[1] - U.clip_center[1], S[2] - U.clip_center[2], U.clip_dir[0], U.clip_dir[1],
                                                U.clip_dir[2]);
                        const float ox = S[0] - (S[0] + U.clip_dir[0] * dist), oy = S[1] - (S[1] + U.clip_dir[1] * dist),
                                    oz = S[2] - (S[2] + U.clip_dir[2] * dist);
                        const float vx = ox * rx, vy = oy * ry, vz = oz * rz;
                        const float vdist = sqrtf(dot3(vx, vy, vz, vx, vy, vz));
                        const float sgn = dist > 0.0f ? 1.0f : (dist < 0.0f ? -1.0f : 0.0f);
                        const float w = fminf(fmaxf(0.5f + (0.57735026919f * vdist * sgn), 0.0f), 1.0f);
                        if (v)
                            w1 = w;
                        else
                            w0 = w;
                    }
                }
                const bool e0 = g0 && w0 > 0.0f, e1 = PX == 2 && g1 && w1 > 0.0f;
                if (e0 || e1) {
                    if (!have_prev) reduce_plane(plane, wp0, wp1, prev);  // the previous slice was skipped: its plane is still raw
                    reduce_plane(plane + 1, wc0, wc1, cur);
                    have_cur = true;
                    float val0, val1 = 0.0f;
                    if (AXIS == 2) {
                        val0 = lerpf(prev.a[0][0], cur.a[0][0], fs);
                        if (PX == 2) val1 = lerpf(prev.a[PX - 1][0], cur.a[PX - 1][0], fs);
                    } else if (AXIS == 1) {
                        val0 = lerpf(lerpf(prev.a[0][0], cur.a[0][0], fs), lerpf(prev.a[0][1], cur.a[0][1], fs), fq);
                        if (PX == 2) val1 = lerpf(lerpf(prev.a[PX - 1][0], cur.a[PX - 1][0], fs), lerpf(prev.a[PX - 1][1], cur.a[PX - 1][1], fs), fq);
                    } else {
                        const float d00 = lerpf(prev.t[0][0], cur.t[0][0], fs), d10 = lerpf(prev.t[1][0], cur.t[1][0], fs);
                        const float d01 = lerpf(prev.t[0][1], cur.t[0][1], fs), d11 = lerpf(prev.t[1][1], cur.t[1][1], fs);
                        val0 = lerpf(lerpf(d00, d10, fp0), lerpf(d01, d11, fp0), fq);
                        if (PX == 2) {
                            const float d20 = lerpf(prev.t[PX][0], cur.t[PX][0], fs), d21 = lerpf(prev.t[PX][1], cur.t[PX][1], fs);
                            val1 = lerpf(lerpf(d10, d20, fp1), lerpf(d11, d21, fp1), fq);
                        }
                    }
                    // the factor the propagated light is multiplied with: 1 - opacity * clip weight
                    if (e0) T[sl][0] = 1.0f - (opacity_from_value(val0, U.win, rwidth, s_alpha, step) * w0);
                    if (e1) T[sl][PX - 1] = 1.0f - (opacity_from_value(val1, U.win, rwidth, s_alpha, step) * w1);
                    ones = ones && T[sl][0] == 1.0f && (PX == 1 || T[sl][PX - 1] == 1.0f);
                }
            }
            prev = cur, have_prev = have_cur;
            wp0 = wc0, wp1 = wc1, gt_prev = gt_cur;
            ++plane;
        }
        // the brick is left out when every factor of the tile x block is exactly 1 (bit pattern of 1.0f: the chain multiplies by it anyway)
        const bool all_ones = __syncthreads_and(ones);
        const size_t brick = (size_t) (tiy * P.ntx + tix) * P.nblocks + nb;
        if (tid == 0) P.tones[brick] = all_ones ? 1 : 0;
        if (!all_ones) {
            float* tb = P.tvol + brick * kBrick + row * TW + lane * PX;
#pragma unroll
            for (int sl = 0; sl < kSB; ++sl) {
                if (PX == 2)
                    *(float2*) (tb + sl * (kTH * TW)) = make_float2(T[sl][0], T[sl][PX - 1]);
                else
                    tb[sl * (kTH * TW)] = T[sl][0];
            }
        }
    }
}

// ====================================================================================================================================
// 2. the propagation chain
// ====================================================================================================================================
// Per slice k every thread
//  (a) issues the load of the halo cell it fetches from the L2 ring (slice k-1 of the upstream tiles) and reads its T from SMEM;
//  (c) checks the halo tag (polls only while the upstream tile is not yet ahead) and parks the value in SMEM;
//  (d) after ONE block barrier propagates: previous-slice taps from SMEM (own tile forwarded through SMEM, halo from the ring),
//      multiplication with T, forwarding, export of the cells other tiles read; the value stays in a register until the block ends.
// SLAB = true adds what a launch needs when it covers only part of the pass (SlabParams; see sweep_tma_kernel.cuh).
template <int AXIS, bool SLAB, int PX>
__global__ void __launch_bounds__(kChThreads, 4)
    sweep_chain_kernel(const __grid_constant__ CUtensorMap light_map, const __grid_constant__ CUtensorMap data_map,
                       const __grid_constant__ CUtensorMap scratch_map, const __grid_constant__ PushMaps push_maps, const TmaParams P,
                     const float4* __restrict__ tf) {
    constexpr int PA = (AXIS == 0) ? 1 : 0, QA = (AXIS == 2) ? 1 : 2, SA = AXIS;
    constexpr int TW = 32 * PX, FPW = TW + 4;
    constexpr int CPX = 2 * PX;             // adjacent pixels of a thread
    constexpr int kBrick = kSB * kTH * TW;  // floats of a T brick / light brick
    static_assert(PX == 1 || PX == 2, "tile width 32 or 64");
    static_assert(kSB == 4, "a pixel's slices of a block form one float4");
    static_assert(kChThreads * CPX == TW * kTH, "one thread per CPX pixels of the tile");
    (void) data_map, (void) scratch_map, (void) tf, (void) push_maps;
    const SweepUniforms& U = P.U;
    const int tx = U.td[0], ty = U.td[1], ns = U.td[2];
    const int tid = threadIdx.x;
    const int tix = (int) blockIdx.x % P.ntx, tiy = (SLAB ? P.S.tile_row0 : 0) + (int) blockIdx.x / P.ntx;
    const int tile = tiy * P.ntx + tix;
    const int row_lo = SLAB ? P.S.tile_row0 : 0, row_hi = SLAB ? P.S.tile_row0 + P.S.tile_rows : P.nty;
    const int k_begin = SLAB ? P.S.k_begin : 0, k_end = SLAB ? P.S.k_end : ns;
    const int x0 = tix * TW, y0 = tiy * kTH;
    const unsigned int plane32 = (unsigned int) (tx * ty);

    extern __shared__ __align__(128) unsigned char smem[];
    float* s_tstage = (float*) smem;                     // kChTStages T bricks  [slot][row][p]
    float* s_lstage = s_tstage + kChTStages * kBrick;    // kChLStages light bricks (native box order)
    float* s_fp = s_lstage + kChLStages * kBrick;        // 2 x footprint (ping-pong over slices)
    uint64_t* t_full = (uint64_t*) (s_fp + 2 * FPW * kFpH);  // [stage] T brick landed (or known to be all ones)
    uint64_t* l_full = t_full + kChTStages;                   // [stage] light brick landed
    __shared__ int s_down[kFusedMaxDeps];
    __shared__ int s_ndown;
    __shared__ volatile int s_abort;  // a wait on another tile / launch timed out; stop waiting (results are void)
    __shared__ int s_tones[kChTStages];  // the stage's brick is all ones (nothing was loaded)
    __shared__ unsigned short s_over_fp[kChHaloOverflow];  // footprint index of the halo cells beyond the per-thread register (0xffff: none)
    __shared__ ulonglong2 s_hstage[kChHaloDepth][kChThreads];  // asynchronously fetched halo cells (the 16-byte pair holding the thread's cell)

    const int fx0 = x0 + P.bmin[0], fy0 = y0 + P.bmin[1], FW = P.bext[0], FH = P.bext[1];
    const int nblocks = (ns + kSB - 1) / kSB;
    const int b_begin = k_begin / kSB, b_end = (k_end + kSB - 1) / kSB;  // SLAB slice ranges are multiples of kSB (host)
    const int nblk = b_end - b_begin;
    auto block_s0 = [&](int b) { return (U.dirn > 0 ? b : nblocks - 1 - b) * kSB; };  // native first slice of block b (sweep order)

    int ones_next = 0;  // service thread: the flag of the brick after the last one issued, fetched a block ahead of its use
    auto issue_t = [&](int n) {  // service thread, n = 0, 1, 2, ...: T brick of block n, unless the occlusion kernel found it to be all ones
        const int st = n % kChTStages;
        const size_t brick = (size_t) tile * nblocks + block_s0(b_begin + n) / kSB;
        const int ones = n == 0 ? (int) P.tones[brick] : ones_next;
        if (n + 1 < nblk) ones_next = P.tones[(size_t) tile * nblocks + block_s0(b_begin + n + 1) / kSB];
        s_tones[st] = ones;
        mbar_expect_tx(&t_full[st], ones ? 0u : (uint32_t) (kBrick * 4));
        if (!ones) bulk_load_1d(s_tstage + st * kBrick, P.tvol + brick * kBrick, (uint32_t) (kBrick * 4), &t_full[st]);
    };
    auto issue_light = [&](int n) {  // service thread
        const int st = n % kChLStages, s0 = block_s0(b_begin + n);
        mbar_expect_tx(&l_full[st], (uint32_t) (kBrick * 4));
        int lc[3];
        lc[PA] = x0, lc[QA] = y0, lc[SA] = s0;
        tma_load_3d(s_lstage + st * kBrick, &light_map, lc[0], lc[1], lc[2], &l_full[st]);
    };

    if (tid == kChService) {
        for (int i = 0; i < kChTStages; ++i) mbar_init(&t_full[i], 1);
        for (int i = 0; i < kChLStages; ++i) mbar_init(&l_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int n = 0; n < min(nblk, kChTStages); ++n) issue_t(n);
        for (int n = 0; n < min(nblk, kChLStages); ++n) issue_light(n);
        int ndown = 0;
        s_abort = 0;
        for (int j = row_lo; j < row_hi; ++j) {  // tiles that read cells of ours: tile (i,j) reads [i*TW + bmin, +FW) x [j*TH + bmin, +FH)
            const int gy = j * kTH + P.bmin[1];
            if (gy + FH <= y0 || gy >= y0 + kTH) continue;
            for (int i = 0; i < P.ntx; ++i) {
                const int gx = i * TW + P.bmin[0];
                if (gx + FW <= x0 || gx >= x0 + TW) continue;
                if ((i != tix || j != tiy) && ndown < kFusedMaxDeps) s_down[ndown++] = j * P.ntx + i;
            }
        }
        s_ndown = ndown;
    }
    for (int c = tid; c < FPW * kFpH; c += kChThreads) {  // out-of-plane cells: border colour for good; in-plane cells: the cleared buffer
        const int gx = fx0 + c % FPW, gy = fy0 + c / FPW;
        const bool in = (unsigned) gx < (unsigned) tx && (unsigned) gy < (unsigned) ty;
        const float v = in ? U.a.light_alpha : U.a.border;
        s_fp[c] = v;
        s_fp[FPW * kFpH + c] = v;
    }

    // ---- per-thread invariants: CPX adjacent pixels of row py -----------------------------------------------------------------
    const int row = tid >> 4, lx = (tid & 15) * CPX;
    const int px = x0 + lx, py = y0 + row;
    const int pxc = min(px, tx - 1), pyc = min(py, ty - 1);
    const bool vy = py < ty;
    const int2 bxa = __ldg(&P.A.bx[pxc]), bya = __ldg(&P.A.by[pyc]);
    const float bfy = __int_as_float(bya.y);
    float bfx[CPX];
    unsigned int vmask = 0, expmask = 0, ownmask = 0;  // per pixel: inside the plane / read by another tile / inside the own footprint
    auto exported = [&](int gx, int gy) {
        const int rx = gx - P.bmin[0], ry = gy - P.bmin[1];  // tile origin i*TW must lie in (rx - FW, rx]
        const int ia = max(0, (rx - FW + TW) / TW), ib = rx >= 0 ? min(P.ntx - 1, rx / TW) : -1;
        const int ja = max(row_lo, (ry - FH + kTH) / kTH), jb = ry >= 0 ? min(row_hi - 1, ry / kTH) : -1;
        for (int j = ja; j <= jb; ++j)
            for (int i = ia; i <= ib; ++i) {
                const int gx0 = i * TW + P.bmin[0], gy0 = j * kTH + P.bmin[1];
                if ((i != tix || j != tiy) && gx >= gx0 && gx < gx0 + FW && gy >= gy0 && gy < gy0 + FH) return true;
            }
        return false;
    };
#pragma unroll
    for (int i = 0; i < CPX; ++i) {
        bfx[i] = __int_as_float(__ldg(&P.A.bx[min(px + i, tx - 1)]).y);
        const bool v = vy && px + i < tx;
        if (v) vmask |= 1u << i;
        if (v && exported(px + i, py)) expmask |= 1u << i;
        if (v && px + i >= fx0 && px + i < fx0 + FW && py >= fy0 && py < fy0 + FH) ownmask |= 1u << i;
    }
    const int tap_idx = (bya.x - fy0) * FPW + (bxa.x - fx0);  // first of the (CPX + 1) x 2 read-buffer taps
    const int own_idx = (py - fy0) * FPW + (px - fx0);
    const unsigned int own_cell = (unsigned int) (px + tx * py);
    // halo cells: footprint cells inside the plane that belong to other tiles; cell h goes to thread h % 256: the first in a register, the
    // rest (footprints shifted far off the tile) through an SMEM-described list
    int halo_fp, halo_ring, n_halo;
    {
        const int ox0 = max(fx0, x0), ox1 = min(fx0 + FW, x0 + TW), oy0 = max(fy0, y0), oy1 = min(fy0 + FH, y0 + kTH);
        const int ow = max(0, ox1 - ox0), oh = (ow > 0) ? max(0, oy1 - oy0) : 0;
        const int top = (oh > 0 ? oy0 - fy0 : FH) * FW, mid = oh * (FW - ow);
        n_halo = FW * FH - ow * oh;
        auto halo_cell = [&](int h, int& fp_idx, int& ring_idx) {
            int gx = -1, gy = -1;
            fp_idx = -1, ring_idx = 0;
            if (h < top) {
                gy = fy0 + h / FW, gx = fx0 + h % FW;
            } else if (h - top < mid) {
                h -= top;
                const int r = h / (FW - ow), c = h % (FW - ow);
                gy = oy0 + r;
                gx = (c < ox0 - fx0) ? fx0 + c : c + ow + fx0;
            } else if (oh > 0) {
                h -= top + mid;
                gy = oy1 + h / FW, gx = fx0 + h % FW;
                if (gy >= fy0 + FH) gy = -1;
            }
            if ((unsigned) gx < (unsigned) tx && (unsigned) gy < (unsigned) ty) {
                fp_idx = (gy - fy0) * FPW + (gx - fx0);
                ring_idx = gy * tx + gx;
                if (SLAB && (gy < P.S.q_lo || gy >= P.S.q_hi)) {  // owned by a neighbouring band: inbox row slot
                    const int slot = gy < P.S.q_lo ? gy - (P.S.q_lo - P.S.reach_lo) : P.S.reach_lo + gy - P.S.q_hi;
                    ring_idx = -1 - (slot * tx + gx);
                }
            }
        };
        halo_cell(tid, halo_fp, halo_ring);
        for (int h = tid + kChThreads; h < n_halo; h += kChThreads) {
            int f, g;
            halo_cell(h, f, g);
            s_over_fp[h - kChThreads] = (unsigned short) (f < 0 ? 0xffff : f);
        }
    }
    auto ring_of = [&](int f) {  // ring / inbox index of a footprint cell (what halo_cell returns), recomputed for the overflow list
        const int gx = fx0 + f % FPW, gy = fy0 + f / FPW;
        if (SLAB && (gy < P.S.q_lo || gy >= P.S.q_hi)) {
            const int slot = gy < P.S.q_lo ? gy - (P.S.q_lo - P.S.reach_lo) : P.S.reach_lo + gy - P.S.q_hi;
            return -1 - (slot * tx + gx);
        }
        return gy * tx + gx;
    };
    const int n_over = max(0, n_halo - kChThreads);
    // SLAB: cells a neighbouring band reads go to its inbox as well (row slot in the RECEIVER's numbering)
    bool xlo = false, xhi = false;
    int xlo_idx = 0, xhi_idx = 0;
    if (SLAB) {
        xlo = P.S.out_lo != nullptr && py - P.S.q_lo < P.S.reach_hi;
        xhi = P.S.out_hi != nullptr && P.S.q_hi - 1 - py < P.S.reach_lo;
        xlo_idx = (P.S.reach_lo + py - P.S.q_lo) * tx + px;    // the lower neighbour's q_hi is our q_lo
        xhi_idx = (py - (P.S.q_hi - P.S.reach_lo)) * tx + px;  // the upper neighbour's q_lo is our q_hi
    }
    const size_t inbox_plane = SLAB ? (size_t) (P.S.reach_lo + P.S.reach_hi) * tx : 0;
    unsigned long long* const ring = (unsigned long long*) P.ring;
    const unsigned int tag_base = P.epoch << 16;
    const int light_off = lx * P.ls_p + row * P.ls_q;  // this thread's first pixel inside a light brick
    __syncthreads();
    const int ndown = s_ndown;

    if (SLAB && k_begin > 0 && P.S.zin != nullptr) {
        // the slices before k_begin belong to the upstream slab: its last slice (tag k_begin) is our "slice k_begin - 1"
        float* fp0 = s_fp + (k_begin & 1) * (FPW * kFpH);
        const unsigned int want = tag_base + (unsigned) k_begin;
        for (int c = tid; c < FPW * kFpH; c += kChThreads) {
            const int gx = fx0 + c % FPW, gy = fy0 + c / FPW;
            if ((unsigned) gx < (unsigned) tx && (unsigned) gy < (unsigned) ty) {
                const unsigned long long v = chain_poll_cell(P.S.zin + (size_t) gy * tx + gx, want, 1, P.timeout_ns, &s_abort, P.error);
                fp0[c] = __uint_as_float((unsigned int) v);
            }
        }
        __syncthreads();
    }

    // Reading a ring cell another SM wrote a moment ago costs ~1 200 cycles here (measured with in-kernel timers: a strong GPU-scope
    // load of such a line, 4-5 times the L2 hit latency) — more than all the work of a slice. The halo cell slice k+2 reads (slice k+1 of
    // the upstream tile) is therefore requested during slice k, as an asynchronous copy into a 3-slot SMEM queue. A request that comes too
    // early (the upstream tile has not exported yet) delivers a stale tag and falls into the polling path once; that delays this tile
    // until it lags its upstream neighbour by about two slices, from where on every request hits. Same idea for the back-pressure probe
    // (a progress word only grows: an early value that suffices stays valid).
    //
    // The loop below is written for instruction count: a slice costs its per-thread bookkeeping, not its arithmetic (24 fp32 operations
    // for 4 pixels), and the pass runs at slices x (instructions of one warp per slice) x (issue interval of a warp). Blocks are whole
    // (the host requires slices % kSB == 0), so the slice index within a block is a compile-time constant: block-level work (TMA issue,
    // probes) sits in fixed iterations, addresses are a base pointer plus (k mod 16) x plane.
    const bool has_halo = halo_fp >= 0;
    const bool halo_inbox = SLAB && halo_ring < 0;
    const unsigned long long* const halo_base = halo_inbox ? P.S.inbox + (-1 - halo_ring) : ring + (unsigned int) halo_ring;
    const bool halo_odd = ((halo_inbox ? -1 - halo_ring : halo_ring) & 1) != 0;  // which half of its 16-byte pair the cell is (planes hold an even number of cells)
    const unsigned long long* const halo_pair = halo_base - (halo_odd ? 1 : 0);
    const unsigned int halo_stride = halo_inbox ? (unsigned int) inbox_plane : plane32;  // cells per slice (the inbox is full depth, the ring wraps)
    const unsigned int halo_wrap = halo_inbox ? 0xffffffffu : (unsigned int) (kRingDepth - 1);
    float* const halo_dst = s_fp + (has_halo ? halo_fp : 0);
    unsigned long long* const exp_base = ring + own_cell;
    unsigned int* const my_flag = P.flags + (size_t) tile * kFlagStride;
    const unsigned int* const probe_flag = P.flags + (size_t) s_down[tid < ndown ? tid : 0] * kFlagStride;
    const bool prober = tid < ndown;
    const float* const tap0 = s_fp + tap_idx;
    float* const own0 = s_fp + own_idx;
    const int t_step = (U.dirn > 0 ? 1 : -1) * (kTH * TW);  // T / light slices are visited in sweep order
    const int t_first = (U.dirn > 0 ? 0 : kSB - 1) * (kTH * TW);
#ifdef TBRM_CHAIN_TIMERS
    long long tsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define TBRM_T(i) { const long long now_ = clock64(); tsum[i] += now_ - tlast; tlast = now_; }
    long long tlast = clock64();
#else
#define TBRM_T(i)
#endif
    unsigned int pv_next = 0;
    for (int i = 0; i < kChHaloDepth; ++i) s_hstage[i][tid] = make_ulonglong2(0ull, 0ull);  // a tag that matches nothing
    int store_pending = -1;  // block whose light brick is complete in SMEM and waits for its TMA store (thread 0)
    for (int n = 0; n < nblk; ++n) {
        const int tst = n % kChTStages, lst = n % kChLStages;
        const int s0 = block_s0(b_begin + n);
        const unsigned int k0 = (unsigned int) (b_begin + n) * kSB;  // sweep position of the block's first slice
        float* s_light = s_lstage + lst * kBrick;
        mbar_wait(&t_full[tst], (uint32_t) ((n / kChTStages) & 1));
        TBRM_T(6)
        const bool t_ones = s_tones[tst] != 0;
        // a brick of ones was not loaded: every slice reads the same vector of ones
        const float* tptr = s_tstage + tst * kBrick + row * TW + lx + t_first;
        const int tstep = t_step;
        float cur[kSB][CPX];
        const bool combine = P.mode == kModeCombine;
        const bool pix_ok = vy && px < tx;
        int c3[3];
        c3[PA] = px, c3[QA] = py, c3[SA] = s0;
        // the removed light's values (kModeCombine) are read at the end of the block: start fetching them now
        const float* rp = P.scratch + ((size_t) c3[0] + (size_t) U.ldims[0] * ((size_t) c3[1] + (size_t) U.ldims[1] * (size_t) c3[2]));
        const size_t rss = AXIS == 1 ? (size_t) U.ldims[0] : (size_t) U.ldims[0] * U.ldims[1];  // scratch stride of a slice (AXIS 1, 2)
        if (combine && pix_ok) {
            if (AXIS == 0) {
#pragma unroll
                for (int i = 0; i < CPX; ++i)
                    if (px + i < tx) prefetch_l1(rp + (size_t) i * U.ldims[0]);
            } else {
#pragma unroll
                for (int v = 0; v < kSB; ++v) prefetch_l1(rp + (size_t) v * rss);
            }
        }

#pragma unroll
        for (int v = 0; v < kSB; ++v) {
            const unsigned int k = k0 + v;  // position in sweep order
            const int fp_par = (int) (k & 1u) * (FPW * kFpH);
            // ---- (a) the halo cell of slice k-1 was requested a slice ago; request the one slice k+1 will read (slice k, tag k+1) ----
            const unsigned int want_tag = tag_base + k;  // slice k-1 carries tag k
            // the copy issued two slices ago has landed (the one issued in the previous slice may still be in flight)
            unsigned long long hv = 0;
            if (has_halo) {
                cp_async_wait_but_one();
                const ulonglong2 pair = s_hstage[k % kChHaloDepth][tid];
                hv = halo_odd ? pair.y : pair.x;
            }
            // a ring slot is reused every kRingDepth slices: before exporting slices k..k+3 every reader must have consumed slice
            // k+3-kRingDepth, i.e. passed the barrier of its slice k+4-kRingDepth. Probed in the first slice of a block, read in the last
            // slice of the block before.
            const unsigned int pv = pv_next;
            if (v == kSB - 1 && prober && k + 5 > kRingDepth) pv_next = ld_relaxed_u32(probe_flag);
            float tv[CPX];
            if (t_ones) {
#pragma unroll
                for (int i = 0; i < CPX; ++i) tv[i] = 1.0f;
            } else if constexpr (CPX == 4) {
                const float4 t4 = *(const float4*) (tptr + v * tstep);
                tv[0] = t4.x, tv[1] = t4.y, tv[2] = t4.z, tv[3] = t4.w;
            } else {
                const float2 t2 = *(const float2*) (tptr + v * tstep);
                tv[0] = t2.x, tv[1] = t2.y;
            }
            TBRM_T(0)
            // ---- (c) the halo of slice k-1 must have arrived; readers of the ring slots we overwrite must have moved on ----
            if (has_halo && k > (unsigned int) k_begin) {
                if ((unsigned int) (hv >> 32) != want_tag)
                    hv = chain_poll_cell(halo_base + (size_t) (((k - 1u) & halo_wrap) * halo_stride), want_tag, halo_inbox, P.timeout_ns, &s_abort, P.error);
                halo_dst[fp_par] = __uint_as_float((unsigned int) hv);
            }
            if (n_over > 0 && k > (unsigned int) k_begin) {  // footprints shifted far off the tile (rare): the cells beyond one per thread
                const unsigned int rd_slot = ((k + kRingDepth - 1u) % kRingDepth) * plane32;
                const unsigned long long* rdi = SLAB ? P.S.inbox + (size_t) (k - 1u) * inbox_plane : nullptr;
                for (int h = tid; h < n_over; h += kChThreads) {
                    const int f = s_over_fp[h];
                    if (f == 0xffff) continue;
                    const int g = ring_of(f);
                    const unsigned long long* cell = (SLAB && g < 0) ? rdi + (-1 - g) : ring + (rd_slot + (unsigned int) g);
                    s_fp[fp_par + f] = __uint_as_float((unsigned int) chain_poll_cell(cell, want_tag, SLAB && g < 0, P.timeout_ns, &s_abort, P.error));
                }
            }
            if (v == 0 && prober && k + 4 > kRingDepth && pv < k + 4 - kRingDepth)
                chain_poll_flag(probe_flag, k + 4 - kRingDepth, P.timeout_ns, &s_abort, P.error);
            TBRM_T(1)
            __syncthreads();
            TBRM_T(2)
            // request the cell slice k+2 will read (slice k+1 of the upstream tile, tag k+2): an asynchronous 16-byte copy L2 -> SMEM, so
            // that no register scoreboard ties the check of an older request to a younger one
            if (has_halo) {
                if (k + 2 < (unsigned int) k_end)
                    cp_async_16(&s_hstage[(k + 2) % kChHaloDepth][tid], halo_pair + (size_t) (((k + 1u) & halo_wrap) * halo_stride));
                cp_async_commit();
            }
            if (tid == kChService) {
                st_relaxed_u32(my_flag, k);  // every read of slice k-1 by this tile is done
                if (v == 0 && store_pending >= 0) {  // the previous block: all its brick updates and T reads happened before the barrier above
                    const int ps0 = block_s0(b_begin + store_pending);
                    int lc[3];
                    lc[PA] = x0, lc[QA] = y0, lc[SA] = ps0;
                    tma_store_3d(&light_map, lc[0], lc[1], lc[2], s_lstage + (store_pending % kChLStages) * kBrick);
                    tma_commit();
                    if (store_pending + kChTStages < nblk) issue_t(store_pending + kChTStages);  // the block that reuses its T stage
                }
                // the light brick of the block that reuses that stage may be loaded once the store (issued a slice ago) has read it
                if (v == 1 && store_pending >= 0) {
                    tma_wait_read<0>();
                    if (store_pending + kChLStages < nblk) issue_light(store_pending + kChLStages);
                    store_pending = -1;
                }
            }
            TBRM_T(3)
            // ---- (d) propagate, forward, export ----
            {
                const float* r0 = tap0 + fp_par;
                const float* r1 = r0 + FPW;
                float ta[CPX + 1], tb[CPX + 1];
#pragma unroll
                for (int i = 0; i <= CPX; ++i) ta[i] = r0[i], tb[i] = r1[i];
                float* fp_own = own0 + (FPW * kFpH - fp_par);  // the other buffer: the footprint slice k+1 reads
#pragma unroll
                for (int i = 0; i < CPX; ++i) {
                    const float prev = lerpf(lerpf(ta[i], ta[i + 1], bfx[i]), lerpf(tb[i], tb[i + 1], bfx[i]), bfy);
                    cur[v][i] = prev * tv[i];
                }
#pragma unroll
                for (int i = 0; i < CPX; ++i)
                    if ((ownmask >> i) & 1u) fp_own[i] = cur[v][i];
                const unsigned long long tag = (unsigned long long) (tag_base + k + 1u) << 32;
                if (expmask) {
                    unsigned long long* wr = exp_base + (size_t) ((k % kRingDepth) * plane32);
#pragma unroll
                    for (int i = 0; i < CPX; ++i)
                        if ((expmask >> i) & 1u) st_relaxed_u64(wr + i, tag | __float_as_uint(cur[v][i]));
                }
                if (SLAB) {
                    const size_t ko = (size_t) k * inbox_plane;
#pragma unroll
                    for (int i = 0; i < CPX; ++i) {
                        if (!((vmask >> i) & 1u)) continue;
                        if (xlo) st_relaxed_sys_u64(P.S.out_lo + ko + xlo_idx + i, tag | __float_as_uint(cur[v][i]));
                        if (xhi) st_relaxed_sys_u64(P.S.out_hi + ko + xhi_idx + i, tag | __float_as_uint(cur[v][i]));
                    }
                    if (P.S.zout != nullptr && k == (unsigned int) (k_end - 1)) {  // hand the last slice of this slab to the next one
                        unsigned long long* zo = P.S.zout + (size_t) px + (size_t) tx * py;
#pragma unroll
                        for (int i = 0; i < CPX; ++i)
                            if ((vmask >> i) & 1u) st_relaxed_sys_u64(zo + i, tag | __float_as_uint(cur[v][i]));
                    }
                }
            }
            TBRM_T(4)
        }
        // ---- fold the block's propagated light into the light brick ----
        mbar_wait(&l_full[lst], (uint32_t) ((n / kChLStages) & 1));  // requested two blocks ago
        {
            // pixels beyond a ragged plane edge update their cells of the SMEM brick too: the TMA store clips the brick to the volume
            auto upd = [&](float& L, float c, float r) {
                if (P.mode == kModeAdd) {  // AddDirLightShader.usf:121-126
                    if (fabsf(c) > 1e-3f) L = L + (c * U.sign);
                } else if (P.mode == kModeStore) {  // the removed light of a ChangeDirLight: its light goes to the scratch volume
                    L = c;
                } else {  // the added light of a ChangeDirLight: LightVolume += added - removed (ChangeDirLightShader.usf:146-153)
                    if (fabsf(c - r) > 1e-3f) L = L + c - r;
                }
            };
            if (AXIS == 0) {  // brick order (s, p, q): one 16-byte vector per pixel holds its kSB slices
#pragma unroll
                for (int i = 0; i < CPX; ++i) {
                    float* lp = s_light + light_off + i * P.ls_p;
                    float4 L4 = *(float4*) lp;
                    float4 R4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (combine && pix_ok && px + i < tx) R4 = __ldg((const float4*) (rp + (size_t) i * U.ldims[0]));  // ns = X is a multiple of 4
                    if (U.dirn > 0) {
                        upd(L4.x, cur[0][i], R4.x), upd(L4.y, cur[1][i], R4.y), upd(L4.z, cur[2][i], R4.z), upd(L4.w, cur[3][i], R4.w);
                    } else {
                        upd(L4.w, cur[0][i], R4.w), upd(L4.z, cur[1][i], R4.z), upd(L4.y, cur[2][i], R4.y), upd(L4.x, cur[3][i], R4.x);
                    }
                    *(float4*) lp = L4;
                }
            } else {  // brick rows run along p = x: one vector per slice holds the thread's pixels
#pragma unroll
                for (int v = 0; v < kSB; ++v) {
                    const int slot = U.dirn > 0 ? v : kSB - 1 - v;
                    float* lp = s_light + light_off + slot * P.ls_s;
                    const bool r_ok = combine && pix_ok;  // tx = X is a multiple of 4: the vector is inside or outside as a whole
                    if constexpr (CPX == 4) {
                        float4 L4 = *(float4*) lp;
                        float4 R4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (r_ok) R4 = __ldg((const float4*) (rp + (size_t) slot * rss));
                        upd(L4.x, cur[v][0], R4.x), upd(L4.y, cur[v][1], R4.y), upd(L4.z, cur[v][2], R4.z), upd(L4.w, cur[v][3], R4.w);
                        *(float4*) lp = L4;
                    } else {
                        float2 L2 = *(float2*) lp;
                        float2 R2 = make_float2(0.f, 0.f);
                        if (r_ok) R2 = __ldg((const float2*) (rp + (size_t) slot * rss));
                        upd(L2.x, cur[v][0], R2.x), upd(L2.y, cur[v][1], R2.y);
                        *(float2*) lp = L2;
                    }
                }
            }
        }
        fence_async_smem();  // the updates must be visible to the TMA store issued after the next barrier
        store_pending = n;
        TBRM_T(5)
    }
#ifdef TBRM_CHAIN_TIMERS
    if (P.dbg != nullptr && (tid & 31) == 0)
        for (int i = 0; i < 8; ++i) P.dbg[((size_t) blockIdx.x * 4 + (tid >> 5)) * 8 + i] = tsum[i];
#endif
    __syncthreads();
    if (tid == kChService) {
        if (store_pending >= 0) {
            const int ps0 = block_s0(b_begin + store_pending);
            int lc[3];
            lc[PA] = x0, lc[QA] = y0, lc[SA] = ps0;
            tma_store_3d(&light_map, lc[0], lc[1], lc[2], s_lstage + (store_pending % kChLStages) * kBrick);
            tma_commit();
        }
        tma_wait_all<0>();
    }
}

}  // namespace tbrm
