// sweep_tma_kernel.cuh — device code of the TMA-staged fused plane-sweep (included by sweep_tma.cuh).
#pragma once

namespace tbrm {

// 64-bit ring cell = (tag << 32) | float bits. One aligned 64-bit store / load is single-copy atomic, so a reader that
// sees the expected tag also sees the matching value: no fence and no separate flag on the critical path.
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// cells another GPU writes over NVLink (slab exchange) are accessed at system scope
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// correctly rounded v/255 for v in 0..255 without a division: with 1/255 = c_hi + c_lo (c_hi = RN(1/255)),
// fma(v, c_hi, v*c_lo) == RN(v/255) for every byte (exhaustive check in tests/test_host_cpu.py)
__device__ __forceinline__ float decode_u8(uint32_t v) {
    const float x = (float) v;
    return __fmaf_rn(x, 0.003921568859368563f, x * -2.319175823606301e-10f);
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
//
// Per slice k every thread
//  (a) issues the loads of the <= 2 halo cells it fetches from the L2 ring (slice k-1 of the upstream tiles);
//  (b) computes the opacity toward the light of its two voxels from the TMA-staged data brick — the bulk of the work,
//      independent of the previous slice, which hides the ring latency;
//  (c) checks the halo tags (re-polls only while the upstream tile is not yet ahead) and parks the values in SMEM;
//  (d) after ONE block barrier propagates: previous-slice taps from SMEM (own tile forwarded through SMEM, halo from the
//      ring), extinction, export of the cells other tiles read, accumulation into the light brick in SMEM.
//
// SLAB = true adds what a launch needs when it covers only part of the pass (SlabParams): a band of tile rows of the buffer
// plane (the Z-slab of a GPU for sweeps along X / Y, or one co-resident wave of a plane too large for the GPU) and / or a
// sub-range of the slices (the Z-slab of a GPU for sweeps along Z). Footprint rows owned by a neighbouring band arrive in a
// full-depth inbox of LL cells (written by that band's launch: an earlier launch on this GPU, or the neighbour GPU's
// concurrent launch through NVLink peer stores); the last slice of a slice sub-range is handed to the next slab as a
// plane of LL cells. The per-voxel arithmetic is untouched, so a sharded pass is bit-identical to an unsharded one.
//
// PX = pixels per thread along p: 2 (tile 64 x 8, the two pixels share their middle tap column) or 1 (tile 32 x 8). A pass costs
// slices x the per-slice instruction chain of ONE tile once its tiles no longer fill the SMs (256^3, the slab of a sharded volume):
// one pixel per thread shortens that chain by the second pixel's work and doubles the tiles; the host picks it for such launches.
// L8 = the light volume is G8 (UNORM8, the reference's default: RaymarchVolume.h:198-199, RaymarchVolume.cpp:857-861): the light brick is a
// byte brick (load -> v / 255, store -> floor(saturate(v) * 255 + 0.5)), and what a slice forwards to the next one is the value its G8
// read / write buffer would hold (the propagation buffers have the light volume's pixel format) — the light volume itself is updated with the
// unquantised value, as in the shader. ChangeDirLight: the removed light's launch keeps an R32F brick (the scratch volume) and only quantises
// what it forwards; the added light's launch has the byte brick with the removed light's R32F brick behind it. A byte brick of 4 slices along
// X has 4-byte rows (below TMA's 16): a sweep along X works on a (y,z,x)-ordered copy of the light volume (the host permutes it there and
// back: three byte transposes, ~0.3 ms at 512^3).
// TH = rows of a tile (= warps of a block): 8, or 7 where 64 x 7 tiles fill every SM with the same number of blocks (the host's choice of th).
// A half-resolution light volume (P.dk = 2 data voxels per light voxel) runs in the one-pixel form: the data box of a tile starts at
// dk * the tile's origin and spans dk * its extent.
template <int AXIS, bool CLIP, bool SLAB, int PX, bool L8 = false, int TH = 8>
__global__ void __launch_bounds__(32 * TH, 4)
    sweep_tma_kernel(const __grid_constant__ CUtensorMap light_map, const __grid_constant__ CUtensorMap data_map,
                     const __grid_constant__ CUtensorMap scratch_map, const __grid_constant__ PushMaps push_maps, const TmaParams P,
                     const float4* __restrict__ tf) {
    constexpr int PA = (AXIS == 0) ? 1 : 0;  // native axis of p
    constexpr int QA = (AXIS == 2) ? 1 : 2;  // native axis of q
    constexpr int SA = AXIS;                 // native axis of s
    constexpr int TW = 32 * PX, FPW = TW + 4;  // tile width, width of the SMEM footprint of the previous slice
    static_assert(PX == 1 || PX == 2, "one or two pixels per thread");
    const SweepUniforms& U = P.U;
    const int tx = U.td[0], ty = U.td[1], ns = U.td[2];
    const int tid = threadIdx.x;
    // tile height (rows = warps of the block): 8, or 7 where that fills every SM with the same number of tiles (see the host's choice of th)
    constexpr int kTH = TH, kTmaThreads = 32 * TH, kHaloOverflow = kFpW * kFpH - kHaloPerThread * kTmaThreads;  // shadow the constants of the 8-row tile
    static_assert(TH == 7 || TH == 8, "tile rows");
    const int tix = (int) blockIdx.x % P.ntx, tiy = (SLAB ? P.S.tile_row0 : 0) + (int) blockIdx.x / P.ntx;
    const int tile = tiy * P.ntx + tix;
    const int row_lo = SLAB ? P.S.tile_row0 : 0, row_hi = SLAB ? P.S.tile_row0 + P.S.tile_rows : P.nty;  // tile rows of this launch
    const int k_begin = SLAB ? P.S.k_begin : 0, k_end = SLAB ? P.S.k_end : ns;
    const int x0 = tix * TW, y0 = P.q_org + tiy * kTH;  // (q_org: tile rows are anchored at the first row of the launch's slab)
    const int q_stop = SLAB ? min(U.td[1], P.S.q_hi) : U.td[1];  // rows from here on are not this launch's (a slab may end inside its last tile row)
    const size_t plane = (size_t) tx * ty;
    const unsigned int plane32 = (unsigned int) (tx * ty);  // ring cells are indexed in 32 bits (the host checks kRingDepth * plane < 2^31)

    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* stage_base = smem;
    float* s_fp = (float*) (smem + (size_t) kStages * P.stage_bytes);  // 2 x footprint (ping-pong over slices)
    float* s_alpha = s_fp + 2 * FPW * kFpH;
    uint64_t* s_bar = (uint64_t*) (s_alpha + 256);
    __shared__ int s_down[kFusedMaxDeps];
    __shared__ int s_ndown;
    __shared__ volatile int s_abort;  // SLAB: an exchange with another launch timed out; stop waiting (results are void)

    const int fx0 = x0 + P.bmin[0], fy0 = y0 + P.bmin[1], FW = P.bext[0], FH = P.bext[1];
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) mbar_init(&s_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // tiles that read cells of ours: tile (i,j) reads [i*TW + bmin, +FW) x [j*TH + bmin, +FH)
        int ndown = 0;
        s_abort = 0;
        for (int j = row_lo; j < row_hi; ++j) {
            const int gy = P.q_org + j * kTH + P.bmin[1];
            if (gy + FH <= y0 || gy >= y0 + kTH) continue;
            for (int i = 0; i < P.ntx; ++i) {
                const int gx = i * TW + P.bmin[0];
                if (gx + FW <= x0 || gx >= x0 + TW) continue;
                if ((i != tix || j != tiy) && ndown < kFusedMaxDeps) s_down[ndown++] = j * P.ntx + i;
            }
        }
        s_ndown = ndown;
    }
    for (int c = tid; c < 256; c += kTmaThreads) s_alpha[c] = __ldg(&tf[c]).w;  // (one per thread for 8 rows)
    // footprint buffers: out-of-plane cells hold the sampler border colour for good, in-plane cells start as the
    // cleared buffer (LightAlpha) = "slice -1"
    for (int c = tid; c < FPW * kFpH; c += kTmaThreads) {
        const int gx = fx0 + c % FPW, gy = fy0 + c / FPW;
        const bool in = (unsigned) gx < (unsigned) tx && (unsigned) gy < (unsigned) ty;
        const float v = in ? (L8 ? decode_u8((uint32_t) quant8(U.a.light_alpha)) : U.a.light_alpha) : U.a.border;
        s_fp[c] = v;
        s_fp[FPW * kFpH + c] = v;
    }

    // ---- per-thread invariants: 2 adjacent pixels (px, px+1) of row py --------------------------------------
    const int lx = (tid & 31) * PX, ly = tid >> 5;
    const int px = x0 + lx, py = y0 + ly;
    const bool v0 = px < tx && py < q_stop, v1 = PX == 2 && px + 1 < tx && py < q_stop;  // PX == 1: every use of the second pixel folds away
    const int pxc = min(px, tx - 1), px1c = min(px + 1, tx - 1), pyc = min(py, ty - 1);
    const int2 mp0 = __ldg(&P.A.ax[PA].meta[pxc]), mp1 = __ldg(&P.A.ax[PA].meta[px1c]), mq = __ldg(&P.A.ax[QA].meta[pyc]);
    const float fp0 = __ldg(&P.A.ax[PA].f[pxc]), fp1 = __ldg(&P.A.ax[PA].f[px1c]), fq = __ldg(&P.A.ax[QA].f[pyc]);
    const int dN_p = P.data_dims_t[0], dN_q = P.data_dims_t[1], dN_s = P.data_dims_t[2];
    const int col = mp0.x - (x0 * P.dk[0] + P.dmin[0]);  // column of the first tap inside the data box
    const int rowq = mq.x - (y0 * P.dk[1] + P.dmin[1]);
    const bool inP0 = (unsigned) mp0.x < (unsigned) dN_p, inP1 = (unsigned) (mp0.x + 1) < (unsigned) dN_p,
               inP2 = PX == 1 || (unsigned) (mp0.x + 2) < (unsigned) dN_p;  // the third tap column belongs to the second pixel
    const bool inQ0 = (unsigned) mq.x < (unsigned) dN_q, inQ1 = (unsigned) (mq.x + 1) < (unsigned) dN_q;
    const bool all_pq = inP0 && inP1 && inP2 && inQ0 && inQ1;
    // AddDirLight samples only where GetUVW + UVWOffset is inside [0,1]^3 (AddDirLightShader.usf:110); ChangeDirLight has no such
    // gate and relies on the border sampler (ChangeDirLightShader.usf:130,136)
    const bool gate = U.gate_saturate != 0;
    const bool inside_pq0 = v0 && (!gate || (mp0.y && mq.y)), inside_pq1 = v1 && (!gate || (mp1.y && mq.y));
    float Sp0 = 0.f, Sp1 = 0.f, Sq = 0.f;
    if (CLIP) {
        Sp0 = __ldg(&P.A.ax[PA].S[pxc]), Sp1 = __ldg(&P.A.ax[PA].S[px1c]), Sq = __ldg(&P.A.ax[QA].S[pyc]);
    }
    const int2 bxa = __ldg(&P.A.bx[pxc]), bxb = __ldg(&P.A.bx[px1c]), bya = __ldg(&P.A.by[pyc]);
    const float bfx0 = __int_as_float(bxa.y), bfx1 = __int_as_float(bxb.y), bfy = __int_as_float(bya.y);
    const int tap_idx = (bya.x - fy0) * FPW + (bxa.x - fx0);  // first of the 3 x 2 read-buffer taps
    // where this thread's own output lives in the footprint (if the footprint covers it)
    const bool own_in0 = v0 && px >= fx0 && px < fx0 + FW && py >= fy0 && py < fy0 + FH;
    const bool own_in1 = v1 && px + 1 >= fx0 && px + 1 < fx0 + FW && py >= fy0 && py < fy0 + FH;
    const int own_idx = (py - fy0) * FPW + (px - fx0);
    const unsigned int own_cell = (unsigned int) (px + tx * py);  // this thread's first pixel in a ring slice (used where v0 / v1 hold)
    // is a pixel read by another tile? tile (i,j) reads [i*TW + bmin, +FW) x [j*TH + bmin, +FH)
    auto exported = [&](int gx, int gy) {
        const int rx = gx - P.bmin[0], ry = gy - P.bmin[1] - P.q_org;  // tile origin i*TW must lie in (rx - FW, rx]
        const int ia = max(0, (rx - FW + TW) / TW), ib = rx >= 0 ? min(P.ntx - 1, rx / TW) : -1;
        const int ja = max(row_lo, (ry - FH + kTH) / kTH), jb = ry >= 0 ? min(row_hi - 1, ry / kTH) : -1;
        for (int j = ja; j <= jb; ++j)
            for (int i = ia; i <= ib; ++i) {
                const int gx0 = i * TW + P.bmin[0], gy0 = P.q_org + j * kTH + P.bmin[1];
                if ((i != tix || j != tiy) && gx >= gx0 && gx < gx0 + FW && gy >= gy0 && gy < gy0 + FH) return true;
            }
        return false;
    };
    const bool exp0 = v0 && exported(px, py), exp1 = v1 && exported(px + 1, py);
    // halo cells this thread fetches: footprint cells inside the plane that belong to other tiles. Cell h of the enumeration
    // (rows above the own tile, left / right of it, rows below) goes to thread h % 256; the first two of a thread live in
    // registers and are prefetched behind the occlusion work, the rest (footprints shifted far off the tile: a second-axis
    // pass of a light close to its main axis) are described in SMEM and fetched in a plain loop.
    int halo_fp[kHaloPerThread], halo_ring[kHaloPerThread];
    __shared__ int s_over_fp[kHaloOverflow], s_over_ring[kHaloOverflow];
    int n_halo;  // footprint cells outside the own tile
    {
        // (rows of the tile past the end of the slab belong to the neighbour: they arrive through the inbox like any other halo cell)
        const int ox0 = max(fx0, x0), ox1 = min(fx0 + FW, x0 + TW), oy0 = max(fy0, y0), oy1 = min(fy0 + FH, SLAB ? min(y0 + kTH, P.S.q_hi) : y0 + kTH);
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
                    // rows only the masked pixels of a last tile row would read (a slab that ends inside the row) are nobody's to deliver
                    if (slot < 0 || slot >= P.S.reach_lo + P.S.reach_hi) fp_idx = -1;
                }
            }
        };
#pragma unroll
        for (int i = 0; i < kHaloPerThread; ++i) halo_cell(tid + i * kTmaThreads, halo_fp[i], halo_ring[i]);
        for (int h = tid + kHaloPerThread * kTmaThreads; h < n_halo; h += kTmaThreads) {
            int f, g;
            halo_cell(h, f, g);
            s_over_fp[h - kHaloPerThread * kTmaThreads] = f, s_over_ring[h - kHaloPerThread * kTmaThreads] = g;
        }
    }
    const int n_over = max(0, n_halo - kHaloPerThread * kTmaThreads);
    // SLAB: cells a neighbouring band reads go to its inbox as well (row slot in the RECEIVER's numbering)
    bool xlo0 = false, xlo1 = false, xhi0 = false, xhi1 = false;
    int xlo_idx = 0, xhi_idx = 0;
    if (SLAB) {
        const bool lo = P.S.out_lo != nullptr && py - P.S.q_lo < P.S.reach_hi, hi = P.S.out_hi != nullptr && P.S.q_hi - 1 - py < P.S.reach_lo;
        xlo0 = lo && v0, xlo1 = lo && v1, xhi0 = hi && v0, xhi1 = hi && v1;
        xlo_idx = (P.S.reach_lo + py - P.S.q_lo) * tx + px;      // the lower neighbour's q_hi is our q_lo
        xhi_idx = (py - (P.S.q_hi - P.S.reach_lo)) * tx + px;    // the upper neighbour's q_lo is our q_hi
    }
    const size_t inbox_plane = SLAB ? (size_t) (P.S.reach_lo + P.S.reach_hi) * tx : 0;
    const float rwidth = 1.0f / U.win.width;
    const float step = U.a.step;
    const int shift = (col & 3) * 8;
    const int thread_data = rowq * P.ds_q + (col & ~3);
    const int light_off = lx * P.ls_p + ly * P.ls_q;
    unsigned long long* const ring = (unsigned long long*) P.ring;
    const unsigned int tag_base = P.epoch << 16;
    __syncthreads();
    const int ndown = s_ndown;
    // give up on an exchange that does not complete (a peer that died or was never launched): flag it, stop waiting
    // Every wait on another tile / launch gives up: waits of partial launches (another GPU may never answer) by the clock, waits inside an
    // unsharded launch by counting polls (an L2 round trip takes well over 256 ns, so poll_limit = timeout / 256 ns bounds the wait by a few
    // timeouts without a timer read or a call in the loop — round 1's unsharded waits had no bound at all)
    auto timed_out = [&](unsigned long long t0) {
        if (global_timer_ns() - t0 < P.timeout_ns) return false;
        s_abort = 1;
        atomicExch(P.error, 1u);
        return true;
    };
    auto give_up = [&]() {
        s_abort = 1;
        atomicExch(P.error, 1u);
    };
    if (SLAB && k_begin > 0 && P.S.zin != nullptr) {
        // the slices before k_begin belong to the upstream slab: its last slice (tag k_begin) is our "slice k_begin - 1"
        float* fp0 = s_fp + (k_begin & 1) * (FPW * kFpH);
        const unsigned int want = tag_base + (unsigned) k_begin;
        for (int c = tid; c < FPW * kFpH; c += kTmaThreads) {
            const int gx = fx0 + c % FPW, gy = fy0 + c / FPW;
            if ((unsigned) gx < (unsigned) tx && (unsigned) gy < (unsigned) ty) {
                const unsigned long long* cell = P.S.zin + (size_t) gy * tx + gx;
                unsigned long long v = ld_relaxed_sys_u64(cell);
                const unsigned long long t0 = global_timer_ns();
                unsigned int polls = 0;
                while ((unsigned int) (v >> 32) != want && !s_abort) {
                    v = ld_relaxed_sys_u64(cell);
                    if ((++polls & 255u) == 0 && timed_out(t0)) break;
                }
                fp0[c] = __uint_as_float((unsigned int) v);
            }
        }
        __syncthreads();
    }

    const int nblocks = (ns + kSB - 1) / kSB;
    const int b_begin = k_begin / kSB, b_end = (k_end + kSB - 1) / kSB;  // SLAB slice ranges are multiples of kSB (host)
    // blocks are aligned to multiples of kSB in native coordinates (TMA: 16-byte aligned inner coordinate when the
    // sweep axis is x); a descending sweep visits them last-to-first
    auto block_s0 = [&](int b) { return (U.dirn > 0 ? b : nblocks - 1 - b) * kSB; };
    auto issue_load = [&](int b) {
        const int st = (b - b_begin) % kStages;
        unsigned char* sb = stage_base + (size_t) st * P.stage_bytes;
        mbar_expect_tx(&s_bar[st], (uint32_t) (P.data_off + P.data_bytes));
        const int s0 = block_s0(b);
        int lc[3], nc[3], dc[3];
        nc[PA] = x0, nc[QA] = y0, nc[SA] = s0;  // light and scratch maps are over native (x,y,z)
        lc[0] = nc[0], lc[1] = nc[1], lc[2] = nc[2];
        if (L8 && AXIS == 0 && P.mode != kModeStore) lc[0] = x0, lc[1] = y0, lc[2] = s0;  // ... except the byte brick of a G8 volume's sweep along X: its (y,z,x)-ordered copy, i.e. (p,q,s)
        tma_load_3d(sb, &light_map, lc[0], lc[1], lc[2], &s_bar[st]);
        if (P.mode == kModeCombine) tma_load_3d(sb + P.light_bytes, &scratch_map, nc[0], nc[1], nc[2], &s_bar[st]);
        // data map: native dims for Z / Y sweeps, the (y,z,x) replica for X sweeps, i.e. (p,q,s)-ordered for X
        if (AXIS == 0) {
            tma_load_3d(sb + P.data_off, &data_map, x0 * P.dk[0] + P.dmin[0], y0 * P.dk[1] + P.dmin[1], s0 * P.dk[2] + P.dmin[2], &s_bar[st]);
        } else {
            dc[PA] = x0 * P.dk[0] + P.dmin[0], dc[QA] = y0 * P.dk[1] + P.dmin[1], dc[SA] = s0 * P.dk[2] + P.dmin[2];
            tma_load_3d(sb + P.data_off, &data_map, dc[0], dc[1], dc[2], &s_bar[st]);
        }
    };
    if (tid == 0) issue_load(b_begin);

    for (int b = b_begin; b < b_end; ++b) {
        const int st = (b - b_begin) % kStages;
        if (tid == 0 && b + 1 < b_end) {
            tma_wait_read<1>();  // the store that last read stage (b+1)%3 (block b-2) has finished reading SMEM
            issue_load(b + 1);
        }
        mbar_wait(&s_bar[st], (uint32_t) (((b - b_begin) / kStages) & 1));
        unsigned char* sb = stage_base + (size_t) st * P.stage_bytes;
        float* s_light = (float*) sb;
        const unsigned char* s_data = sb + P.data_off;
        const int s0 = block_s0(b);

#pragma unroll 1
        for (int sl = 0; sl < kSB; ++sl) {
            const int loop = s0 + (U.dirn > 0 ? sl : kSB - 1 - sl);
            if (loop >= ns) continue;
            const int k = U.dirn > 0 ? loop : ns - 1 - loop;  // position in sweep order
            const int fp_par = (k & 1) * (FPW * kFpH);
            float* fp_cur = s_fp + fp_par;
            float* fp_next = s_fp + (FPW * kFpH - fp_par);
            // ---- (a) issue the halo loads of slice k-1 (and, every 4th slice, the back-pressure probes) ----
            const unsigned int want_tag = tag_base + (unsigned) k;  // slice k-1 carries tag k
            const unsigned int rd_slot = (((unsigned int) k + kRingDepth - 1u) % kRingDepth) * plane32;  // first cell of slice k-1 in the ring
            const unsigned long long* rdi = SLAB ? P.S.inbox + (size_t) max(k - 1, 0) * inbox_plane : nullptr;  // inbox is full depth
            auto halo_load = [&](int i) {
                if (SLAB && halo_ring[i] < 0) return ld_relaxed_sys_u64(rdi + (-1 - halo_ring[i]));
                return ld_relaxed_u64(ring + (rd_slot + (unsigned int) halo_ring[i]));
            };
            unsigned long long hv[kHaloPerThread];
#pragma unroll
            for (int i = 0; i < kHaloPerThread; ++i) hv[i] = 0;
            if (k > k_begin) {
#pragma unroll
                for (int i = 0; i < kHaloPerThread; ++i)
                    if (halo_fp[i] >= 0) hv[i] = halo_load(i);
            }
            // a ring slot is reused every kRingDepth slices: before exporting slices k..k+3 every reader must have
            // consumed slice k+3-kRingDepth, i.e. passed the barrier of its slice k+4-kRingDepth
            const bool probe = (k & 3) == 0 && k + 4 > kRingDepth && tid < ndown;
            unsigned int pv = 0;
            if (probe) pv = ld_relaxed_u32(P.flags + (size_t) s_down[tid] * kFlagStride);

            // ---- (b) opacity toward the light for this thread's two voxels (independent of the previous slice) ----
            const int2 ms = __ldg(&P.A.ax[SA].meta[loop]);
            const float fs = __ldg(&P.A.ax[SA].f[loop]);
            const int rows = ms.x - (s0 * P.dk[2] + P.dmin[2]);
            const bool inS01 = (unsigned) ms.x < (unsigned) (dN_s - 1);  // taps ms.x and ms.x + 1 both inside (dN_s >= 1)
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
                    if (v)
                        w1 = w;
                    else
                        w0 = w;
                }
            }
            const bool gs = !gate || ms.y;
            const bool g0 = w0 > 0.0f && inside_pq0 && gs, g1 = w1 > 0.0f && inside_pq1 && gs;
            float cs0 = 0.0f, cs1 = 0.0f;
            if (g0 || g1) {
                // 4 rows (q, s) x 3 columns of taps; two aligned 32-bit loads + a funnel shift per row
                float t[3][2][2];
                uint32_t wq[2][2];
                const unsigned char* base = s_data + thread_data + rows * P.ds_s;
#pragma unroll
                for (int js = 0; js < 2; ++js)
#pragma unroll
                    for (int jq = 0; jq < 2; ++jq) {
                        const unsigned char* rowp = base + jq * P.ds_q + js * P.ds_s;
                        const uint32_t a = *(const uint32_t*) rowp, bb = *(const uint32_t*) (rowp + 4);
                        wq[jq][js] = __funnelshift_r(a, bb, shift);  // bytes 0..2 = the three taps; byte 3 is not a tap and is never looked at
                    }
                const bool all_in = all_pq && inS01;
                // Exact empty-space skip (SWAR "some byte > T", T = largest byte the low cut-off rejects): a trilinear value
                // lies between its smallest and largest tap and the window position is monotone in the value, so if no tap
                // exceeds T both samples are rejected and return exactly 0 (WindowedSampling.usf:28). Carries only travel upward, so the
                // stray byte 3 cannot disturb the three tap bytes, and the final mask drops its own flag.
                constexpr uint32_t kTapFlags = PX == 2 ? 0x00808080u : 0x00008080u;  // the flag bits of the tap bytes in use
                uint32_t any_gt = 1u;
                if (P.cut_lo_mode == 1)
                    any_gt = (((wq[0][0] + P.cut_lo_add) | wq[0][0]) | ((wq[0][1] + P.cut_lo_add) | wq[0][1]) |
                              ((wq[1][0] + P.cut_lo_add) | wq[1][0]) | ((wq[1][1] + P.cut_lo_add) | wq[1][1])) & kTapFlags;
                else if (P.cut_lo_mode == 2)
                    any_gt = ((((wq[0][0] & 0x7f7f7f7fu) + P.cut_lo_add) & wq[0][0]) | (((wq[0][1] & 0x7f7f7f7fu) + P.cut_lo_add) & wq[0][1]) |
                              (((wq[1][0] & 0x7f7f7f7fu) + P.cut_lo_add) & wq[1][0]) | (((wq[1][1] & 0x7f7f7f7fu) + P.cut_lo_add) & wq[1][1])) & kTapFlags;
                if (!(all_in && any_gt == 0u)) {
#pragma unroll
                    for (int js = 0; js < 2; ++js)
#pragma unroll
                        for (int jq = 0; jq < 2; ++jq) {
                            const uint32_t w = wq[jq][js];
                            t[0][jq][js] = decode_u8(w & 0xffu);
                            t[1][jq][js] = decode_u8((w >> 8) & 0xffu);
                            t[2][jq][js] = decode_u8((w >> 16) & 0xffu);
                        }
                    if (!all_in) {  // cold (volume faces only): the per-tap bounds are recomputed here rather than kept live through the loop
                        const bool ip[3] = {(unsigned) mp0.x < (unsigned) dN_p, (unsigned) (mp0.x + 1) < (unsigned) dN_p,
                                            PX == 1 || (unsigned) (mp0.x + 2) < (unsigned) dN_p};
                        const bool iq[2] = {(unsigned) mq.x < (unsigned) dN_q, (unsigned) (mq.x + 1) < (unsigned) dN_q};
                        const bool is[2] = {(unsigned) ms.x < (unsigned) dN_s, (unsigned) (ms.x + 1) < (unsigned) dN_s};
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
            }

            // ---- (c) the halo of slice k-1 must have arrived; readers of the slots we overwrite must have moved on ----
            if (k > k_begin) {
#pragma unroll
                for (int i = 0; i < kHaloPerThread; ++i)
                    if (halo_fp[i] >= 0) {
                        if (!SLAB) {
                            unsigned int polls = 0;
                            while ((unsigned int) (hv[i] >> 32) != want_tag && ++polls <= P.poll_limit) hv[i] = ld_relaxed_u64(ring + (rd_slot + (unsigned int) halo_ring[i]));
                            if (polls > P.poll_limit) give_up();
                        } else if ((unsigned int) (hv[i] >> 32) != want_tag) {
                            const unsigned long long t0 = global_timer_ns();
                            unsigned int polls = 0;
                            while ((unsigned int) (hv[i] >> 32) != want_tag && !s_abort) {
                                hv[i] = halo_load(i);
                                if ((++polls & 255u) == 0 && timed_out(t0)) break;
                            }
                        }
                        fp_cur[halo_fp[i]] = __uint_as_float((unsigned int) hv[i]);
                    }
                for (int h = tid; h < n_over; h += kTmaThreads) {
                    const int f = s_over_fp[h], g = s_over_ring[h];
                    if (f < 0) continue;
                    const unsigned long long* cell = (SLAB && g < 0) ? rdi + (-1 - g) : ring + (rd_slot + (unsigned int) g);
                    unsigned long long v = SLAB ? ld_relaxed_sys_u64(cell) : ld_relaxed_u64(cell);
                    const unsigned long long t0 = SLAB ? global_timer_ns() : 0ull;
                    unsigned int polls = 0;
                    while ((unsigned int) (v >> 32) != want_tag && !(SLAB && s_abort)) {
                        v = SLAB ? ld_relaxed_sys_u64(cell) : ld_relaxed_u64(cell);
                        ++polls;
                        if (SLAB ? ((polls & 255u) == 0 && timed_out(t0)) : polls > P.poll_limit) {
                            if (!SLAB) give_up();
                            break;
                        }
                    }
                    fp_cur[f] = __uint_as_float((unsigned int) v);
                }
            }
            if (probe) {
                const unsigned int need = (unsigned) (k + 4 - kRingDepth);
                unsigned int polls = 0;
                while (pv < need && !(SLAB && s_abort) && ++polls <= P.poll_limit) pv = ld_relaxed_u32(P.flags + (size_t) s_down[tid] * kFlagStride);
                if (polls > P.poll_limit) give_up();
            }
            __syncthreads();
            // progress counter for back-pressure: every read of slice k-1 by this tile is done
            if (tid == 0) st_relaxed_u32(P.flags + (size_t) tile * kFlagStride, (unsigned) k);

            // ---- (d) propagate, export, accumulate into the light brick ----
            {
                const float* r0 = fp_cur + tap_idx;
                const float* r1 = r0 + FPW;
                const float t00 = r0[0], t10 = r0[1], t20 = r0[2], t01 = r1[0], t11 = r1[1], t21 = r1[2];
                const float prev0 = lerpf(lerpf(t00, t10, bfx0), lerpf(t01, t11, bfx0), bfy);
                const float prev1 = lerpf(lerpf(t10, t20, bfx1), lerpf(t11, t21, bfx1), bfy);
                const float cur0 = prev0 * (1.0f - cs0), cur1 = prev1 * (1.0f - cs1);
                // what the next slice reads: the value as a read / write buffer of the light volume's format holds it
                const float fwd0 = L8 ? decode_u8((uint32_t) quant8(cur0)) : cur0, fwd1 = L8 ? decode_u8((uint32_t) quant8(cur1)) : cur1;
                if (own_in0) fp_next[own_idx] = fwd0;
                if (own_in1) fp_next[own_idx + 1] = fwd1;
                unsigned long long* wr = ring + (((unsigned int) k % kRingDepth) * plane32 + own_cell);
                const unsigned long long tag = (unsigned long long) (tag_base + (unsigned) k + 1u) << 32;
                if (exp0) st_relaxed_u64(wr, tag | __float_as_uint(fwd0));
                if (exp1) st_relaxed_u64(wr + 1, tag | __float_as_uint(fwd1));
                if (SLAB) {
                    const size_t ko = (size_t) k * inbox_plane;
                    if (xlo0) st_relaxed_sys_u64(P.S.out_lo + ko + xlo_idx, tag | __float_as_uint(fwd0));
                    if (xlo1) st_relaxed_sys_u64(P.S.out_lo + ko + xlo_idx + 1, tag | __float_as_uint(fwd1));
                    if (xhi0) st_relaxed_sys_u64(P.S.out_hi + ko + xhi_idx, tag | __float_as_uint(fwd0));
                    if (xhi1) st_relaxed_sys_u64(P.S.out_hi + ko + xhi_idx + 1, tag | __float_as_uint(fwd1));
                    if (P.S.zout != nullptr && k == k_end - 1) {  // hand the last slice of this slab to the next one
                        unsigned long long* zo = P.S.zout + (size_t) px + (size_t) tx * py;
                        if (v0) st_relaxed_sys_u64(zo, tag | __float_as_uint(fwd0));
                        if (v1) st_relaxed_sys_u64(zo + 1, tag | __float_as_uint(fwd1));
                    }
                }
                // pixels beyond a ragged plane edge (v0 / v1 false) update their cell of the SMEM brick too: the TMA store clips the brick
                // to the light volume, so those cells never reach memory — no validity test per slice
                float* lp = s_light + light_off + (loop - s0) * P.ls_s;
                if (L8 && P.mode == kModeAdd) {  // AddDirLightShader.usf:121-126 on a G8 light volume
                    unsigned char* lp8 = (unsigned char*) s_light + light_off + (loop - s0) * P.ls_s;
                    if (fabsf(cur0) > 1e-3f) lp8[0] = quant8(decode_u8((uint32_t) lp8[0]) + (cur0 * U.sign));
                    if (PX == 2 && fabsf(cur1) > 1e-3f) lp8[P.ls_p] = quant8(decode_u8((uint32_t) lp8[P.ls_p]) + (cur1 * U.sign));
                } else if (L8 && P.mode == kModeCombine) {  // ChangeDirLightShader.usf:146-153 on a G8 light volume: byte brick, R32F brick of the removed light behind it
                    unsigned char* lp8 = (unsigned char*) s_light + light_off + (loop - s0) * P.ls_s;
                    const float* rp = (const float*) ((const unsigned char*) s_light + P.light_bytes) + lx * P.ss_p + ly * P.ss_q + (loop - s0) * P.ss_s;
                    const float r0c = rp[0], r1c = PX == 2 ? rp[P.ss_p] : 0.0f;
                    if (fabsf(cur0 - r0c) > 1e-3f) lp8[0] = quant8(decode_u8((uint32_t) lp8[0]) + cur0 - r0c);
                    if (PX == 2 && fabsf(cur1 - r1c) > 1e-3f) lp8[P.ls_p] = quant8(decode_u8((uint32_t) lp8[P.ls_p]) + cur1 - r1c);
                } else if (P.mode == kModeAdd) {  // AddDirLightShader.usf:121-126
                    if (fabsf(cur0) > 1e-3f) lp[0] = lp[0] + (cur0 * U.sign);
                    if (PX == 2 && fabsf(cur1) > 1e-3f) lp[P.ls_p] = lp[P.ls_p] + (cur1 * U.sign);
                } else if (P.mode == kModeStore) {  // the removed light of a ChangeDirLight: its light goes to the scratch volume
                    lp[0] = cur0;
                    if (PX == 2) lp[P.ls_p] = cur1;
                } else {  // the added light of a ChangeDirLight: LightVolume += added - removed (ChangeDirLightShader.usf:146-153)
                    const float* rp = lp + P.light_bytes / 4;
                    const float r0c = rp[0], r1c = PX == 2 ? rp[P.ls_p] : 0.0f;
                    if (fabsf(cur0 - r0c) > 1e-3f) lp[0] = lp[0] + cur0 - r0c;
                    if (PX == 2 && fabsf(cur1 - r1c) > 1e-3f) lp[P.ls_p] = lp[P.ls_p] + cur1 - r1c;
                }
            }
        }
        // ---- write the updated light brick back ----
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            int lc[3];
            lc[PA] = x0, lc[QA] = y0, lc[SA] = s0;
            if (L8 && AXIS == 0 && P.mode != kModeStore) lc[0] = x0, lc[1] = y0, lc[2] = s0;
            tma_store_3d(&light_map, lc[0], lc[1], lc[2], s_light);
            // push-gather: the finished brick also goes into every other rank's light volume (NVLink), in the same bulk group — the transfer
            // overlaps the sweep tile by tile, and no all-gather of the light slabs follows the sweep
            for (int pr = 0; pr < P.n_push; ++pr) tma_store_3d(&push_maps.m[pr], lc[0], lc[1], lc[2], s_light);
            tma_commit();
        }
    }
    if (tid == 0) tma_wait_all<0>();
}

}  // namespace tbrm
