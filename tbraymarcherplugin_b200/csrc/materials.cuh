// materials.cuh — the reference's other two raymarch materials and the octree they read, sm_100a.
// Included at the end of raymarch.cu (shares its camera / cube-setup / sampler device functions).
//   GenerateOctreeShader.usf:28-107 + OctreeShaders.cpp:28-54      -> octree_build_kernel (one launch for all 4 mips)
//   PerformWindowedIntensityRaymarch, WindowedRaymarchMaterials.usf:187-242 -> raymarch_intensity_kernel
//   PerformWindowedRaymarchOctree,    WindowedRaymarchMaterials.usf:99-183  -> raymarch_octree_kernel
// Same fp32 arithmetic contract as the lit march: results are bit-identical to the oracle (tests/test_gpu_zz_materials.py), which is
// itself bit-identical to the reference's shader code compiled for the CPU (tests/test_ref_materials_cpu.py).
#pragma once

namespace tbrm {

// ---- octree: 4-level max pyramid of the data volume in a pow-2 sized UNORM16 volume ---------------------------------------------------
// The reference runs ONE thread per 8^3 leaf ([numthreads(1,1,1)]) which copies 512 voxels and reduces them three times through the UAV.
// Here one 256-thread CTA owns a 128 x 8 x 8 strip (16 leaves along x): a warp reads one 128-voxel row segment with 4-voxel loads and
// writes its 256 bytes of mip 0 with 8-byte stores; mip 0 stays in 16 KiB of shared memory, mips 1-3 are reduced from shared memory — the data volume is read once, every octree texel is written once:
// algorithmic bytes = B_d per data voxel + 2 * (1 + 1/8 + 1/64 + 1/512) per octree voxel. UNORM16 values round-trip exactly through
// the float load / store of the shader (v/65535 -> floor(v'*65535 + .5) is the identity on 0..65535), so the reductions are integer maxima.
struct OctreeUniforms {
    int ddims[3];     // data volume
    int odims[4][3];  // octree mip dimensions
};

__device__ __forceinline__ unsigned int quant16(float v) { return (unsigned int) floorf(saturatef(v) * 65535.0f + 0.5f); }

// The shader's store of a loaded texel into the UNORM16 UAV, floor(saturate(Load(p).r) * 65535 + .5): for UNORM8 data that is the byte times
// 257, for UNORM16 data the identity (both checked exhaustively in tests/test_host_cpu.py) — integer work; float data takes the float path.
__device__ __forceinline__ unsigned int octree_texel(uint8_t v) { return (unsigned int) v * 257u; }
__device__ __forceinline__ unsigned int octree_texel(uint16_t v) { return v; }
__device__ __forceinline__ unsigned int octree_texel(float v) { return quant16(v * 1.0f); }

template <typename T, int N>
struct alignas(sizeof(T) * N) OctPack {
    T v[N];
};

// mips 1-3 of one 128 x 8 x 8 strip whose mip 0 sits in shared memory (s0); called by all 256 threads of the CTA
__device__ __forceinline__ void octree_reduce_strip(const OctreeUniforms& U, const uint16_t (*s0)[8][128], uint16_t (*s1)[4][64], uint16_t (*s2)[2][32],
                                                    uint16_t* __restrict__ mip1, uint16_t* __restrict__ mip2, uint16_t* __restrict__ mip3, int x0,
                                                    int y0, int z0, int t) {
    __syncthreads();
    // strip texels beyond the octree's bounds hold 0 in shared memory (they lie outside the data volume: octree sides >= data sides),
    // which is what the shader's Load returns for them
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // mip 1: 64 x 4 x 4 texels of this strip, four per thread
        const int e = t + 256 * k;
        const int mx = e & 63, my = (e >> 6) & 3, mz = e >> 8;
        unsigned int m = 0;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int a = 0; a < 2; ++a) m = max(m, (unsigned int) s0[2 * mz + c][2 * my + b][2 * mx + a]);
        s1[mz][my][mx] = (uint16_t) m;
        const int gx = x0 / 2 + mx, gy = y0 / 2 + my, gz = z0 / 2 + mz;
        if (gx < U.odims[1][0] && gy < U.odims[1][1] && gz < U.odims[1][2])
            mip1[(size_t) gx + (size_t) U.odims[1][0] * ((size_t) gy + (size_t) U.odims[1][1] * gz)] = (uint16_t) m;
    }
    __syncthreads();
    if (t < 128) {  // mip 2: 32 x 2 x 2
        const int mx = t & 31, my = (t >> 5) & 1, mz = t >> 6;
        unsigned int m = 0;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int a = 0; a < 2; ++a) m = max(m, (unsigned int) s1[2 * mz + c][2 * my + b][2 * mx + a]);
        s2[mz][my][mx] = (uint16_t) m;
        const int gx = x0 / 4 + mx, gy = y0 / 4 + my, gz = z0 / 4 + mz;
        if (gx < U.odims[2][0] && gy < U.odims[2][1] && gz < U.odims[2][2])
            mip2[(size_t) gx + (size_t) U.odims[2][0] * ((size_t) gy + (size_t) U.odims[2][1] * gz)] = (uint16_t) m;
    }
    __syncthreads();
    if (t < 16) {  // mip 3: 16 x 1 x 1
        unsigned int m = 0;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int a = 0; a < 2; ++a) m = max(m, (unsigned int) s2[c][b][2 * t + a]);
        const int gx = x0 / 8 + t, gy = y0 / 8, gz = z0 / 8;
        if (gx < U.odims[3][0] && gy < U.odims[3][1] && gz < U.odims[3][2])
            mip3[(size_t) gx + (size_t) U.odims[3][0] * ((size_t) gy + (size_t) U.odims[3][1] * gz)] = (uint16_t) m;
    }
}

// R8 data with X % 16 == 0 (and a 16-byte aligned base): a thread owns 16 voxels along x — one 16-byte load, two 16-byte stores of mip 0.
// UNORM8 -> UNORM16 through the shader's float round trip, floor(saturate(b / 255) * 65535 + .5), equals b * 257 for every byte
// (tests/test_host_cpu.py checks all 256), i.e. the byte replicated into both halves: one PRMT per two voxels.
__global__ void __launch_bounds__(256) octree_build_u8x16_kernel(const OctreeUniforms U, const uint8_t* __restrict__ data, uint16_t* __restrict__ mip0,
                                                                 uint16_t* __restrict__ mip1, uint16_t* __restrict__ mip2, uint16_t* __restrict__ mip3) {
    __shared__ __align__(16) uint16_t s0[8][8][128];
    __shared__ uint16_t s1[4][4][64];
    __shared__ uint16_t s2[2][2][32];
    const int x0 = blockIdx.x * 128, y0 = blockIdx.y * 8, z0 = blockIdx.z * 8;
    const int t = threadIdx.x, seg = t & 7, ly = (t >> 3) & 7, lzq = t >> 6;  // 8 segments x 8 rows x 4 slices per round
    const int X = U.ddims[0], Y = U.ddims[1], Z = U.ddims[2];
    const int OX = U.odims[0][0], OY = U.odims[0][1], OZ = U.odims[0][2];
    const int x = x0 + 16 * seg, y = y0 + ly;
    uint4 w[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {  // both loads in flight before the first store
        const int z = z0 + 4 * it + lzq;
        w[it] = make_uint4(0u, 0u, 0u, 0u);  // Load outside the data volume returns 0 (GenerateOctreeShader.usf:36-49)
        if (x < X && y < Y && z < Z) w[it] = __ldg(reinterpret_cast<const uint4*>(data + (size_t) x + (size_t) X * ((size_t) y + (size_t) Y * z)));
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int lz = 4 * it + lzq, z = z0 + lz;
        const unsigned int in[4] = {w[it].x, w[it].y, w[it].z, w[it].w};
        uint4 o[2];
        o[0] = make_uint4(__byte_perm(in[0], 0u, 0x1100u), __byte_perm(in[0], 0u, 0x3322u), __byte_perm(in[1], 0u, 0x1100u), __byte_perm(in[1], 0u, 0x3322u));
        o[1] = make_uint4(__byte_perm(in[2], 0u, 0x1100u), __byte_perm(in[2], 0u, 0x3322u), __byte_perm(in[3], 0u, 0x1100u), __byte_perm(in[3], 0u, 0x3322u));
        uint4* sdst = reinterpret_cast<uint4*>(&s0[lz][ly][16 * seg]);
        sdst[0] = o[0], sdst[1] = o[1];
        if (x < OX && y < OY && z < OZ) {  // OX is a power of two >= X >= 16: whole segments only
            uint4* dst = reinterpret_cast<uint4*>(mip0 + (size_t) x + (size_t) OX * ((size_t) y + (size_t) OY * z));
            dst[0] = o[0], dst[1] = o[1];
        }
    }
    octree_reduce_strip(U, s0, s1, s2, mip1, mip2, mip3, x0, y0, z0, t);
}

// VEC4: the data rows can be read 4 voxels at a time (X % 4 == 0 and an aligned base pointer)
template <typename DataT, bool VEC4>
__global__ void __launch_bounds__(256) octree_build_kernel(const OctreeUniforms U, const DataT* __restrict__ data, uint16_t* __restrict__ mip0,
                                                           uint16_t* __restrict__ mip1, uint16_t* __restrict__ mip2, uint16_t* __restrict__ mip3) {
    __shared__ __align__(8) uint16_t s0[8][8][128];
    __shared__ uint16_t s1[4][4][64];
    __shared__ uint16_t s2[2][2][32];
    const int x0 = blockIdx.x * 128, y0 = blockIdx.y * 8, z0 = blockIdx.z * 8;
    const int t = threadIdx.x, tx = t & 31, ly = t >> 5;
    const int X = U.ddims[0], Y = U.ddims[1], Z = U.ddims[2];
    const int OX = U.odims[0][0], OY = U.odims[0][1], OZ = U.odims[0][2];
    // mip 0: OctreeVolumeMip0[p] = Volume.Load(p).r * MinMaxValues.y (= 1); Load outside the data volume returns 0 (:36-49).
    // A thread owns 4 voxels along x: a warp reads one 128-voxel row segment and writes 256 bytes of it.
    const int x = x0 + 4 * tx, y = y0 + ly;
#pragma unroll
    for (int lz = 0; lz < 8; ++lz) {
        const int z = z0 + lz;
        unsigned int q[4] = {0, 0, 0, 0};
        if (y < Y && z < Z) {
            const size_t row = (size_t) X * ((size_t) y + (size_t) Y * z);
            if (VEC4 && x + 3 < X) {
                const OctPack<DataT, 4> p = *reinterpret_cast<const OctPack<DataT, 4>*>(data + row + x);
#pragma unroll
                for (int k = 0; k < 4; ++k) q[k] = octree_texel(p.v[k]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (x + k < X) q[k] = octree_texel(__ldg(data + row + x + k));
            }
        }
        OctPack<uint16_t, 4> o;
#pragma unroll
        for (int k = 0; k < 4; ++k) o.v[k] = (uint16_t) q[k];
        *reinterpret_cast<OctPack<uint16_t, 4>*>(&s0[lz][ly][4 * tx]) = o;
        if (y < OY && z < OZ) {  // out-of-bounds UAV stores are dropped
            uint16_t* dst = mip0 + (size_t) x + (size_t) OX * ((size_t) y + (size_t) OY * z);
            if (x + 3 < OX && (OX & 3) == 0) {
                *reinterpret_cast<OctPack<uint16_t, 4>*>(dst) = o;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (x + k < OX) dst[k] = o.v[k];
            }
        }
    }
    octree_reduce_strip(U, s0, s1, s2, mip1, mip2, mip3, x0, y0, z0, t);
}

// ---- per-pixel preamble shared by the two marches (WindowedRaymarchMaterials.usf:113-130, 196-208) -------------------------------------
struct MaterialUniforms {
    MarchUniforms M;
    const uint16_t* mip;  // octree march: the mip that is sampled
    int mdims[3];         // its dimensions
    float octree_depth0;  // OctreeDepthConst: depth of octree mip 0
    float rwidth;         // RN(1 / window width), for the division-free window position (div_exact)
};

struct PixelMarch {
    V3 cur, sv;
    int max_steps;
    float fin, ss;
};
__device__ __forceinline__ PixelMarch pixel_march(const MarchUniforms& U, int ix, int iy) {
    const V3 V = camera_vector(U.cam, ix, iy);
    V3 cur, lcv;
    float thick;
    cube_setup(U.cam, V, cur, thick, lcv);
    PixelMarch p;
    p.ss = 1 / U.step_count;
    const float fas = U.step_count * thick;
    const float fl = floorf(fas);
    p.max_steps = (int) fl;
    p.fin = fas - fl;
    p.sv = v3(lcv.x * p.ss, lcv.y * p.ss, lcv.z * p.ss);
    if (U.cam.jitter) {
        const float rnd = (float) pcg16_x(ix, iy, U.cam.frame_mod8) / 65535.0f;
        cur = v3(cur.x - p.sv.x * rnd, cur.y - p.sv.y * rnd, cur.z - p.sv.z * rnd);
    }
    p.cur = cur;
    return p;
}
__device__ __forceinline__ bool is_clipped(const MarchUniforms& U, V3 p) {  // IsCurPosClipped — RaymarcherCommon.usf:22-25
    return dot3(p.x - U.clip_center[0], p.y - U.clip_center[1], p.z - U.clip_center[2], U.clip_dir[0], U.clip_dir[1], U.clip_dir[2]) <= 0.0f;
}
__device__ __forceinline__ void count_steps(unsigned long long* steps_out, unsigned int steps) {
    if (!steps_out) return;
    unsigned int s = steps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(steps_out, (unsigned long long) s);
}

// PerformWindowedIntensityRaymarch: the first unclipped sample, windowed to grey, alpha 1
template <typename DataT>
__global__ void __launch_bounds__(256) raymarch_intensity_kernel(const MarchUniforms U, const DataT* __restrict__ data, float4* __restrict__ out,
                                                                 unsigned long long* __restrict__ steps_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ix = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int iy = U.row_begin + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    unsigned int steps = 0;
    if (ix < U.cam.width && iy < U.row_end) {
        const PixelMarch m = pixel_march(U, ix, iy);
        V3 cur = m.cur;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);  // :241 "didn't hit anything"
        bool hit = false;
        for (int i = 0; i < m.max_steps; i++) {  // :211
            cur = v3(cur.x + m.sv.x, cur.y + m.sv.y, cur.z + m.sv.z);
            ++steps;
            const V3 sp = v3(saturatef(cur.x), saturatef(cur.y), saturatef(cur.z));
            if (!is_clipped(U, sp)) {  // :215 tests the SATURATED position
                const float v = sample_data<DataT>(data, U.ddims, sp, false);  // :217 Material.Clamp_WorldGroupSettings
                float pos;
                tf_position(v, U.win, pos);
                pos = saturatef(pos);  // :220 clamp(..., 0, 1)
                o = make_float4(pos, pos, pos, 1.0f);
                hit = true;
                break;  // :222 return
            }
        }
        if (!hit && m.fin > 0.0f) {  // :227
            cur = v3(cur.x + m.sv.x * m.fin, cur.y + m.sv.y * m.fin, cur.z + m.sv.z * m.fin);
            ++steps;
            if (!is_clipped(U, cur)) {  // :231 tests the raw position
                const float v = sample_data<DataT>(data, U.ddims, cur, false);
                float pos;
                tf_position(v, U.win, pos);
                pos = saturatef(pos);
                o = make_float4(pos, pos, pos, 1.0f);
            }
        }
        out[(size_t) (iy - U.row_begin) * U.cam.width + ix] = o;
    }
    count_steps(steps_out, steps);
}

// one sample of PerformWindowedRaymarchOctree: point Load from the octree mip -> windowed TF -> AccumulateLightEnergy (no light volume)
// v / 65535 for v in 0..65535 without a division: with 1/65535 = c_hi + c_lo (c_hi = RN(1/65535)), fma(v, c_hi, v * c_lo) == RN(v / 65535)
// for every 16-bit value (exhaustive check in tests/test_host_cpu.py), like decode_u8_exact
__device__ __forceinline__ float decode_u16_exact(uint32_t v) {
    const float x = (float) v;
    return __fmaf_rn(x, 1.5259021893143654e-05f, x * 3.5527678889091252e-15f);
}

__device__ __forceinline__ void octree_sample(const MaterialUniforms& F, const float4* s_tf, V3 p, float step, float ow, float oh, float od, float dd,
                                              float4& acc) {
    // int3 VoxelPos = float3(CurPos.x * OctreeWidth, CurPos.y * OctreeHeight, (CurPos.z * DataVolumeDepth / OctreeDepthConst) * OctreeDepth) — :148
    const int vx = (int) (p.x * ow), vy = (int) (p.y * oh), vz = (int) (((p.z * dd) / F.octree_depth0) * od);
    float v = 0.0f;  // Texture3D.Load out of bounds returns 0
    if ((unsigned) vx < (unsigned) F.mdims[0] && (unsigned) vy < (unsigned) F.mdims[1] && (unsigned) vz < (unsigned) F.mdims[2])
        v = decode_u16_exact(__ldg(F.mip + (size_t) vx + (size_t) F.mdims[0] * ((size_t) vy + (size_t) F.mdims[1] * vz)));
    // GetTransferFuncPosition + cut-offs (tf_position) with the lit march's correctly rounded division-free quotient
    const Windowing& w = F.M.win;
    const float pos = div_exact(v - w.center + (w.width / 2.0f), w.width, F.rwidth);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!((pos < 0.0f && w.low > 0.0f) || (pos > 1.0f && w.high > 0.0f))) {
        int i0, i1;
        float f;
        tf_taps(pos, i0, i1, f);
        const float4 a = s_tf[i0], b = s_tf[i1];
        s = make_float4(lerpf(a.x, b.x, f), lerpf(a.y, b.y, f), lerpf(a.z, b.z, f), lerpf(a.w, b.w, f));
        s.w = step_opacity(s.w, step);
    }
    const float oma = 1.0f - acc.w;
    acc.x = acc.x + ((s.x * s.w) * oma);
    acc.y = acc.y + ((s.y * s.w) * oma);
    acc.z = acc.z + ((s.z * s.w) * oma);
    acc.w = acc.w + (s.w * oma);
}

__global__ void __launch_bounds__(256) raymarch_octree_kernel(const MaterialUniforms F, const float4* __restrict__ tf, float4* __restrict__ out,
                                                              unsigned long long* __restrict__ steps_out) {
    const MarchUniforms& U = F.M;
    __shared__ float4 s_tf[256];
    s_tf[threadIdx.x] = __ldg(&tf[threadIdx.x]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ix = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int iy = U.row_begin + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    unsigned int steps = 0;
    if (ix < U.cam.width && iy < U.row_end) {
        const PixelMarch m = pixel_march(U, ix, iy);
        V3 cur = m.cur;
        const float ssw = 100.0f * m.ss;  // :119
        // GetDimensions of the requested mip (:135)
        const float ow = (float) F.mdims[0], oh = (float) F.mdims[1], od = (float) F.mdims[2], dd = (float) U.ddims[2];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int i = 0;
        for (i = 0; i < m.max_steps; i++) {  // :138
            cur = v3(cur.x + m.sv.x, cur.y + m.sv.y, cur.z + m.sv.z);
            ++steps;
            if (!is_clipped(U, cur)) {
                octree_sample(F, s_tf, cur, ssw, ow, oh, od, dd, acc);
                if (acc.w > 0.95f) {  // :156-160
                    acc.w = 1.0f;
                    break;
                }
            }
        }
        if (i == m.max_steps && m.fin > 0.0f) {  // :165 — the opacity step stays StepSizeWorld (:173), unlike the lit march
            cur = v3(cur.x + m.sv.x * m.fin, cur.y + m.sv.y * m.fin, cur.z + m.sv.z * m.fin);
            ++steps;
            if (!is_clipped(U, cur)) octree_sample(F, s_tf, cur, ssw, ow, oh, od, dd, acc);
        }
        out[(size_t) (iy - U.row_begin) * U.cam.width + ix] = acc;
    }
    count_steps(steps_out, steps);
}

// ---- host side -------------------------------------------------------------------------------------------------------------------
static inline int round_up_pow2(int v) {  // FMath::RoundUpToPowerOfTwo
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
void octree_mip_dims(const tbrm_resources& r, int mip, int32_t dims[3]) {  // RaymarchVolume.cpp:873-877: pow-2 sides, 4 mips
    for (int k = 0; k < 3; ++k) dims[k] = std::max(1, round_up_pow2(r.ddims[k]) >> mip);
}

template <typename DataT>
static cudaError_t launch_octree(tbrm_resources& r, const OctreeUniforms& U) {
    const dim3 grid((U.odims[0][0] + 127) / 128, (U.odims[0][1] + 7) / 8, (U.odims[0][2] + 7) / 8);
    const bool vec4 = (U.ddims[0] & 3) == 0 && (reinterpret_cast<uintptr_t>(r.data) % (4 * sizeof(DataT))) == 0;
    if (std::is_same<DataT, uint8_t>::value && (U.ddims[0] & 15) == 0 && (reinterpret_cast<uintptr_t>(r.data) & 15) == 0)
        octree_build_u8x16_kernel<<<grid, 256, 0, r.stream>>>(U, (const uint8_t*) r.data, (uint16_t*) r.octree[0], (uint16_t*) r.octree[1],
                                                             (uint16_t*) r.octree[2], (uint16_t*) r.octree[3]);
    else if (vec4)
        octree_build_kernel<DataT, true><<<grid, 256, 0, r.stream>>>(U, (const DataT*) r.data, (uint16_t*) r.octree[0], (uint16_t*) r.octree[1],
                                                                    (uint16_t*) r.octree[2], (uint16_t*) r.octree[3]);
    else
        octree_build_kernel<DataT, false><<<grid, 256, 0, r.stream>>>(U, (const DataT*) r.data, (uint16_t*) r.octree[0], (uint16_t*) r.octree[1],
                                                                     (uint16_t*) r.octree[2], (uint16_t*) r.octree[3]);
    count_launch();
    return cudaGetLastError();
}

// GenerateOctreeForVolume_RenderThread — OctreeShaders.cpp:28-54
cudaError_t generate_octree(tbrm_resources& r) {
    OctreeUniforms U;
    cudaError_t e;
    for (int m = 0; m < 4; ++m) {
        int32_t d[3];
        octree_mip_dims(r, m, d);
        for (int k = 0; k < 3; ++k) U.odims[m][k] = d[k];
        if (!r.octree[m] && (e = cudaMalloc(&r.octree[m], (size_t) d[0] * d[1] * d[2] * sizeof(uint16_t))) != cudaSuccess) return e;
    }
    for (int k = 0; k < 3; ++k) U.ddims[k] = r.ddims[k];
    switch (r.data_fmt) {
        case TBRM_FMT_G8: e = launch_octree<uint8_t>(r, U); break;
        case TBRM_FMT_G16: e = launch_octree<uint16_t>(r, U); break;
        default: e = launch_octree<float>(r, U); break;
    }
    if (e == cudaSuccess) r.octree_valid = true;
    return e;
}

static void fill_march_uniforms(const tbrm_resources& r, const host::CameraUniforms& cam, const float clip_center[3], const float clip_dir[3],
                                float step_count, int row_begin, int row_end, MarchUniforms& U) {
    U.cam = to_raycam(cam);
    for (int k = 0; k < 3; ++k) {
        U.ddims[k] = r.ddims[k];
        U.ldims[k] = r.ldims[k];
        U.clip_center[k] = clip_center[k];
        U.clip_dir[k] = clip_dir[k];
    }
    U.win = Windowing{r.windowing.center, r.windowing.width, r.windowing.low_cutoff ? 1.0f : 0.0f, r.windowing.high_cutoff ? 1.0f : 0.0f};
    U.step_count = step_count;
    U.row_begin = row_begin, U.row_end = row_end;
    U.row_block = 8, U.block_stride = 1;
    U.data_wrap = 0;
}

cudaError_t raymarch_intensity(tbrm_resources& r, const host::CameraUniforms& cam, const float clip_center[3], const float clip_dir[3],
                               float step_count, int row_begin, int row_end, float* d_out, unsigned long long* d_steps) {
    MarchUniforms U;
    fill_march_uniforms(r, cam, clip_center, clip_dir, step_count, row_begin, row_end, U);
    const dim3 grid((cam.width + 31) / 32, (row_end - row_begin + 7) / 8);
    switch (r.data_fmt) {
        case TBRM_FMT_G8: raymarch_intensity_kernel<uint8_t><<<grid, 256, 0, r.stream>>>(U, (const uint8_t*) r.data, (float4*) d_out, d_steps); break;
        case TBRM_FMT_G16: raymarch_intensity_kernel<uint16_t><<<grid, 256, 0, r.stream>>>(U, (const uint16_t*) r.data, (float4*) d_out, d_steps); break;
        default: raymarch_intensity_kernel<float><<<grid, 256, 0, r.stream>>>(U, (const float*) r.data, (float4*) d_out, d_steps); break;
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t raymarch_octree(tbrm_resources& r, const host::CameraUniforms& cam, const float clip_center[3], const float clip_dir[3],
                            float step_count, int octree_mip, int row_begin, int row_end, float* d_out, unsigned long long* d_steps) {
    MaterialUniforms F;
    fill_march_uniforms(r, cam, clip_center, clip_dir, step_count, row_begin, row_end, F.M);
    int32_t d0[3];
    octree_mip_dims(r, 0, d0);
    F.octree_depth0 = (float) d0[2];
    F.rwidth = 1.0f / F.M.win.width;
    int32_t dm[3];
    octree_mip_dims(r, octree_mip, dm);  // 0 <= octree_mip < 4: checked by the C entry point
    for (int k = 0; k < 3; ++k) F.mdims[k] = dm[k];
    F.mip = (const uint16_t*) r.octree[octree_mip];
    const dim3 grid((cam.width + 31) / 32, (row_end - row_begin + 7) / 8);
    raymarch_octree_kernel<<<grid, 256, 0, r.stream>>>(F, r.tf, (float4*) d_out, d_steps);
    count_launch();
    return cudaGetLastError();
}

}  // namespace tbrm
