// tbrm_internal.hpp — state behind the opaque tbrm_resources handle and the launcher interfaces between the
// translation units of libtbrm.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/tbrm.h"
#include "host_plan.hpp"
#include "sweep_common.cuh"

// FBasicRaymarchRenderingResources (RaymarchTypes.h:87-129) + the stream that plays the render-thread queue
struct tbrm_resources {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool stream_owned = true;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;

    // DataVolumeTextureRef
    int32_t ddims[3] = {0, 0, 0};
    tbrm_format data_fmt = TBRM_FMT_G8;
    void* data = nullptr;
    bool data_owned = false;
    bool data_ready = false;

    // streaming upload / download (tbrm_upload_volume_async, tbrm_raymarch_lit_to_host_async)
    void* data_back = nullptr;               // back buffer of the data volume
    cudaStream_t upload_stream = nullptr, download_stream = nullptr;
    cudaEvent_t ev_uploaded = nullptr;       // the copy into data_back has finished
    cudaEvent_t ev_back_free = nullptr;      // the render queue no longer reads data_back (recorded when it was swapped out)
    bool upload_pending = false;
    void* frame_dev[2] = {nullptr, nullptr}; // device frames, used alternately
    size_t frame_bytes = 0;
    cudaEvent_t ev_frame_done[2] = {nullptr, nullptr};   // the raymarch into frame_dev[i] has finished
    cudaEvent_t ev_frame_copied[2] = {nullptr, nullptr}; // its copy to the host has finished
    bool frame_in_flight[2] = {false, false};
    int frame_next = 0;

    // TFTextureRef collapsed to 256 x RGBA fp32 (fp16-rounded values)
    float4* tf = nullptr;
    bool tf_ready = false;

    // LightVolumeRenderTarget
    int32_t ldims[3] = {0, 0, 0};
    tbrm_format light_fmt = TBRM_FMT_R32F;
    void* light = nullptr;
    bool half_res = false;

    tbrm_windowing windowing = {0.5f, 1.0f, 1, 1};  // VolumeInfo.h:36-46 defaults
    tbrm_options options = {};

    // XYZReadWriteBuffers[axis].Buffers[0..3] (RaymarchVolume.cpp:864-866,889-891), light pixel format
    void* rw[3][4] = {};
    // scratch of the fused sweep: ring of propagated slices + per-tile progress flags
    void* ring = nullptr;
    size_t ring_bytes = 0;
    unsigned int ring_epoch = 0;  // TMA sweep: tag epoch of the ring cells (0 = ring holds no valid tags)
    unsigned int* flags = nullptr;
    size_t flags_count = 0;
    unsigned int* sweep_err = nullptr;  // device word raised by a fused sweep launch whose wait on another tile timed out
    int32_t last_geom[4] = {0, 0, 0, 0};  // the last TMA-staged sweep launch: tile rows, pixels per thread, tiles of the plane, bands (tbrm_debug_download_derived 4)
    void* light_perm[2] = {nullptr, nullptr};  // G8 light volume, sweeps along X: its (y,z,x)- and (z,x,y)-ordered copies
    size_t light_perm_bytes = 0;
    void* tvol = nullptr;               // split sweep: T = 1 - occlusion bricks of the pass in flight (4 B per light voxel)
    size_t tvol_bytes = 0;
    void* dbg = nullptr;                // TBRM_CHAIN_TIMERS builds only
    void* tones = nullptr;              // ... and one byte per brick: every T of the brick is exactly 1
    size_t tones_bytes = 0;
    unsigned long long* counters = nullptr;  // device scratch for step / iteration counts
    // TMA sweep: (y,z,x)-ordered replica of the data volume for sweeps along X, and the per-pass sampler tables
    void* data_yzx = nullptr;
    bool data_yzx_valid = false;
    void* bricks = nullptr;  // raymarch: per-8^3-brick max of the data volume (exact empty-space skipping)
    bool bricks_valid = false;
    // OctreeVolumeRenderTarget (RaymarchTypes.h:104-106): 4 UNORM16 mips, sides rounded up to powers of two (RaymarchVolume.cpp:873-877)
    void* octree[4] = {nullptr, nullptr, nullptr, nullptr};
    bool octree_valid = false;
    void* tables = nullptr;
    size_t tables_bytes = 0;

    void* joined_buf = nullptr;  // joined sweeps: 2 propagation buffers per member pass (light pixel format)
    size_t joined_bytes = 0;

    bool light_owned = true;
    void* change_scratch = nullptr;  // TMA sweep: the removed light's propagated light of a ChangeDirLight (R32F, light volume dims)

    // Z-slab sharding of this volume over several GPUs and the exchange arena of partial sweep launches
    tbrm_slab slab = {0, 1, 0, 0};
    unsigned int pass_seq = 0;     // TMA passes executed: the tag sequence of ring / arena cells
    void* arena = nullptr;         // header + 4 regions x (inbox + hand-off plane) of LL cells, written by the neighbours
    size_t arena_bytes = 0;
    void* peer_arena[2] = {nullptr, nullptr};  // arenas of the slabs below / above (peer-mapped or same-process pointers)
    bool peer_ipc[2] = {false, false};
    int slab_timeout_ms = 0;       // 0 = default (4 s)
    // push-gather: the light volumes of ALL ranks of a sharded volume (peer-mapped); the last axis pass of a sweep call may store every
    // finished light brick into them as well, which replaces the trailing all-gather of the light slabs (tbrm_slab_push_light)
    static constexpr int kMaxPushPeers = 15;
    void* peer_light[kMaxPushPeers + 1] = {};
    bool peer_light_ipc[kMaxPushPeers + 1] = {};
    bool push_light = false;       // the caller asked for the next AddDirLight / ChangeDirLight call to push its last pass
    bool push_this_pass = false;   // set by the pass driver for that last pass

    size_t light_voxels() const { return (size_t) ldims[0] * ldims[1] * ldims[2]; }
    size_t data_voxels() const { return (size_t) ddims[0] * ddims[1] * ddims[2]; }
    size_t light_elem() const { return light_fmt == TBRM_FMT_G8 ? 1 : 4; }
    size_t data_elem() const { return data_fmt == TBRM_FMT_G8 ? 1 : (data_fmt == TBRM_FMT_G16 ? 2 : 4); }
};

namespace tbrm {

extern std::atomic<long long> g_kernel_launches;
void set_last_error(const std::string& msg);
inline void count_launch(int n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

// sweep.cu
cudaError_t sweep_fill_buffer(tbrm_resources& r, void* buf, size_t count, float value);
cudaError_t sweep_pass_per_slice(tbrm_resources& r, const SweepUniforms& u, bool change, int* launches);
constexpr int kMaxJoined = 8;  // member passes of one joined sweep
cudaError_t sweep_pass_joined(tbrm_resources& r, const SweepUniforms& u, const LightPass* members, int n_members, int* launches);
cudaError_t sweep_pass_fused(tbrm_resources& r, const SweepUniforms& u, bool change, int* launches, bool* handled);
cudaError_t sweep_pass_tma(tbrm_resources& r, const SweepUniforms& u, bool change, int* launches, bool* handled);
cudaError_t clear_light(tbrm_resources& r, float value);
cudaError_t build_replica_for_tests(tbrm_resources& r);  // the (y,z,x)-ordered replica of an R8 data volume (tbrm_debug_download_derived)
int slab_pass_order(const SweepUniforms& u);  // +1 lower slabs first, -1 higher slabs first, 2 only concurrently
size_t slab_arena_bytes(const int32_t ldims[3]);
cudaError_t slab_ensure_arena(tbrm_resources& r);
// the partition rule every rank applies: slabs are multiples of 8 slices, the last rank takes the remainder
inline void slab_partition(int Z, int nranks, int rank, int32_t* z_begin, int32_t* z_end) {
    const int per = std::max(8, (Z / nranks + 7) / 8 * 8);
    *z_begin = std::min(Z, rank * per);
    *z_end = rank + 1 == nranks ? Z : std::min(Z, (rank + 1) * per);
}

// raymarch.cu
cudaError_t raymarch_cube_setup(tbrm_resources& r, const host::CameraUniforms& cam, float* d_out);
cudaError_t raymarch_lit(tbrm_resources& r, const host::CameraUniforms& cam, const float clip_center[3], const float clip_dir[3],
                         float step_count, int row_begin, int row_end, int row_block, int block_stride, float* d_out,
                         unsigned long long* d_steps);
cudaError_t ensure_bricks(tbrm_resources& r);  // brick max-grid: 1 byte per 8^3 brick (max over [8b, 8b+8] per axis)
int raymarch_local_rows(int row_begin, int row_end, int row_block, int block_stride);
// materials.cuh (part of raymarch.cu): octree generation, intensity and octree marches
void octree_mip_dims(const tbrm_resources& r, int mip, int32_t dims[3]);
cudaError_t generate_octree(tbrm_resources& r);
cudaError_t raymarch_intensity(tbrm_resources& r, const host::CameraUniforms& cam, const float clip_center[3], const float clip_dir[3],
                               float step_count, int row_begin, int row_end, float* d_out, unsigned long long* d_steps);
cudaError_t raymarch_octree(tbrm_resources& r, const host::CameraUniforms& cam, const float clip_center[3], const float clip_dir[3],
                            float step_count, int octree_mip, int row_begin, int row_end, float* d_out, unsigned long long* d_steps);

// mandelbulb.cu
cudaError_t mandelbulb_march(cudaStream_t stream, const tbrm_mandelbulb& mb, const host::CameraUniforms& cam, int row_begin,
                             int row_end, float* d_out, unsigned long long* d_iters);

cudaError_t mandelbulb_march_normal(cudaStream_t stream, const tbrm_mandelbulb& mb, float derivation_distance, const host::CameraUniforms& cam,
                                    int row_begin, int row_end, float* d_out, unsigned long long* d_iters);
float mandelbulb_sdf_p8_host(float px, float py, float pz, float bailout, int iterations, unsigned int* iters);
cudaError_t mandelbulb_sdf_bake(cudaStream_t stream, const int32_t dims[3], const float center[3], float extent, float power, int g16, void* d_out,
                                unsigned long long* d_iters);

// ingest.cu
int voxel_format_bytes(int fmt);
size_t ingest_partials_bytes();
cudaError_t ingest_normalize(cudaStream_t stream, int fmt, const void* d_in, size_t n, void* d_out, void* d_partials, float* d_minmax);
cudaError_t ingest_to_float(cudaStream_t stream, int fmt, const void* d_in, size_t n, float* d_out);
bool mhd_parse_header(const std::string& text, tbrm_volume_info& out);
bool load_voxel_file(const std::string& path, const tbrm_volume_info& info, std::vector<uint8_t>& voxels, std::string& err);

// synth.cu
cudaError_t synth_volume_u8(cudaStream_t stream, int kind, const int32_t dims[3], uint32_t seed, uint8_t* d_out);

}  // namespace tbrm
