// api.cu — the C ABI of include/tbrm.h: resource lifetime, validation, the host drivers of the sweep
// (AddDirLightToSingleLightVolume_RenderThread / ChangeDirLightInSingleLightVolume_RenderThread,
// Source/Raymarcher/Private/Rendering/LightingShaders.cpp:35-326) and the raymarch / Mandelbulb entry points.
// Every op is enqueued on the resource set's stream, which plays the role of UE's render-thread queue.
#include <cuda_fp16.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "tbrm_internal.hpp"

// NVTX ranges around the C-ABI operations — what SCOPED_GPU_STAT / SCOPED_DRAW_EVENT are to the reference's render commands
// (LightingShaders.cpp:25-30, 57-58): Nsight Systems / Compute show "tbrm::AddDirLight" etc. Header-only NVTX v3: no link dependency, a
// no-op unless a tool is attached.
#ifndef TBRM_HOST_EMULATION
#include <nvtx3/nvToolsExt.h>
struct TbrmRange {
    explicit TbrmRange(const char* name) { nvtxRangePushA(name); }
    ~TbrmRange() { nvtxRangePop(); }
};
#else
struct TbrmRange {
    explicit TbrmRange(const char*) {}
};
#endif
#define TBRM_RANGE(name) TbrmRange tbrm_range_(name)

namespace tbrm {
std::atomic<long long> g_kernel_launches{0};
static thread_local std::string t_last_error;
void set_last_error(const std::string& msg) { t_last_error = msg; }
}  // namespace tbrm

using namespace tbrm;

#define TBRM_CUDA(expr)                                                                            \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
            return TBRM_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

#define TBRM_REQUIRE(cond, msg)                 \
    do {                                        \
        if (!(cond)) {                          \
            set_last_error(msg);                \
            return TBRM_ERR_INVALID_ARGUMENT;   \
        }                                       \
    } while (0)

// run `enqueue(d_out)` with a device output buffer and deliver it to a host or device destination
template <typename F>
static tbrm_status with_output(int device, cudaStream_t stream, void* dst, size_t bytes, int dst_is_device, F enqueue) {
    if (dst_is_device) {
        TBRM_CUDA(enqueue(dst));
        return TBRM_OK;
    }
    // grow-only staging buffer per host thread and device: cudaFree synchronises the whole device and would stall the upload / download
    // streams of the streaming pipeline once per frame
    struct Staging {
        void* p = nullptr;
        size_t cap = 0;
        int dev = -1;
        ~Staging() {
            if (p && dev >= 0 && cudaSetDevice(dev) == cudaSuccess) cudaFree(p);
        }
    };
    thread_local Staging st[8];
    Staging& b = st[(unsigned) device % 8u];
    if (b.dev != device || b.cap < bytes) {
        if (b.p && b.dev >= 0) {
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice(b.dev);
            cudaFree(b.p);
            cudaSetDevice(cur);
        }
        b.p = nullptr, b.cap = 0, b.dev = device;
        TBRM_CUDA(cudaMalloc(&b.p, bytes));
        b.cap = bytes;
    }
    cudaError_t e = enqueue(b.p);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, b.p, bytes, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    TBRM_CUDA(e);
    return TBRM_OK;
}

struct DeviceCounter {  // a device word that is released however the function is left
    unsigned long long* p = nullptr;
    ~DeviceCounter() {
        if (p) cudaFree(p);
    }
};

// runs `enqueue(d_out, d_counter)` on the per-thread stream, delivers `bytes` of output and the optional iteration count
template <typename F>
static tbrm_status mandelbulb_op(int device, void* dst, size_t bytes, int dst_is_device, uint64_t* out_iterations, F enqueue) {
    if (tbrm_device_count() <= 0) {
        set_last_error("no CUDA device visible");
        return TBRM_ERR_NO_DEVICE;
    }
    TBRM_CUDA(cudaSetDevice(device));
    DeviceCounter counter;  // freed on every exit path
    unsigned long long*& d_iters = counter.p;
    if (out_iterations) {
        TBRM_CUDA(cudaMalloc((void**) &d_iters, sizeof(unsigned long long)));
        TBRM_CUDA(cudaMemsetAsync(d_iters, 0, sizeof(unsigned long long), cudaStreamPerThread));
    }
    tbrm_status s = with_output(device, cudaStreamPerThread, dst, bytes, dst_is_device, [&](void* d) { return enqueue(d, d_iters); });
    if (s == TBRM_OK && out_iterations) {
        unsigned long long h = 0;
        cudaError_t e = cudaMemcpyAsync(&h, d_iters, sizeof(h), cudaMemcpyDeviceToHost, cudaStreamPerThread);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);
        *out_iterations = h;
        if (e != cudaSuccess) s = TBRM_ERR_CUDA;
    }
    if (s == TBRM_OK && dst_is_device && cudaStreamSynchronize(cudaStreamPerThread) != cudaSuccess) s = TBRM_ERR_CUDA;
    return s;
}

// ConvertData (VolumeLoader.cpp:97-128) + VoxelFormatToPixelFormat (VolumeInfo.cpp:99-122)
static void converted_format(const tbrm_volume_info& info, int normalize, int convert_to_float, int& texture_format, int& actual_format) {
    const int of = info.original_format;
    if (normalize)
        actual_format = info.bytes_per_voxel > 1 ? TBRM_VOXEL_U16 : TBRM_VOXEL_U8;  // "normalize and cap at G16"
    else if (convert_to_float && of != TBRM_VOXEL_F32)
        actual_format = TBRM_VOXEL_F32;
    else
        actual_format = of;
    switch (actual_format) {
        case TBRM_VOXEL_U8:
        case TBRM_VOXEL_I8: texture_format = TBRM_FMT_G8; break;    // bits as stored
        case TBRM_VOXEL_U16:
        case TBRM_VOXEL_I16: texture_format = TBRM_FMT_G16; break;
        case TBRM_VOXEL_F32: texture_format = TBRM_FMT_R32F; break;
        default: texture_format = -1; break;                        // PF_R32_SINT: "experimental" in the reference, not sampled by the path
    }
}

// the conversions share: staging of a host source, the two scratch buffers, delivery to a host or device destination
struct IngestBuffers {
    void* d_in = nullptr;
    void* d_out = nullptr;
    bool own_in = false, own_out = false;
    ~IngestBuffers() {
        if (own_in && d_in) cudaFree(d_in);
        if (own_out && d_out) cudaFree(d_out);
    }
};
// per host thread and device: the (min, max) partials of the reduction + the final pair (ops run on the per-thread stream, so the
// scratch is per thread too; ~10 KB, kept for the life of the process: a device-to-device conversion allocates nothing)
static void* ingest_scratch(int device) {
    static thread_local void* cache[64] = {};
    if (device < 0 || device >= 64) return nullptr;
    if (!cache[device] && cudaMalloc(&cache[device], ingest_partials_bytes() + 64) != cudaSuccess) cache[device] = nullptr;
    return cache[device];
}

static tbrm_status ingest_stage(IngestBuffers& b, const void* src, int src_is_device, size_t in_bytes, void* dst, int dst_is_device,
                                size_t out_bytes) {
    if (src_is_device) {
        b.d_in = const_cast<void*>(src);
    } else {
        TBRM_CUDA(cudaMalloc(&b.d_in, in_bytes));
        b.own_in = true;
        TBRM_CUDA(cudaMemcpyAsync(b.d_in, src, in_bytes, cudaMemcpyHostToDevice, cudaStreamPerThread));
    }
    if (dst_is_device) {
        b.d_out = dst;
    } else {
        TBRM_CUDA(cudaMalloc(&b.d_out, out_bytes));
        b.own_out = true;
    }
    return TBRM_OK;
}

extern "C" {

int tbrm_abi_version(void) { return TBRM_ABI_VERSION; }

const char* tbrm_status_string(int status) {
    switch (status) {
        case TBRM_OK: return "ok";
        case TBRM_ERR_INVALID_ARGUMENT: return "invalid argument";
        case TBRM_ERR_NOT_INITIALIZED: return "resources not initialized";
        case TBRM_ERR_CUDA: return "CUDA error";
        case TBRM_ERR_UNSUPPORTED: return "unsupported configuration";
        case TBRM_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown status";
    }
}
const char* tbrm_last_error(void) { return t_last_error.c_str(); }

int tbrm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int64_t tbrm_kernel_launch_count(void) { return (int64_t) g_kernel_launches.load(); }

tbrm_status tbrm_plan_dir_light(const int32_t light_dims[3], const tbrm_windowing* win, const tbrm_options* opts,
                                const tbrm_dir_light* light, const tbrm_world* world, tbrm_light_plan* out) {
    TBRM_REQUIRE(light_dims && win && light && world && out, "tbrm_plan_dir_light: null argument");
    TBRM_REQUIRE(light_dims[0] > 0 && light_dims[1] > 0 && light_dims[2] > 0, "tbrm_plan_dir_light: empty light volume");
    host::plan_dir_light(light_dims, *win, opts ? opts->border_exact != 0 : false, *light, *world, *out);
    return TBRM_OK;
}

// ---- resources --------------------------------------------------------------------------------------------
tbrm_status tbrm_create(int device, const int32_t data_dims[3], tbrm_format data_fmt, tbrm_format light_fmt, int half_res,
                        tbrm_resources** out) {
    TBRM_REQUIRE(out && data_dims, "tbrm_create: null argument");
    *out = nullptr;
    // RaymarchVolume.cpp:833-841: a volume with a zero dimension is not initialised
    TBRM_REQUIRE(data_dims[0] > 0 && data_dims[1] > 0 && data_dims[2] > 0, "tbrm_create: data volume has a zero dimension");
    TBRM_REQUIRE(data_fmt == TBRM_FMT_G8 || data_fmt == TBRM_FMT_G16 || data_fmt == TBRM_FMT_R32F, "tbrm_create: bad data format");
    TBRM_REQUIRE(light_fmt == TBRM_FMT_G8 || light_fmt == TBRM_FMT_R32F, "tbrm_create: light volume must be G8 or R32F");
    if (tbrm_device_count() <= 0) {
        set_last_error("tbrm_create: no CUDA device visible");
        return TBRM_ERR_NO_DEVICE;
    }
    TBRM_CUDA(cudaSetDevice(device));
    auto* r = new tbrm_resources();
    r->device = device;
    r->data_fmt = data_fmt;
    r->light_fmt = light_fmt;
    r->half_res = half_res != 0;
    for (int k = 0; k < 3; ++k) {
        r->ddims[k] = data_dims[k];
        r->ldims[k] = half_res ? (data_dims[k] + 1) / 2 : data_dims[k];  // RaymarchVolume.cpp:850-855
    }
    auto fail = [&](cudaError_t e, const char* what) {
        set_last_error(std::string(what) + ": " + cudaGetErrorString(e));
        tbrm_destroy(r);
        return TBRM_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    if ((e = cudaEventCreate(&r->ev_begin)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&r->ev_end)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaMalloc(&r->light, r->light_voxels() * r->light_elem())) != cudaSuccess) return fail(e, "cudaMalloc(light volume)");
    if ((e = cudaMalloc((void**) &r->tf, 256 * sizeof(float4))) != cudaSuccess) return fail(e, "cudaMalloc(tf)");
    if ((e = cudaMalloc((void**) &r->counters, 8 * sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cudaMalloc(counters)");
    // 4 R/W buffers per axis sized X:(Y,Z) Y:(X,Z) Z:(X,Y) in the light pixel format (RaymarchVolume.cpp:864-866,889-891)
    const size_t bsz[3] = {(size_t) r->ldims[1] * r->ldims[2], (size_t) r->ldims[0] * r->ldims[2], (size_t) r->ldims[0] * r->ldims[1]};
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 4; ++b)
            if ((e = cudaMalloc(&r->rw[a][b], bsz[a] * r->light_elem())) != cudaSuccess) return fail(e, "cudaMalloc(rw buffer)");
    // a freshly created render target is cleared (UTextureRenderTargetVolume::Init)
    if ((e = cudaMemsetAsync(r->light, 0, r->light_voxels() * r->light_elem(), r->stream)) != cudaSuccess) return fail(e, "cudaMemset");
    *out = r;
    return TBRM_OK;
}

tbrm_status tbrm_destroy(tbrm_resources* r) {
    if (!r) return TBRM_OK;
    cudaSetDevice(r->device);
    if (r->stream) cudaStreamSynchronize(r->stream);
    if (r->upload_stream) cudaStreamSynchronize(r->upload_stream);
    if (r->download_stream) cudaStreamSynchronize(r->download_stream);
    if (r->data_owned && r->data) cudaFree(r->data);
    if (r->data_back) cudaFree(r->data_back);
    for (int i = 0; i < 2; ++i) {
        if (r->frame_dev[i]) cudaFree(r->frame_dev[i]);
        if (r->ev_frame_done[i]) cudaEventDestroy(r->ev_frame_done[i]);
        if (r->ev_frame_copied[i]) cudaEventDestroy(r->ev_frame_copied[i]);
    }
    if (r->ev_uploaded) cudaEventDestroy(r->ev_uploaded);
    if (r->ev_back_free) cudaEventDestroy(r->ev_back_free);
    if (r->upload_stream) cudaStreamDestroy(r->upload_stream);
    if (r->download_stream) cudaStreamDestroy(r->download_stream);
    if (r->light && r->light_owned) cudaFree(r->light);
    for (int i = 0; i < 2; ++i)
        if (r->peer_arena[i] && r->peer_ipc[i]) cudaIpcCloseMemHandle(r->peer_arena[i]);
    for (int i = 0; i <= tbrm_resources::kMaxPushPeers; ++i)
        if (r->peer_light[i] && r->peer_light_ipc[i]) cudaIpcCloseMemHandle(r->peer_light[i]);
    if (r->arena) cudaFree(r->arena);
    if (r->change_scratch) cudaFree(r->change_scratch);
    if (r->tf) cudaFree(r->tf);
    if (r->counters) cudaFree(r->counters);
    if (r->ring) cudaFree(r->ring);
    if (r->data_yzx) cudaFree(r->data_yzx);
    if (r->tables) cudaFree(r->tables);
    if (r->bricks) cudaFree(r->bricks);
    if (r->joined_buf) cudaFree(r->joined_buf);
    for (int m = 0; m < 4; ++m)
        if (r->octree[m]) cudaFree(r->octree[m]);

    if (r->flags) cudaFree(r->flags);
    if (r->sweep_err) cudaFree(r->sweep_err);
    if (r->light_perm[0]) cudaFree(r->light_perm[0]);
    if (r->light_perm[1]) cudaFree(r->light_perm[1]);
    if (r->tvol) cudaFree(r->tvol);
    if (r->tones) cudaFree(r->tones);
    for (auto& axis : r->rw)
        for (void* b : axis)
            if (b) cudaFree(b);
    if (r->ev_begin) cudaEventDestroy(r->ev_begin);
    if (r->ev_end) cudaEventDestroy(r->ev_end);
    if (r->stream && r->stream_owned) cudaStreamDestroy(r->stream);
    delete r;
    return TBRM_OK;
}

tbrm_status tbrm_set_options(tbrm_resources* r, const tbrm_options* opts) {
    TBRM_REQUIRE(r && opts, "tbrm_set_options: null argument");
    r->options = *opts;
    return TBRM_OK;
}

tbrm_status tbrm_upload_volume(tbrm_resources* r, const void* src, int src_is_device) {
    TBRM_RANGE("tbrm::SetDataVolume");
    TBRM_REQUIRE(r && src, "tbrm_upload_volume: null argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    const size_t bytes = r->data_voxels() * r->data_elem();
    if (!r->data_owned) {
        r->data = nullptr;
        TBRM_CUDA(cudaMalloc(&r->data, bytes));
        r->data_owned = true;
    }
    TBRM_CUDA(cudaMemcpyAsync(r->data, src, bytes, src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, r->stream));
    r->data_ready = true;
    r->data_yzx_valid = false;
    r->bricks_valid = false;
    r->octree_valid = false;  // bRequestedOctreeRebuild (RaymarchVolume.cpp:553-554)
    return TBRM_OK;
}

// ---- streaming upload: back buffer + upload stream ------------------------------------------------------------
static tbrm_status ensure_streaming(tbrm_resources* r) {
    if (!r->upload_stream) TBRM_CUDA(cudaStreamCreateWithFlags(&r->upload_stream, cudaStreamNonBlocking));
    if (!r->download_stream) TBRM_CUDA(cudaStreamCreateWithFlags(&r->download_stream, cudaStreamNonBlocking));
    if (!r->ev_uploaded) TBRM_CUDA(cudaEventCreateWithFlags(&r->ev_uploaded, cudaEventDisableTiming));
    if (!r->ev_back_free) TBRM_CUDA(cudaEventCreateWithFlags(&r->ev_back_free, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        if (!r->ev_frame_done[i]) TBRM_CUDA(cudaEventCreateWithFlags(&r->ev_frame_done[i], cudaEventDisableTiming));
        if (!r->ev_frame_copied[i]) TBRM_CUDA(cudaEventCreateWithFlags(&r->ev_frame_copied[i], cudaEventDisableTiming));
    }
    return TBRM_OK;
}

tbrm_status tbrm_upload_volume_async(tbrm_resources* r, const void* src_host) {
    TBRM_RANGE("tbrm::SetDataVolumeAsync");
    TBRM_REQUIRE(r && src_host, "tbrm_upload_volume_async: null argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    tbrm_status s = ensure_streaming(r);
    if (s != TBRM_OK) return s;
    const size_t bytes = r->data_voxels() * r->data_elem();
    if (!r->data_back) {
        TBRM_CUDA(cudaMalloc(&r->data_back, bytes));
        TBRM_CUDA(cudaEventRecord(r->ev_back_free, r->stream));  // nothing reads a fresh buffer
    }
    // the back buffer was the front buffer until the last swap: wait until the render queue is done with it
    TBRM_CUDA(cudaStreamWaitEvent(r->upload_stream, r->ev_back_free, 0));
    TBRM_CUDA(cudaMemcpyAsync(r->data_back, src_host, bytes, cudaMemcpyHostToDevice, r->upload_stream));
    TBRM_CUDA(cudaEventRecord(r->ev_uploaded, r->upload_stream));
    r->upload_pending = true;
    return TBRM_OK;
}

tbrm_status tbrm_present_volume(tbrm_resources* r) {
    TBRM_RANGE("tbrm::PresentDataVolume");
    TBRM_REQUIRE(r, "tbrm_present_volume: null argument");
    TBRM_REQUIRE(r->upload_pending && r->data_back, "tbrm_present_volume: no tbrm_upload_volume_async since the last present");
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(cudaStreamWaitEvent(r->stream, r->ev_uploaded, 0));  // the render queue continues once the copy has landed
    void* old_front = r->data_owned ? r->data : nullptr;
    r->data = r->data_back;
    r->data_owned = true;
    r->data_back = old_front;  // a caller-bound front buffer is not ours to reuse: the next upload allocates a back buffer
    // everything enqueued so far may still read the old front buffer; later ops read the new one
    TBRM_CUDA(cudaEventRecord(r->ev_back_free, r->stream));
    r->upload_pending = false;
    r->data_ready = true;
    r->data_yzx_valid = false;
    r->bricks_valid = false;
    r->octree_valid = false;  // bRequestedOctreeRebuild (RaymarchVolume.cpp:553-554)
    return TBRM_OK;
}

tbrm_status tbrm_bind_volume_device(tbrm_resources* r, const void* dptr) {
    TBRM_REQUIRE(r && dptr, "tbrm_bind_volume_device: null argument");
    if (r->data_owned && r->data) {
        cudaSetDevice(r->device);
        cudaStreamSynchronize(r->stream);
        cudaFree(r->data);
    }
    r->data = const_cast<void*>(dptr);
    r->data_owned = false;
    r->data_ready = true;
    r->data_yzx_valid = false;
    r->bricks_valid = false;
    r->octree_valid = false;  // bRequestedOctreeRebuild (RaymarchVolume.cpp:553-554)
    return TBRM_OK;
}

// PF_FloatRGBA texels sampled at v = 0.5 by a bilinear clamp sampler collapse to one 256-entry row
static tbrm_status upload_tf_rows(tbrm_resources* r, const float* rgba, int width, int height) {
    float table[256 * 4];
    const float y = 0.5f * (float) height - 0.5f;
    const float fl = floorf(y), fy = y - fl;
    const int j0 = (int) fminf(fmaxf(fl, 0.0f), (float) (height - 1));
    const int j1 = (int) fminf(fmaxf(fl + 1.0f, 0.0f), (float) (height - 1));
    for (int i = 0; i < 256; ++i)
        for (int c = 0; c < 4; ++c) {
            const float a = __half2float(__float2half_rn(rgba[((size_t) j0 * width + i) * 4 + c]));
            const float b = __half2float(__float2half_rn(rgba[((size_t) j1 * width + i) * 4 + c]));
            table[4 * i + c] = fmaf(fy, b - a, a);
        }
    TBRM_CUDA(cudaSetDevice(r->device));
    // the table lives on the stack: synchronous w.r.t. the host, ordered on the stream
    TBRM_CUDA(cudaMemcpyAsync(r->tf, table, sizeof(table), cudaMemcpyHostToDevice, r->stream));
    TBRM_CUDA(cudaStreamSynchronize(r->stream));
    r->tf_ready = true;
    return TBRM_OK;
}

tbrm_status tbrm_set_transfer_function(tbrm_resources* r, const float* rgba, int width, int height) {
    TBRM_REQUIRE(r && rgba, "tbrm_set_transfer_function: null argument");
    TBRM_REQUIRE(width == 256 && height >= 1, "tbrm_set_transfer_function: the TF texture is 256 x H (RaymarchUtils.cpp:145-148)");
    return upload_tf_rows(r, rgba, width, height);
}

tbrm_status tbrm_make_default_tf(tbrm_resources* r) {
    TBRM_REQUIRE(r, "tbrm_make_default_tf: null argument");
    float samples[256 * 4];
    for (unsigned i = 0; i < 256; ++i) {  // RaymarchUtils.cpp:121-128
        const float whiteness = (float) i / (float) (256 - 1);
        samples[4 * i] = samples[4 * i + 1] = samples[4 * i + 2] = whiteness;
        samples[4 * i + 3] = 1.0f;
    }
    return upload_tf_rows(r, samples, 256, 1);
}

tbrm_status tbrm_set_windowing(tbrm_resources* r, const tbrm_windowing* w) {
    TBRM_REQUIRE(r && w, "tbrm_set_windowing: null argument");
    r->windowing = *w;
    return TBRM_OK;
}

// ---- sweep ------------------------------------------------------------------------------------------------
tbrm_status tbrm_clear_light_volume(tbrm_resources* r, float clear_value) {
    TBRM_RANGE("tbrm::ClearLightVolume");
    if (!r || !r->light) return TBRM_OK;  // RaymarchUtils.cpp:106-109: silently returns without a render target
    TBRM_CUDA(cudaSetDevice(r->device));
    if (r->slab.nranks > 1) {  // a sharded volume: every rank clears the slab it owns
        const size_t plane = (size_t) r->ldims[0] * r->ldims[1];
        TBRM_CUDA(sweep_fill_buffer(*r, (char*) r->light + plane * r->slab.z_begin * r->light_elem(),
                                    plane * (size_t) (r->slab.z_end - r->slab.z_begin), clear_value));
        return TBRM_OK;
    }
    TBRM_CUDA(clear_light(*r, clear_value));
    return TBRM_OK;
}

static bool resources_valid(const tbrm_resources* r) {  // the 7 null checks of RaymarchUtils.cpp:39-41
    return r && r->data && r->data_ready && r->tf && r->tf_ready && r->light;
}

static void fill_pass(LightPass& lp, const tbrm_pass_plan& p) {
    lp.uv_off[0] = p.uv_offset[0], lp.uv_off[1] = p.uv_offset[1];
    for (int k = 0; k < 3; ++k) lp.uvw_off[k] = p.uvw_offset[k];
    lp.step = p.step_size * 100.0f;  // StepSize * VOLUME_DENSITY (AddDirLightShader.usf:112)
    lp.border = p.border;
    lp.light_alpha = p.light_alpha;
}

static void fill_uniforms(const tbrm_resources& r, const tbrm_light_plan& plan, int pass, SweepUniforms& u) {
    const tbrm_pass_plan& p = plan.pass[pass];
    u.axis = p.axis, u.dirn = p.dirn, u.start = p.start;
    for (int k = 0; k < 3; ++k) {
        u.td[k] = p.td[k];
        u.ldims[k] = r.ldims[k];
        u.ddims[k] = r.ddims[k];
        u.clip_center[k] = plan.clip_center[k];
        u.clip_dir[k] = plan.clip_dir[k];
    }
    u.data_border = plan.data_border;
    u.win = Windowing{r.windowing.center, r.windowing.width, r.windowing.low_cutoff ? 1.0f : 0.0f, r.windowing.high_cutoff ? 1.0f : 0.0f};
    u.sign = 1.0f;
    fill_pass(u.a, p);
    u.r = u.a;
    u.gate_saturate = 1;
}

static tbrm_status run_pass(tbrm_resources& r, const SweepUniforms& u, bool change, int gpu_sync, tbrm_sweep_stats* stats) {
    int launches = 0;
    bool handled = false;
    int used = 1;
    // 0 auto (gpu_sync ? fused : per-slice), 1 per-slice, 2 fused (TMA path when eligible), 3 generic fused only
    const int impl = r.options.sweep_impl;
    const bool sharded = r.slab.nranks > 1;
    const bool want_fused = impl == 2 || impl == 3 || (impl == 0 && gpu_sync) || sharded;
    if (want_fused && (impl != 3 || sharded)) {
        TBRM_CUDA(sweep_pass_tma(r, u, change, &launches, &handled));
        if (handled) used = 3;
    }
    if (sharded && !handled) {  // only the TMA-staged sweep knows how to exchange light between slabs
        set_last_error(std::string("sharded volume: only the TMA-staged sweep exchanges light between slabs; ") + tbrm_last_error());
        return TBRM_ERR_UNSUPPORTED;
    }
    if (want_fused && !handled) {
        TBRM_CUDA(sweep_pass_fused(r, u, change, &launches, &handled));
        if (handled) used = 2;
        if (!handled && impl != 0) {
            set_last_error("fused sweep does not support this configuration");
            return TBRM_ERR_UNSUPPORTED;
        }
    }
    if (!handled) TBRM_CUDA(sweep_pass_per_slice(r, u, change, &launches));
    if (stats) {
        if (stats->passes < 4) {
            stats->faces[stats->passes] = u.axis * 2 + (u.dirn > 0 ? 1 : 0);
            stats->impl[stats->passes] = used;
        }
        stats->passes += 1;
        stats->voxels += (int64_t) u.td[0] * u.td[1] * u.td[2];
        stats->kernel_launches += launches;
    }
    return TBRM_OK;
}

// AddDirLightToSingleLightVolume_RenderThread — LightingShaders.cpp:35-166
static tbrm_status add_dir_light_impl(tbrm_resources& r, const tbrm_dir_light& light, bool added, const tbrm_world& world, int gpu_sync,
                                      tbrm_sweep_stats* stats, int only_pass = -1) {
    tbrm_light_plan plan;
    host::plan_dir_light(r.ldims, r.windowing, r.options.border_exact != 0, light, world, plan);
    if (plan.zero_direction) return TBRM_OK;  // :41-46
    for (int i = 0; i < plan.add_passes; ++i) {  // "break if the axis weight == 0", :65-68, :94-97
        if (only_pass >= 0 && i != only_pass) continue;
        SweepUniforms u;
        fill_uniforms(r, plan, i, u);
        u.sign = added ? 1.0f : -1.0f;
        // push-gather (tbrm_slab_push_light): the call's last axis pass leaves final light values in every brick it stores
        r.push_this_pass = r.push_light && only_pass < 0 && i == plan.add_passes - 1;
        tbrm_status s = run_pass(r, u, false, gpu_sync, stats);
        r.push_this_pass = false;
        if (s != TBRM_OK) return s;
    }
    return TBRM_OK;
}

// Same-face passes of several lights in one sweep (not in the reference — Readme.md:165-166, 186-187). Lights are
// planned like AddDirLight; passes are grouped by cube face in order of first appearance (at most kMaxJoined per group).
static tbrm_status add_dir_lights_joined_impl(tbrm_resources& r, const tbrm_dir_light* lights, int n_lights, bool added, const tbrm_world& world,
                                              int* lights_added, tbrm_sweep_stats* stats) {
    struct Group {
        SweepUniforms u;
        int face;
        std::vector<LightPass> members;
    };
    std::vector<Group> groups;
    int n_added = 0;
    for (int i = 0; i < n_lights; ++i) {
        tbrm_light_plan plan;
        host::plan_dir_light(r.ldims, r.windowing, r.options.border_exact != 0, lights[i], world, plan);
        if (plan.zero_direction) continue;  // LightingShaders.cpp:41-46
        ++n_added;
        for (int p = 0; p < plan.add_passes; ++p) {
            size_t g = 0;
            while (g < groups.size() && !(groups[g].face == plan.pass[p].face && (int) groups[g].members.size() < kMaxJoined)) ++g;
            if (g == groups.size()) {
                Group G;
                fill_uniforms(r, plan, p, G.u);
                G.u.sign = added ? 1.0f : -1.0f;
                G.face = plan.pass[p].face;
                groups.push_back(G);
            }
            LightPass lp;
            fill_pass(lp, plan.pass[p]);
            groups[g].members.push_back(lp);
        }
    }
    if (lights_added) *lights_added = n_added;
    for (Group& G : groups) {
        int launches = 0;
        TBRM_CUDA(sweep_pass_joined(r, G.u, G.members.data(), (int) G.members.size(), &launches));
        if (stats) {
            if (stats->passes < 4) {
                stats->faces[stats->passes] = G.face;
                stats->impl[stats->passes] = 4;
            }
            stats->passes += 1;
            stats->voxels += (int64_t) G.u.td[0] * G.u.td[1] * G.u.td[2] * (int64_t) G.members.size();
            stats->kernel_launches += launches;
        }
    }
    return TBRM_OK;
}

static void reset_stats(tbrm_sweep_stats* s) {
    if (!s) return;
    memset(s, 0, sizeof(*s));
    for (int& f : s->faces) f = -1;
}

tbrm_status tbrm_add_dir_light_stats(tbrm_resources* r, const tbrm_dir_light* light, int added, const tbrm_world* world,
                                     int* light_added, int gpu_sync, tbrm_sweep_stats* stats) {
    TBRM_RANGE("tbrm::AddDirLight");
    reset_stats(stats);
    if (!resources_valid(r)) {  // RaymarchUtils.cpp:39-45
        if (light_added) *light_added = 0;
        return TBRM_ERR_NOT_INITIALIZED;
    }
    TBRM_REQUIRE(light && world, "tbrm_add_dir_light: null light or world parameters");
    if (light_added) *light_added = 1;  // :48
    TBRM_CUDA(cudaSetDevice(r->device));
    return add_dir_light_impl(*r, *light, added != 0, *world, gpu_sync, stats);
}

tbrm_status tbrm_add_dir_light(tbrm_resources* r, const tbrm_dir_light* light, int added, const tbrm_world* world, int* light_added,
                               int gpu_sync) {
    return tbrm_add_dir_light_stats(r, light, added, world, light_added, gpu_sync, nullptr);
}

// ChangeDirLightInSingleLightVolume_RenderThread — LightingShaders.cpp:168-326
tbrm_status tbrm_add_dir_lights_joined(tbrm_resources* r, const tbrm_dir_light* lights, int n_lights, int added, const tbrm_world* world,
                                       int* lights_added, tbrm_sweep_stats* stats) {
    TBRM_RANGE("tbrm::AddDirLightsJoined");
    reset_stats(stats);
    if (lights_added) *lights_added = 0;
    if (!resources_valid(r)) return TBRM_ERR_NOT_INITIALIZED;
    TBRM_REQUIRE((lights || n_lights == 0) && n_lights >= 0 && world, "tbrm_add_dir_lights_joined: null argument");
    if (r->slab.nranks > 1) {
        set_last_error("tbrm_add_dir_lights_joined: joined sweeps run the per-slice schedule, which does not exchange light between slabs");
        return TBRM_ERR_UNSUPPORTED;
    }
    TBRM_CUDA(cudaSetDevice(r->device));
    return add_dir_lights_joined_impl(*r, lights, n_lights, added != 0, *world, lights_added, stats);
}

tbrm_status tbrm_change_dir_light_stats(tbrm_resources* r, const tbrm_dir_light* old_light, const tbrm_dir_light* new_light,
                                        const tbrm_world* world, int* light_added, int gpu_sync, tbrm_sweep_stats* stats) {
    TBRM_RANGE("tbrm::ChangeDirLight");
    reset_stats(stats);
    if (!resources_valid(r)) {  // RaymarchUtils.cpp:74-80
        if (light_added) *light_added = 0;
        return TBRM_ERR_NOT_INITIALIZED;
    }
    TBRM_REQUIRE(old_light && new_light && world, "tbrm_change_dir_light: null light or world parameters");
    if (light_added) *light_added = 1;
    TBRM_CUDA(cudaSetDevice(r->device));
    tbrm_light_plan rem, add;
    host::plan_dir_light(r->ldims, r->windowing, r->options.border_exact != 0, *old_light, *world, rem);
    host::plan_dir_light(r->ldims, r->windowing, r->options.border_exact != 0, *new_light, *world, add);
    if (rem.zero_direction || add.zero_direction) return TBRM_OK;  // :173-179
    if (rem.pass[0].face != add.pass[0].face || rem.pass[1].face != add.pass[1].face) {  // :192-198
        if (stats) stats->fell_back = 1;
        tbrm_status s = add_dir_light_impl(*r, *old_light, false, *world, gpu_sync, stats);
        if (s != TBRM_OK) return s;
        return add_dir_light_impl(*r, *new_light, true, *world, gpu_sync, stats);
    }
    for (int i = 0; i < 2; ++i) {  // both axes, no weight test (:203, :238)
        SweepUniforms u;
        fill_uniforms(*r, rem, i, u);
        fill_pass(u.r, rem.pass[i]);
        fill_pass(u.a, add.pass[i]);
        u.gate_saturate = 0;  // ChangeDirLightShader.usf:130,136 has no saturate gate
        tbrm_status s = run_pass(*r, u, true, gpu_sync, stats);
        if (s != TBRM_OK) return s;
    }
    return TBRM_OK;
}

tbrm_status tbrm_change_dir_light(tbrm_resources* r, const tbrm_dir_light* old_light, const tbrm_dir_light* new_light,
                                  const tbrm_world* world, int* light_added, int gpu_sync) {
    return tbrm_change_dir_light_stats(r, old_light, new_light, world, light_added, gpu_sync, nullptr);
}

tbrm_status tbrm_light_volume_dims(const tbrm_resources* r, int32_t dims[3]) {
    TBRM_REQUIRE(r && dims, "tbrm_light_volume_dims: null argument");
    for (int k = 0; k < 3; ++k) dims[k] = r->ldims[k];
    return TBRM_OK;
}

tbrm_status tbrm_download_light_volume(tbrm_resources* r, void* dst_host) {
    TBRM_REQUIRE(r && dst_host, "tbrm_download_light_volume: null argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(cudaMemcpyAsync(dst_host, r->light, r->light_voxels() * r->light_elem(), cudaMemcpyDeviceToHost, r->stream));
    TBRM_CUDA(cudaStreamSynchronize(r->stream));
    return TBRM_OK;
}

tbrm_status tbrm_upload_light_volume(tbrm_resources* r, const void* src_host) {
    TBRM_REQUIRE(r && src_host, "tbrm_upload_light_volume: null argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(cudaMemcpyAsync(r->light, src_host, r->light_voxels() * r->light_elem(), cudaMemcpyHostToDevice, r->stream));
    TBRM_CUDA(cudaStreamSynchronize(r->stream));
    return TBRM_OK;
}

tbrm_status tbrm_bind_light_volume_device(tbrm_resources* r, void* dptr) {
    TBRM_REQUIRE(r && dptr, "tbrm_bind_light_volume_device: null argument");
    TBRM_REQUIRE(((uintptr_t) dptr & 15) == 0, "tbrm_bind_light_volume_device: the light volume must be 16-byte aligned");
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(cudaStreamSynchronize(r->stream));
    if (r->light && r->light_owned) cudaFree(r->light);
    r->light = dptr;
    r->light_owned = false;
    return TBRM_OK;
}

// ---- Z-slab sharding ----------------------------------------------------------------------------------------
void tbrm_slab_partition(int32_t z_slices, int32_t nranks, int32_t rank, int32_t* z_begin, int32_t* z_end) {
    if (!z_begin || !z_end || nranks <= 0 || rank < 0 || rank >= nranks) return;
    slab_partition(z_slices, nranks, rank, z_begin, z_end);
}

tbrm_status tbrm_slab_configure(tbrm_resources* r, const tbrm_slab* slab) {
    TBRM_REQUIRE(r && slab, "tbrm_slab_configure: null argument");
    TBRM_REQUIRE(slab->nranks >= 1 && slab->rank >= 0 && slab->rank < slab->nranks, "tbrm_slab_configure: bad rank");
    if (slab->nranks > 1) {
        int32_t zb, ze;
        slab_partition(r->ldims[2], slab->nranks, slab->rank, &zb, &ze);
        TBRM_REQUIRE(zb == slab->z_begin && ze == slab->z_end && zb < ze, "tbrm_slab_configure: the slab must follow tbrm_slab_partition and be non-empty");
        const bool l8 = r->light_fmt == TBRM_FMT_G8;  // byte bricks: the light volume's own X and Y are TMA strides
        if (r->data_fmt != TBRM_FMT_G8 || (r->light_fmt != TBRM_FMT_R32F && !l8) || r->ddims[0] % 16 || r->ddims[1] % 16 || r->ddims[2] % 8 || r->ldims[2] % 8 ||
            (l8 && (r->ldims[0] % 16 || r->ldims[1] % 16))) {
            set_last_error("tbrm_slab_configure: sharding needs R8 data, an R32F or G8 light volume, X % 16 == 0, Y % 16 == 0, Z % 8 == 0 (G8: of the light volume too)");
            return TBRM_ERR_UNSUPPORTED;
        }
    }
    TBRM_CUDA(cudaSetDevice(r->device));
    r->slab = *slab;
    if (slab->nranks > 1) TBRM_CUDA(slab_ensure_arena(*r));
    return TBRM_OK;
}

tbrm_status tbrm_slab_arena(tbrm_resources* r, void** dptr, size_t* bytes) {
    TBRM_REQUIRE(r && dptr && bytes, "tbrm_slab_arena: null argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(slab_ensure_arena(*r));
    *dptr = r->arena, *bytes = r->arena_bytes;
    return TBRM_OK;
}

tbrm_status tbrm_slab_ipc_handle(tbrm_resources* r, void* handle64) {
    TBRM_REQUIRE(r && handle64, "tbrm_slab_ipc_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(slab_ensure_arena(*r));
    TBRM_CUDA(cudaStreamSynchronize(r->stream));  // the arena has been cleared before anybody can write to it
    cudaIpcMemHandle_t h;
    TBRM_CUDA(cudaIpcGetMemHandle(&h, r->arena));
    memcpy(handle64, &h, sizeof(h));
    return TBRM_OK;
}

static tbrm_status drop_peer(tbrm_resources* r, int i) {
    if (r->peer_arena[i] && r->peer_ipc[i]) TBRM_CUDA(cudaIpcCloseMemHandle(r->peer_arena[i]));
    r->peer_arena[i] = nullptr, r->peer_ipc[i] = false;
    return TBRM_OK;
}

tbrm_status tbrm_slab_open_peer(tbrm_resources* r, int side, const void* handle64) {
    TBRM_REQUIRE(r && handle64 && (side == -1 || side == 1), "tbrm_slab_open_peer: bad argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    const int i = side < 0 ? 0 : 1;
    tbrm_status s = drop_peer(r, i);
    if (s != TBRM_OK) return s;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    TBRM_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    r->peer_arena[i] = p, r->peer_ipc[i] = true;
    return TBRM_OK;
}

tbrm_status tbrm_slab_set_peer(tbrm_resources* r, int side, void* peer_arena_dptr) {
    TBRM_REQUIRE(r && (side == -1 || side == 1), "tbrm_slab_set_peer: bad argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    const int i = side < 0 ? 0 : 1;
    tbrm_status s = drop_peer(r, i);
    if (s != TBRM_OK) return s;
    r->peer_arena[i] = peer_arena_dptr;
    return TBRM_OK;
}

// ---- push-gather: every rank's light volume mapped into every other rank ----------------------------------------------------
tbrm_status tbrm_slab_light_ipc_handle(tbrm_resources* r, void* handle64) {
    TBRM_REQUIRE(r && handle64, "tbrm_slab_light_ipc_handle: null argument");
    TBRM_REQUIRE(r->light && r->light_owned, "tbrm_slab_light_ipc_handle: the light volume must be the library's own allocation (an IPC handle names a whole allocation)");
    TBRM_CUDA(cudaSetDevice(r->device));
    cudaIpcMemHandle_t h;
    TBRM_CUDA(cudaIpcGetMemHandle(&h, r->light));
    memcpy(handle64, &h, sizeof(h));
    return TBRM_OK;
}

static tbrm_status drop_peer_light(tbrm_resources* r, int rank) {
    if (r->peer_light[rank] && r->peer_light_ipc[rank]) TBRM_CUDA(cudaIpcCloseMemHandle(r->peer_light[rank]));
    r->peer_light[rank] = nullptr, r->peer_light_ipc[rank] = false;
    return TBRM_OK;
}

tbrm_status tbrm_slab_open_peer_light(tbrm_resources* r, int peer_rank, const void* handle64) {
    TBRM_REQUIRE(r && handle64 && peer_rank >= 0 && peer_rank <= tbrm_resources::kMaxPushPeers, "tbrm_slab_open_peer_light: bad argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    tbrm_status s = drop_peer_light(r, peer_rank);
    if (s != TBRM_OK) return s;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    TBRM_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    r->peer_light[peer_rank] = p, r->peer_light_ipc[peer_rank] = true;
    return TBRM_OK;
}

tbrm_status tbrm_slab_set_peer_light(tbrm_resources* r, int peer_rank, void* peer_light_dptr) {
    TBRM_REQUIRE(r && peer_rank >= 0 && peer_rank <= tbrm_resources::kMaxPushPeers, "tbrm_slab_set_peer_light: bad argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    tbrm_status s = drop_peer_light(r, peer_rank);
    if (s != TBRM_OK) return s;
    r->peer_light[peer_rank] = peer_light_dptr;
    return TBRM_OK;
}

tbrm_status tbrm_slab_push_light(tbrm_resources* r, int enable) {
    TBRM_REQUIRE(r, "tbrm_slab_push_light: null argument");
    r->push_light = enable != 0;
    return TBRM_OK;
}

tbrm_status tbrm_slab_reset_comm(tbrm_resources* r) {
    TBRM_REQUIRE(r, "tbrm_slab_reset_comm: null argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(cudaStreamSynchronize(r->stream));
    if (r->arena) TBRM_CUDA(cudaMemsetAsync(r->arena, 0, r->arena_bytes, r->stream));
    if (r->ring) TBRM_CUDA(cudaMemsetAsync(r->ring, 0, r->ring_bytes, r->stream));
    TBRM_CUDA(cudaStreamSynchronize(r->stream));
    r->pass_seq = 0;
    return TBRM_OK;
}

tbrm_status tbrm_slab_check(tbrm_resources* r) {
    TBRM_REQUIRE(r, "tbrm_slab_check: null argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(cudaStreamSynchronize(r->stream));
    if (r->sweep_err) {  // unsharded fused sweeps: a tile that waited on another tile longer than the timeout
        unsigned int werr = 0;
        TBRM_CUDA(cudaMemcpy(&werr, r->sweep_err, sizeof(werr), cudaMemcpyDeviceToHost));
        if (werr) {
            TBRM_CUDA(cudaMemset(r->sweep_err, 0, sizeof(werr)));
            set_last_error("fused sweep timed out waiting for another tile of the same launch (results are void)");
            return TBRM_ERR_CUDA;
        }
    }
    if (!r->arena) return TBRM_OK;
    unsigned int err = 0;
    TBRM_CUDA(cudaMemcpy(&err, (unsigned int*) r->arena + 2, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) {
        TBRM_CUDA(cudaMemset((unsigned int*) r->arena + 2, 0, sizeof(err)));
        set_last_error("slab exchange timed out: a neighbouring slab did not deliver its light (results are void)");
        return TBRM_ERR_CUDA;
    }
    return TBRM_OK;
}

tbrm_status tbrm_slab_set_timeout_ms(tbrm_resources* r, int timeout_ms) {
    TBRM_REQUIRE(r && timeout_ms >= 0, "tbrm_slab_set_timeout_ms: bad argument");
    r->slab_timeout_ms = timeout_ms;
    return TBRM_OK;
}

tbrm_status tbrm_add_dir_light_pass(tbrm_resources* r, const tbrm_dir_light* light, int added, const tbrm_world* world, int pass,
                                    int gpu_sync, tbrm_sweep_stats* stats) {
    TBRM_RANGE("tbrm::AddDirLightPass");
    reset_stats(stats);
    if (!resources_valid(r)) return TBRM_ERR_NOT_INITIALIZED;
    TBRM_REQUIRE(light && world && (pass == 0 || pass == 1), "tbrm_add_dir_light_pass: bad argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    return add_dir_light_impl(*r, *light, added != 0, *world, gpu_sync, stats, pass);
}

tbrm_status tbrm_slab_pass_order(tbrm_resources* r, const tbrm_dir_light* light, const tbrm_world* world, int pass, int* order) {
    TBRM_REQUIRE(r && light && world && order && (pass == 0 || pass == 1), "tbrm_slab_pass_order: bad argument");
    *order = 0;
    tbrm_light_plan plan;
    host::plan_dir_light(r->ldims, r->windowing, r->options.border_exact != 0, *light, *world, plan);
    if (plan.zero_direction || pass >= plan.add_passes) return TBRM_OK;
    SweepUniforms u;
    fill_uniforms(*r, plan, pass, u);
    *order = slab_pass_order(u);
    return TBRM_OK;
}

void* tbrm_light_volume_device_ptr(tbrm_resources* r) { return r ? r->light : nullptr; }
void* tbrm_data_volume_device_ptr(tbrm_resources* r) { return r ? r->data : nullptr; }

// ---- raymarch ---------------------------------------------------------------------------------------------
static bool camera_valid(const tbrm_camera* cam) { return cam && cam->width > 0 && cam->height > 0 && cam->hfov_deg > 0.0; }

tbrm_status tbrm_raymarch_cube_setup(tbrm_resources* r, const tbrm_camera* cam, const tbrm_world* world, float* out, int out_is_device) {
    TBRM_REQUIRE(r && world && out, "tbrm_raymarch_cube_setup: null argument");
    TBRM_REQUIRE(camera_valid(cam), "tbrm_raymarch_cube_setup: invalid camera");
    TBRM_CUDA(cudaSetDevice(r->device));
    host::CameraUniforms cu;
    host::plan_camera(*cam, *world, cu);
    const size_t bytes = (size_t) cam->width * cam->height * 4 * sizeof(float);
    return with_output(r->device, r->stream, out, bytes, out_is_device,
                       [&](void* d) { return raymarch_cube_setup(*r, cu, (float*) d); });
}

static tbrm_status raymarch_lit_impl(tbrm_resources* r, const tbrm_camera* cam, const tbrm_world* world, float step_count, int row_begin,
                                     int row_end, int row_block, int block_stride, float* out_rgba, int out_is_device, uint64_t* out_steps) {
    if (!resources_valid(r)) return TBRM_ERR_NOT_INITIALIZED;
    TBRM_REQUIRE(world && out_rgba, "tbrm_raymarch_lit: null argument");
    TBRM_REQUIRE(camera_valid(cam), "tbrm_raymarch_lit: invalid camera");
    TBRM_REQUIRE(step_count > 0.0f, "tbrm_raymarch_lit: step count must be positive");
    TBRM_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= cam->height, "tbrm_raymarch_lit: bad row range");
    const int rows = raymarch_local_rows(row_begin, row_end, row_block, block_stride);
    if (rows == 0) {
        if (out_steps) *out_steps = 0;
        return TBRM_OK;
    }
    TBRM_CUDA(cudaSetDevice(r->device));
    host::CameraUniforms cu;
    host::plan_camera(*cam, *world, cu);
    float cc[3], cd[3];
    host::plan_clip(*world, cc, cd);  // SetMaterialClippingParameters, RaymarchVolume.cpp:705-728
    unsigned long long* d_steps = out_steps ? r->counters : nullptr;
    if (d_steps) TBRM_CUDA(cudaMemsetAsync(d_steps, 0, sizeof(unsigned long long), r->stream));
    const size_t bytes = (size_t) cam->width * rows * 4 * sizeof(float);
    tbrm_status s = with_output(r->device, r->stream, out_rgba, bytes, out_is_device, [&](void* d) {
        return raymarch_lit(*r, cu, cc, cd, step_count, row_begin, row_end, row_block, block_stride, (float*) d, d_steps);
    });
    if (s != TBRM_OK) return s;
    if (out_steps) {
        unsigned long long h = 0;
        TBRM_CUDA(cudaMemcpyAsync(&h, d_steps, sizeof(h), cudaMemcpyDeviceToHost, r->stream));
        TBRM_CUDA(cudaStreamSynchronize(r->stream));
        *out_steps = h;
    }
    return TBRM_OK;
}

tbrm_status tbrm_raymarch_lit(tbrm_resources* r, const tbrm_camera* cam, const tbrm_world* world, float step_count, int row_begin,
                              int row_end, float* out_rgba, int out_is_device, uint64_t* out_steps) {
    TBRM_RANGE("tbrm::LitRaymarch");
    return raymarch_lit_impl(r, cam, world, step_count, row_begin, row_end, 8, 1, out_rgba, out_is_device, out_steps);
}

tbrm_status tbrm_raymarch_lit_interleaved(tbrm_resources* r, const tbrm_camera* cam, const tbrm_world* world, float step_count,
                                          int block_rows, int first_block, int block_stride, float* out_rgba, int out_is_device,
                                          uint64_t* out_steps) {
    TBRM_RANGE("tbrm::LitRaymarchInterleaved");
    TBRM_REQUIRE(cam && block_rows > 0 && block_rows % 8 == 0 && first_block >= 0 && block_stride >= 1 && first_block < block_stride,
                 "tbrm_raymarch_lit_interleaved: block_rows must be a positive multiple of 8, 0 <= first_block < block_stride");
    const int row_begin = std::min(first_block * block_rows, cam->height);
    return raymarch_lit_impl(r, cam, world, step_count, row_begin, cam->height, block_rows, block_stride, out_rgba, out_is_device, out_steps);
}

tbrm_status tbrm_raymarch_lit_to_host_async(tbrm_resources* r, const tbrm_camera* cam, const tbrm_world* world, float step_count,
                                            float* out_host) {
    TBRM_RANGE("tbrm::LitRaymarchToHostAsync");
    if (!resources_valid(r)) return TBRM_ERR_NOT_INITIALIZED;
    TBRM_REQUIRE(world && out_host && camera_valid(cam), "tbrm_raymarch_lit_to_host_async: bad argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    tbrm_status s = ensure_streaming(r);
    if (s != TBRM_OK) return s;
    const size_t bytes = (size_t) cam->width * cam->height * 4 * sizeof(float);
    if (r->frame_bytes < bytes) {
        TBRM_CUDA(cudaStreamSynchronize(r->download_stream));
        TBRM_CUDA(cudaStreamSynchronize(r->stream));
        for (int i = 0; i < 2; ++i) {
            if (r->frame_dev[i]) cudaFree(r->frame_dev[i]);
            r->frame_dev[i] = nullptr, r->frame_in_flight[i] = false;
            TBRM_CUDA(cudaMalloc(&r->frame_dev[i], bytes));
        }
        r->frame_bytes = bytes;
    }
    const int f = r->frame_next;
    r->frame_next ^= 1;
    // the device frame is reused every second call: its previous copy to the host must have finished
    if (r->frame_in_flight[f]) TBRM_CUDA(cudaStreamWaitEvent(r->stream, r->ev_frame_copied[f], 0));
    s = raymarch_lit_impl(r, cam, world, step_count, 0, cam->height, 8, 1, (float*) r->frame_dev[f], 1, nullptr);
    if (s != TBRM_OK) return s;
    TBRM_CUDA(cudaEventRecord(r->ev_frame_done[f], r->stream));
    TBRM_CUDA(cudaStreamWaitEvent(r->download_stream, r->ev_frame_done[f], 0));
    TBRM_CUDA(cudaMemcpyAsync(out_host, r->frame_dev[f], bytes, cudaMemcpyDeviceToHost, r->download_stream));
    TBRM_CUDA(cudaEventRecord(r->ev_frame_copied[f], r->download_stream));
    r->frame_in_flight[f] = true;
    return TBRM_OK;
}

tbrm_status tbrm_download_wait(tbrm_resources* r) {
    TBRM_REQUIRE(r, "tbrm_download_wait: null argument");
    if (!r->download_stream) return TBRM_OK;
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(cudaStreamSynchronize(r->download_stream));
    return TBRM_OK;
}

int tbrm_raymarch_interleaved_rows(int height, int block_rows, int first_block, int block_stride) {
    if (height <= 0 || block_rows <= 0 || block_stride < 1 || first_block < 0) return 0;
    return raymarch_local_rows(std::min(first_block * block_rows, height), height, block_rows, block_stride);
}

tbrm_status tbrm_mandelbulb_march(int device, const tbrm_mandelbulb* params, const tbrm_camera* cam, const tbrm_world* world,
                                  int row_begin, int row_end, float* out_xy, int out_is_device, uint64_t* out_iterations) {
    TBRM_RANGE("tbrm::MandelbulbMarch");
    TBRM_REQUIRE(params && world && out_xy, "tbrm_mandelbulb_march: null argument");
    TBRM_REQUIRE(camera_valid(cam), "tbrm_mandelbulb_march: invalid camera");
    TBRM_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= cam->height, "tbrm_mandelbulb_march: bad row range");
    TBRM_REQUIRE(params->extent != 0.0f, "tbrm_mandelbulb_march: extent must be non-zero");
    if (tbrm_device_count() <= 0) {
        set_last_error("tbrm_mandelbulb_march: no CUDA device visible");
        return TBRM_ERR_NO_DEVICE;
    }
    if (row_begin == row_end) {
        if (out_iterations) *out_iterations = 0;
        return TBRM_OK;
    }
    TBRM_CUDA(cudaSetDevice(device));
    host::CameraUniforms cu;
    host::plan_camera(*cam, *world, cu);
    DeviceCounter counter;  // freed on every exit path
    unsigned long long*& d_iters = counter.p;
    if (out_iterations) {
        TBRM_CUDA(cudaMalloc((void**) &d_iters, sizeof(unsigned long long)));
        TBRM_CUDA(cudaMemsetAsync(d_iters, 0, sizeof(unsigned long long), cudaStreamPerThread));
    }
    const size_t bytes = (size_t) cam->width * (row_end - row_begin) * 2 * sizeof(float);
    tbrm_status s = with_output(device, cudaStreamPerThread, out_xy, bytes, out_is_device, [&](void* d) {
        return mandelbulb_march(cudaStreamPerThread, *params, cu, row_begin, row_end, (float*) d, d_iters);
    });
    if (s == TBRM_OK && out_iterations) {
        unsigned long long h = 0;
        cudaError_t e = cudaMemcpyAsync(&h, d_iters, sizeof(h), cudaMemcpyDeviceToHost, cudaStreamPerThread);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);
        *out_iterations = h;
        if (e != cudaSuccess) s = TBRM_ERR_CUDA;
    }
    if (s == TBRM_OK && out_is_device) {
        if (cudaStreamSynchronize(cudaStreamPerThread) != cudaSuccess) s = TBRM_ERR_CUDA;
    }
    return s;
}

// ---- the other materials and the octree ----------------------------------------------
tbrm_status tbrm_generate_octree(tbrm_resources* r) {
    TBRM_RANGE("tbrm::GenerateOctree");
    // URaymarchUtils::GenerateOctree enqueues without checks (RaymarchUtils.cpp:94-102); the shader needs the data volume
    if (!r || !r->data || !r->data_ready) return TBRM_ERR_NOT_INITIALIZED;
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(generate_octree(*r));
    return TBRM_OK;
}

tbrm_status tbrm_octree_mip_dims(const tbrm_resources* r, int mip, int32_t dims[3]) {
    TBRM_REQUIRE(r && dims && mip >= 0 && mip < 4, "tbrm_octree_mip_dims: bad argument (4 mips)");
    octree_mip_dims(*r, mip, dims);
    return TBRM_OK;
}

tbrm_status tbrm_download_octree_mip(tbrm_resources* r, int mip, void* dst_host) {
    TBRM_REQUIRE(r && dst_host && mip >= 0 && mip < 4, "tbrm_download_octree_mip: bad argument (4 mips)");
    if (!r->octree_valid) return TBRM_ERR_NOT_INITIALIZED;
    TBRM_CUDA(cudaSetDevice(r->device));
    int32_t d[3];
    octree_mip_dims(*r, mip, d);
    TBRM_CUDA(cudaMemcpyAsync(dst_host, r->octree[mip], (size_t) d[0] * d[1] * d[2] * sizeof(uint16_t), cudaMemcpyDeviceToHost, r->stream));
    TBRM_CUDA(cudaStreamSynchronize(r->stream));
    return TBRM_OK;
}

// material: 1 intensity, 2 octree
static tbrm_status raymarch_material_impl(tbrm_resources* r, int material, const tbrm_camera* cam, const tbrm_world* world, float step_count,
                                          int octree_mip, int row_begin, int row_end, float* out_rgba, int out_is_device, uint64_t* out_steps) {
    if (!r || !r->data || !r->data_ready) return TBRM_ERR_NOT_INITIALIZED;
    if (material == 2 && (!r->tf_ready || !r->octree_valid)) return TBRM_ERR_NOT_INITIALIZED;
    TBRM_REQUIRE(world && out_rgba, "raymarch: null argument");
    TBRM_REQUIRE(camera_valid(cam), "raymarch: invalid camera");
    TBRM_REQUIRE(step_count > 0.0f, "raymarch: step count must be positive");
    TBRM_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= cam->height, "raymarch: bad row range");
    TBRM_REQUIRE(material != 2 || (octree_mip >= 0 && octree_mip < 4), "tbrm_raymarch_octree: the octree has mips 0..3");
    if (row_begin == row_end) {
        if (out_steps) *out_steps = 0;
        return TBRM_OK;
    }
    TBRM_CUDA(cudaSetDevice(r->device));
    host::CameraUniforms cu;
    host::plan_camera(*cam, *world, cu);
    float cc[3], cd[3];
    host::plan_clip(*world, cc, cd);
    unsigned long long* d_steps = out_steps ? r->counters : nullptr;
    if (d_steps) TBRM_CUDA(cudaMemsetAsync(d_steps, 0, sizeof(unsigned long long), r->stream));
    const size_t bytes = (size_t) cam->width * (row_end - row_begin) * 4 * sizeof(float);
    tbrm_status s = with_output(r->device, r->stream, out_rgba, bytes, out_is_device, [&](void* d) {
        return material == 1 ? raymarch_intensity(*r, cu, cc, cd, step_count, row_begin, row_end, (float*) d, d_steps)
                             : raymarch_octree(*r, cu, cc, cd, step_count, octree_mip, row_begin, row_end, (float*) d, d_steps);
    });
    if (s != TBRM_OK) return s;
    if (out_steps) {
        unsigned long long h = 0;
        TBRM_CUDA(cudaMemcpyAsync(&h, d_steps, sizeof(h), cudaMemcpyDeviceToHost, r->stream));
        TBRM_CUDA(cudaStreamSynchronize(r->stream));
        *out_steps = h;
    }
    return TBRM_OK;
}

tbrm_status tbrm_raymarch_intensity(tbrm_resources* r, const tbrm_camera* cam, const tbrm_world* world, float step_count, int row_begin,
                                    int row_end, float* out_rgba, int out_is_device, uint64_t* out_steps) {
    TBRM_RANGE("tbrm::IntensityRaymarch");
    return raymarch_material_impl(r, 1, cam, world, step_count, 0, row_begin, row_end, out_rgba, out_is_device, out_steps);
}

tbrm_status tbrm_raymarch_octree(tbrm_resources* r, const tbrm_camera* cam, const tbrm_world* world, float step_count, int octree_mip,
                                 int row_begin, int row_end, float* out_rgba, int out_is_device, uint64_t* out_steps) {
    TBRM_RANGE("tbrm::OctreeRaymarch");
    return raymarch_material_impl(r, 2, cam, world, step_count, octree_mip, row_begin, row_end, out_rgba, out_is_device, out_steps);
}

// ---- Mandelbulb variants ---------------------------------------------------------------
tbrm_status tbrm_mandelbulb_march_normal(int device, const tbrm_mandelbulb* params, float derivation_distance, const tbrm_camera* cam,
                                         const tbrm_world* world, int row_begin, int row_end, float* out_rgba, int out_is_device,
                                         uint64_t* out_iterations) {
    TBRM_RANGE("tbrm::MandelbulbMarchNormal");
    TBRM_REQUIRE(params && world && out_rgba, "tbrm_mandelbulb_march_normal: null argument");
    TBRM_REQUIRE(camera_valid(cam), "tbrm_mandelbulb_march_normal: invalid camera");
    TBRM_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= cam->height, "tbrm_mandelbulb_march_normal: bad row range");
    TBRM_REQUIRE(params->extent != 0.0f, "tbrm_mandelbulb_march_normal: extent must be non-zero");
    if (row_begin == row_end) {
        if (out_iterations) *out_iterations = 0;
        return TBRM_OK;
    }
    host::CameraUniforms cu;
    host::plan_camera(*cam, *world, cu);
    const size_t bytes = (size_t) cam->width * (row_end - row_begin) * 4 * sizeof(float);
    return mandelbulb_op(device, out_rgba, bytes, out_is_device, out_iterations, [&](void* d, unsigned long long* it) {
        return mandelbulb_march_normal(cudaStreamPerThread, *params, derivation_distance, cu, row_begin, row_end, (float*) d, it);
    });
}

float tbrm_debug_mandelbulb_sdf_p8(const float position[3], float bailout, int iterations, uint32_t* out_iterations) {
    unsigned int it = 0;
    const float d = position ? mandelbulb_sdf_p8_host(position[0], position[1], position[2], bailout, iterations, &it) : 0.0f;
    if (out_iterations) *out_iterations = it;
    return d;
}

tbrm_status tbrm_debug_download_derived(tbrm_resources* res, int which, void* dst, size_t capacity) {
    TBRM_REQUIRE(res && dst, "tbrm_debug_download_derived: null argument");
    if (which == 3) {  // TBRM_CHAIN_TIMERS builds: the chain kernel's section timers of the last pass
        TBRM_REQUIRE(res->dbg != nullptr, "tbrm_debug_download_derived: not a TBRM_CHAIN_TIMERS build");
        TBRM_CUDA(cudaSetDevice(res->device));
        TBRM_CUDA(cudaStreamSynchronize(res->stream));
        TBRM_CUDA(cudaMemcpy(dst, res->dbg, std::min(capacity, (size_t) 4096 * 4 * 8 * sizeof(long long)), cudaMemcpyDeviceToHost));
        return TBRM_OK;
    }
    if (which == 4) {  // geometry of the last TMA-staged sweep launch: tile rows, pixels per thread, tiles this GPU launched, bands
        TBRM_REQUIRE(capacity >= sizeof(res->last_geom), "tbrm_debug_download_derived: destination too small");
        memcpy(dst, res->last_geom, sizeof(res->last_geom));
        return TBRM_OK;
    }
    TBRM_REQUIRE(which >= 0 && which <= 2, "tbrm_debug_download_derived: which must be 0 (brick grid), 1 (yzx replica), 2 (T-brick flags of the last split sweep pass) or 4 (launch geometry)");
    tbrm_resources* r = res;
    if (r->data_fmt != TBRM_FMT_G8 || !r->data_ready) {
        set_last_error("tbrm_debug_download_derived: needs an uploaded R8 data volume");
        return TBRM_ERR_NOT_INITIALIZED;
    }
    TBRM_CUDA(cudaSetDevice(r->device));
    size_t bytes = r->data_voxels();
    const void* src = nullptr;
    if (which == 0) {
        TBRM_CUDA(ensure_bricks(*r));
        bytes = (size_t) ((r->ddims[0] + 7) / 8) * ((r->ddims[1] + 7) / 8) * ((r->ddims[2] + 7) / 8);
        src = r->bricks;
    } else if (which == 1) {
        TBRM_CUDA(build_replica_for_tests(*r));
        src = r->data_yzx;
    } else {
        TBRM_REQUIRE(r->tones != nullptr, "tbrm_debug_download_derived: no split sweep pass has run");
        bytes = std::min(capacity, r->tones_bytes);
        src = r->tones;
    }
    TBRM_REQUIRE(capacity >= bytes, "tbrm_debug_download_derived: destination too small");
    TBRM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, r->stream));
    TBRM_CUDA(cudaStreamSynchronize(r->stream));
    return TBRM_OK;
}

tbrm_status tbrm_mandelbulb_sdf(int device, const int32_t dims[3], const float center[3], float extent, float power, tbrm_format out_fmt,
                                void* dst, int dst_is_device, uint64_t* out_iterations) {
    TBRM_REQUIRE(dims && center && dst, "tbrm_mandelbulb_sdf: null argument");
    TBRM_REQUIRE(dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && dims[2] <= 65535, "tbrm_mandelbulb_sdf: bad dimensions");
    TBRM_REQUIRE(out_fmt == TBRM_FMT_G16 || out_fmt == TBRM_FMT_R32F, "tbrm_mandelbulb_sdf: output is G16 (the reference's texture) or R32F");
    if (out_iterations) *out_iterations = 0;
    if (!(extent > 0.0f)) return TBRM_OK;  // EnqueueRenderCommand_CalculateMandelbulbSDF: Extent <= 0 -> return (FractalShaders.cpp:28-31)
    const size_t bytes = (size_t) dims[0] * dims[1] * dims[2] * (out_fmt == TBRM_FMT_G16 ? 2 : 4);
    return mandelbulb_op(device, dst, bytes, dst_is_device, out_iterations, [&](void* d, unsigned long long* it) {
        return mandelbulb_sdf_bake(cudaStreamPerThread, dims, center, extent, power, out_fmt == TBRM_FMT_G16, d, it);
    });
}

// ---- volume ingest -------------------------------------------------------------------
tbrm_status tbrm_mhd_parse_header(const char* header_text, tbrm_volume_info* out) {
    TBRM_REQUIRE(header_text && out, "tbrm_mhd_parse_header: null argument");
    if (!mhd_parse_header(header_text, *out)) {
        set_last_error("tbrm_mhd_parse_header: DimSize, ElementSpacing | ElementSize, ElementType (MET_*) and ElementDataFile are required");
        return TBRM_ERR_INVALID_ARGUMENT;
    }
    return TBRM_OK;
}

// FVolumeInfo::NormalizeValue ... DenormalizeRange — VolumeInfo.cpp:18-55
float tbrm_volume_info_normalize_value(const tbrm_volume_info* i, float v) {
    return (!i || !i->is_normalized) ? v : ((v - i->min_value) / (i->max_value - i->min_value));
}
float tbrm_volume_info_denormalize_value(const tbrm_volume_info* i, float v) {
    return (!i || !i->is_normalized) ? v : ((v * (i->max_value - i->min_value)) + i->min_value);
}
float tbrm_volume_info_normalize_range(const tbrm_volume_info* i, float v) {
    return (!i || !i->is_normalized) ? v : (v / (i->max_value - i->min_value));
}
float tbrm_volume_info_denormalize_range(const tbrm_volume_info* i, float v) {
    return (!i || !i->is_normalized) ? v : (v * (i->max_value - i->min_value));
}

tbrm_status tbrm_converted_format(const tbrm_volume_info* info, int normalize, int convert_to_float, int* out_texture_format,
                                  int* out_actual_format) {
    TBRM_REQUIRE(info && out_texture_format && out_actual_format, "tbrm_converted_format: null argument");
    TBRM_REQUIRE(voxel_format_bytes(info->original_format) > 0, "tbrm_converted_format: unknown voxel format");
    converted_format(*info, normalize, convert_to_float, *out_texture_format, *out_actual_format);
    return TBRM_OK;
}

tbrm_status tbrm_normalize_volume(int device, int voxel_format, const void* src, int src_is_device, uint64_t count, void* dst,
                                  int dst_is_device, float* out_min, float* out_max) {
    TBRM_RANGE("tbrm::NormalizeVolume");
    TBRM_REQUIRE(src && dst && count > 0, "tbrm_normalize_volume: null argument or empty volume");
    const int ib = voxel_format_bytes(voxel_format);
    TBRM_REQUIRE(ib > 0, "tbrm_normalize_volume: unknown voxel format");
    if (tbrm_device_count() <= 0) {
        set_last_error("tbrm_normalize_volume: no CUDA device visible");
        return TBRM_ERR_NO_DEVICE;
    }
    TBRM_CUDA(cudaSetDevice(device));
    const size_t ob = ib == 1 ? 1 : 2;
    IngestBuffers b;
    tbrm_status s = ingest_stage(b, src, src_is_device, (size_t) count * ib, dst, dst_is_device, (size_t) count * ob);
    if (s != TBRM_OK) return s;
    void* scratch = ingest_scratch(device);
    if (!scratch) {
        set_last_error("tbrm_normalize_volume: cannot allocate the reduction scratch");
        return TBRM_ERR_CUDA;
    }
    float* d_minmax = (float*) ((char*) scratch + ingest_partials_bytes());
    TBRM_CUDA(ingest_normalize(cudaStreamPerThread, voxel_format, b.d_in, (size_t) count, b.d_out, scratch, d_minmax));
    float mm[2] = {0.0f, 0.0f};
    TBRM_CUDA(cudaMemcpyAsync(mm, d_minmax, sizeof(mm), cudaMemcpyDeviceToHost, cudaStreamPerThread));
    if (!dst_is_device) TBRM_CUDA(cudaMemcpyAsync(dst, b.d_out, (size_t) count * ob, cudaMemcpyDeviceToHost, cudaStreamPerThread));
    TBRM_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
    if (out_min) *out_min = mm[0];
    if (out_max) *out_max = mm[1];
    return TBRM_OK;
}

tbrm_status tbrm_convert_volume_to_float(int device, int voxel_format, const void* src, int src_is_device, uint64_t count, float* dst,
                                         int dst_is_device) {
    TBRM_REQUIRE(src && dst && count > 0, "tbrm_convert_volume_to_float: null argument or empty volume");
    const int ib = voxel_format_bytes(voxel_format);
    TBRM_REQUIRE(ib > 0 && voxel_format != TBRM_VOXEL_F32, "tbrm_convert_volume_to_float: integer voxel formats only (ConvertArrayToFloat)");
    if (tbrm_device_count() <= 0) {
        set_last_error("tbrm_convert_volume_to_float: no CUDA device visible");
        return TBRM_ERR_NO_DEVICE;
    }
    TBRM_CUDA(cudaSetDevice(device));
    IngestBuffers b;
    tbrm_status s = ingest_stage(b, src, src_is_device, (size_t) count * ib, dst, dst_is_device, (size_t) count * sizeof(float));
    if (s != TBRM_OK) return s;
    TBRM_CUDA(ingest_to_float(cudaStreamPerThread, voxel_format, b.d_in, (size_t) count, (float*) b.d_out));
    if (!dst_is_device) TBRM_CUDA(cudaMemcpyAsync(dst, b.d_out, (size_t) count * sizeof(float), cudaMemcpyDeviceToHost, cudaStreamPerThread));
    TBRM_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
    return TBRM_OK;
}

// IVolumeLoader::LoadAndConvertData (VolumeLoader.cpp:88-128) + CreateVolumeTextureTransient + InitializeRaymarchResources: the voxels of
// `data_path` described by `info`, converted on the GPU, as the data volume of a new resource set
static tbrm_status load_volume_from_info(int device, const std::string& data_path, int normalize, int convert_to_float, tbrm_format light_fmt,
                                         int half_res, tbrm_volume_info* info, tbrm_resources** out, const char* who) {
    TBRM_REQUIRE(info->dims[0] > 0 && info->dims[1] > 0 && info->dims[2] > 0, std::string(who) + ": the volume has a zero dimension");
    const int of = info->original_format;
    int tex = -1, actual = of;
    converted_format(*info, normalize, convert_to_float, tex, actual);
    if (tex < 0) {
        set_last_error(std::string(who) + ": unnormalised 32-bit integer voxels map to PF_R32_SINT, which the path does not sample");
        return TBRM_ERR_UNSUPPORTED;
    }
    const tbrm_format data_fmt = (tbrm_format) tex;
    std::vector<uint8_t> voxels;
    std::string err;
    if (!load_voxel_file(data_path, *info, voxels, err)) {
        set_last_error(std::string(who) + ": " + err);
        return TBRM_ERR_INVALID_ARGUMENT;
    }
    tbrm_status s = tbrm_create(device, info->dims, data_fmt, light_fmt, half_res, out);
    if (s != TBRM_OK) return s;
    tbrm_resources* r = *out;
    const uint64_t count = (uint64_t) r->data_voxels();
    if (cudaMalloc(&r->data, (size_t) count * r->data_elem()) != cudaSuccess) {
        set_last_error(std::string(who) + ": cudaMalloc(data volume) failed");
        r->data = nullptr;
        tbrm_destroy(r);
        *out = nullptr;
        return TBRM_ERR_CUDA;
    }
    r->data_owned = true;
    if (normalize) {
        s = tbrm_normalize_volume(device, of, voxels.data(), 0, count, r->data, 1, &info->min_value, &info->max_value);
        info->is_normalized = 1;
        if (info->bytes_per_voxel > 1) info->bytes_per_voxel = 2;  // VolumeLoader.cpp:106-110; the float conversion leaves it alone (:116-121)
    } else if (convert_to_float && of != TBRM_VOXEL_F32) {
        s = tbrm_convert_volume_to_float(device, of, voxels.data(), 0, count, (float*) r->data, 1);
    } else {
        s = cudaMemcpy(r->data, voxels.data(), voxels.size(), cudaMemcpyHostToDevice) == cudaSuccess ? TBRM_OK : TBRM_ERR_CUDA;
    }
    if (s != TBRM_OK) {
        tbrm_destroy(r);
        *out = nullptr;
        return s;
    }
    info->actual_format = actual;
    r->data_ready = true;
    return TBRM_OK;
}

// no C++ exception may cross the C ABI (a header that promises 2^60 bytes must come back as a status, not terminate the host process)
extern "C++" {
template <typename F>
static tbrm_status guarded(const char* what, F body) {
    try {
        return body();
    } catch (const std::exception& e) {
        set_last_error(std::string(what) + ": " + e.what());
        return TBRM_ERR_INVALID_ARGUMENT;
    } catch (...) {
        set_last_error(std::string(what) + ": unknown failure");
        return TBRM_ERR_INVALID_ARGUMENT;
    }
}
}  // extern "C++"

static tbrm_status load_mhd_volume_impl(int device, const char* mhd_path, int normalize, int convert_to_float, tbrm_format light_fmt, int half_res,
                                        tbrm_volume_info* info, tbrm_resources** out) {
    TBRM_REQUIRE(mhd_path && info && out, "tbrm_load_mhd_volume: null argument");
    *out = nullptr;
    std::string text;
    {
        FILE* f = std::fopen(mhd_path, "rb");
        if (!f) {
            set_last_error(std::string("tbrm_load_mhd_volume: cannot read ") + mhd_path);
            return TBRM_ERR_INVALID_ARGUMENT;
        }
        char buf[4096];
        size_t n;
        while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
        std::fclose(f);
    }
    tbrm_status s = tbrm_mhd_parse_header(text.c_str(), info);
    if (s != TBRM_OK) return s;
    std::string dir(mhd_path);
    const size_t slash = dir.find_last_of("/\\");
    dir = slash == std::string::npos ? std::string(".") : dir.substr(0, slash);
    // LoadRawDataFileFromInfo: FilePath + "/" + DataFileName
    return load_volume_from_info(device, dir + "/" + info->data_file, normalize, convert_to_float, light_fmt, half_res, info, out, "tbrm_load_mhd_volume");
}

tbrm_status tbrm_load_mhd_volume(int device, const char* mhd_path, int normalize, int convert_to_float, tbrm_format light_fmt, int half_res,
                                 tbrm_volume_info* info, tbrm_resources** out) {
    return guarded("tbrm_load_mhd_volume", [&] { return load_mhd_volume_impl(device, mhd_path, normalize, convert_to_float, light_fmt, half_res, info, out); });
}

static tbrm_status load_raw_volume_impl(int device, const char* raw_path, const int32_t dims[3], int voxel_format, int64_t compressed_bytes, int normalize,
                                        int convert_to_float, tbrm_format light_fmt, int half_res, tbrm_volume_info* info, tbrm_resources** out) {
    TBRM_REQUIRE(raw_path && dims && info && out, "tbrm_load_raw_volume: null argument");
    *out = nullptr;
    TBRM_REQUIRE(voxel_format_bytes(voxel_format) > 0, "tbrm_load_raw_volume: unknown voxel format");
    TBRM_REQUIRE(dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && compressed_bytes >= 0, "tbrm_load_raw_volume: bad dimensions / compressed size");
    std::memset(info, 0, sizeof(*info));
    info->parse_ok = 1;
    for (int k = 0; k < 3; ++k) info->dims[k] = dims[k], info->spacing[k] = 1.0, info->world_dims[k] = (double) dims[k];
    info->original_format = info->actual_format = voxel_format;
    info->bytes_per_voxel = voxel_format_bytes(voxel_format);
    info->is_signed = (voxel_format == TBRM_VOXEL_I8 || voxel_format == TBRM_VOXEL_I16 || voxel_format == TBRM_VOXEL_I32 || voxel_format == TBRM_VOXEL_F32);
    info->min_value = -1000.0f, info->max_value = 3000.0f;
    info->is_compressed = compressed_bytes > 0, info->compressed_bytes = compressed_bytes > 0 ? compressed_bytes : 0;
    std::snprintf(info->data_file, sizeof(info->data_file), "%s", raw_path);
    return load_volume_from_info(device, raw_path, normalize, convert_to_float, light_fmt, half_res, info, out, "tbrm_load_raw_volume");
}

tbrm_status tbrm_load_raw_volume(int device, const char* raw_path, const int32_t dims[3], int voxel_format, int64_t compressed_bytes, int normalize,
                                 int convert_to_float, tbrm_format light_fmt, int half_res, tbrm_volume_info* info, tbrm_resources** out) {
    return guarded("tbrm_load_raw_volume",
                   [&] { return load_raw_volume_impl(device, raw_path, dims, voxel_format, compressed_bytes, normalize, convert_to_float, light_fmt, half_res, info, out); });
}

// ---- queue control ----------------------------------------------------------------------------------------
tbrm_status tbrm_flush(tbrm_resources* r) {
    TBRM_REQUIRE(r, "tbrm_flush: null argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(cudaStreamSynchronize(r->stream));
    return TBRM_OK;
}

void* tbrm_stream(tbrm_resources* r) { return r ? (void*) r->stream : nullptr; }

tbrm_status tbrm_set_stream(tbrm_resources* r, void* cuda_stream) {
    TBRM_REQUIRE(r, "tbrm_set_stream: null argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(cudaStreamSynchronize(r->stream));
    if (r->stream_owned && r->stream) cudaStreamDestroy(r->stream);
    r->stream = (cudaStream_t) cuda_stream;
    r->stream_owned = false;
    return TBRM_OK;
}

tbrm_status tbrm_timer_begin(tbrm_resources* r) {
    TBRM_REQUIRE(r, "tbrm_timer_begin: null argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(cudaEventRecord(r->ev_begin, r->stream));
    return TBRM_OK;
}

tbrm_status tbrm_timer_end(tbrm_resources* r, float* out_ms) {
    TBRM_REQUIRE(r && out_ms, "tbrm_timer_end: null argument");
    TBRM_CUDA(cudaSetDevice(r->device));
    TBRM_CUDA(cudaEventRecord(r->ev_end, r->stream));
    TBRM_CUDA(cudaEventSynchronize(r->ev_end));
    TBRM_CUDA(cudaEventElapsedTime(out_ms, r->ev_begin, r->ev_end));
    return TBRM_OK;
}

tbrm_status tbrm_synth_volume_u8(int device, int kind, const int32_t dims[3], uint32_t seed, void* dst, int dst_is_device) {
    TBRM_REQUIRE(dims && dst, "tbrm_synth_volume_u8: null argument");
    TBRM_REQUIRE(dims[0] > 0 && dims[1] > 0 && dims[2] > 0, "tbrm_synth_volume_u8: empty volume");
    TBRM_REQUIRE(kind == TBRM_SYNTH_SPHERE || kind == TBRM_SYNTH_PERLIN_CT, "tbrm_synth_volume_u8: unknown kind");
    if (tbrm_device_count() <= 0) {
        set_last_error("tbrm_synth_volume_u8: no CUDA device visible");
        return TBRM_ERR_NO_DEVICE;
    }
    TBRM_CUDA(cudaSetDevice(device));
    const size_t bytes = (size_t) dims[0] * dims[1] * dims[2];
    tbrm_status s = with_output(device, cudaStreamPerThread, dst, bytes, dst_is_device,
                                [&](void* d) { return synth_volume_u8(cudaStreamPerThread, kind, dims, seed, (uint8_t*) d); });
    if (s == TBRM_OK && dst_is_device) TBRM_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
    return s;
}

}  // extern "C"
