/*
 * tbrm.h — C ABI of the B200-native raymarch / illumination-sweep hot path.
 *
 * This is the drop-in boundary for the two data-parallel hot paths of
 * tommybazar/TBRaymarcherPlugin (reference paths are relative to the plugin root):
 *
 *   URaymarchUtils::AddDirLightToSingleVolume      Source/Raymarcher/Public/Util/RaymarchUtils.h:33-35
 *   URaymarchUtils::ChangeDirLightInSingleVolume   Source/Raymarcher/Public/Util/RaymarchUtils.h:39-41
 *   URaymarchUtils::ClearResourceLightVolumes      Source/Raymarcher/Public/Util/RaymarchUtils.h:49
 *   URaymarchUtils::MakeDefaultTFTexture           Source/Raymarcher/Public/Util/RaymarchUtils.h:60
 *   URaymarchUtils::ColorCurveToTexture            Source/Raymarcher/Public/Util/RaymarchUtils.h:64
 *   PerformRaymarchCubeSetup                       Source/Raymarcher/Shaders/Private/RaymarchMaterialCommon.usf:23-69
 *   PerformWindowedLitRaymarch                     Source/Raymarcher/Shaders/Private/WindowedRaymarchMaterials.usf:36-96
 *   PerformMandelbulbRaymarchReturnDistance        Source/FractalMarcher/Shaders/Private/SDFMarcher.usf:61-112
 * and, next to the path:
 *   URaymarchUtils::GenerateOctree                 Source/Raymarcher/Public/Util/RaymarchUtils.h:45
 *   PerformWindowedIntensityRaymarch / ...Octree   Source/Raymarcher/Shaders/Private/WindowedRaymarchMaterials.usf:99-242
 *   PerformMandelbulbRaymarchReturnNormal          Source/FractalMarcher/Shaders/Private/SDFMarcher.usf:117-188
 *   EnqueueRenderCommand_CalculateMandelbulbSDF    Source/FractalMarcher/Public/Rendering/FractalShaders.h
 *   UMHDLoader / IVolumeLoader / UVolumeTextureToolkit::NormalizeArrayByFormat   Source/VolumeTextureToolkit
 *
 * Conventions
 *   - plain C, no torch / CUDA types in any signature; device pointers are passed as void*.
 *   - every op is enqueued on the resource set's CUDA stream and returns immediately
 *     (the reference enqueues onto UE's render thread, RaymarchUtils.cpp:63-66,87-91);
 *     tbrm_flush() is FlushRenderingCommands().
 *   - functions return a tbrm_status; the reference's `bool& LightAdded` out-parameter and its
 *     silent no-op rules (zero light direction, LightingShaders.cpp:41-46,173-179) are preserved.
 *   - volumes are x-fastest: idx = x + y*X + z*X*Y (TextureUtilities.cpp:43-78).
 */
#ifndef TBRM_H_
#define TBRM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TBRM_ABI_VERSION 1

typedef enum tbrm_status {
    TBRM_OK = 0,
    TBRM_ERR_INVALID_ARGUMENT = 1,
    TBRM_ERR_NOT_INITIALIZED = 2, /* a resource the reference null-checks is missing (RaymarchUtils.cpp:39-45) */
    TBRM_ERR_CUDA = 3,
    TBRM_ERR_UNSUPPORTED = 4,
    TBRM_ERR_NO_DEVICE = 5
} tbrm_status;

/* EPixelFormat subset the path uses (RaymarchVolume.cpp:857-861, VolumeInfo.h) */
typedef enum tbrm_format {
    TBRM_FMT_G8 = 0,   /* UNORM8  */
    TBRM_FMT_G16 = 1,  /* UNORM16 (data volume only) */
    TBRM_FMT_R32F = 2  /* float   */
} tbrm_format;

/* FDirLightParameters — Source/Raymarcher/Public/Rendering/RaymarchTypes.h:20-41 */
typedef struct tbrm_dir_light {
    double direction[3]; /* FVector LightDirection (world space, need not be unit) */
    float intensity;     /* float LightIntensity */
} tbrm_dir_light;

/* FClippingPlaneParameters — RaymarchTypes.h:45-71 */
typedef struct tbrm_clip_plane {
    double center[3];
    double direction[3]; /* "the direction from the center that is NOT clipped away" */
} tbrm_clip_plane;

/* FRaymarchWorldParameters — RaymarchTypes.h:136-153. VolumeTransform is an FTransform
 * {Translation, Rotation (quaternion x,y,z,w), Scale3D}. */
typedef struct tbrm_world {
    double translation[3];
    double rotation[4];
    double scale[3];
    tbrm_clip_plane clip;
} tbrm_world;

/* FWindowingParameters — Source/VolumeTextureToolkit/Public/VolumeAsset/VolumeInfo.h:32-53 */
typedef struct tbrm_windowing {
    float center;
    float width;
    int32_t low_cutoff;
    int32_t high_cutoff;
} tbrm_windowing;

/* Stand-in for the UE view the material reads (ResolvedView.WorldCameraOrigin, CameraVector,
 * ViewToTranslatedWorld, SvPosition, View.StateFrameIndexMod8 — RaymarchMaterialCommon.usf:26-48,75-76):
 * an explicit pinhole camera. Pixel (ix,iy) looks through the centre of the pixel; iy grows downwards. */
typedef struct tbrm_camera {
    double eye[3];     /* world space */
    double look_at[3]; /* world space */
    double up[3];      /* world space, need not be orthogonal to the view direction */
    double hfov_deg;   /* horizontal field of view */
    int32_t width;
    int32_t height;
    float scene_depth;  /* CalcSceneDepth stand-in (constant over the image); <= 0 selects 1e8 */
    int32_t frame_index; /* View.StateFrameIndexMod8 is frame_index % 8 */
    int32_t jitter;      /* 1 = JitterEntryPos with Rand3DPCG16 (reference), 0 = off */
} tbrm_camera;

/* Parameters of PerformMandelbulbRaymarchReturnDistance — SDFMarcher.usf:61-72 */
typedef struct tbrm_mandelbulb {
    float center[3];
    float extent;
    float power;
    float max_steps;
    float max_iterations;
    float bailout;
    float high_precision_eps;
    float low_precision_eps;
} tbrm_mandelbulb;

/* Engine-semantics switches. Zero-initialised = reference-faithful. */
typedef struct tbrm_options {
    int32_t border_exact;     /* 0: 8-bit FColor round trip of sampler border colours (Q1,Q2); 1: exact values */
    int32_t data_addr_wrap;   /* raymarch data sampler address mode: 0 clamp (default), 1 wrap (Q5) */
    int32_t sweep_impl;       /* 0: auto (gpu_sync ? fused : per-slice), 1: per-slice launches (reference schedule),
                                 2: fused persistent sweep (TMA-staged when eligible), 3: generic fused sweep only */
    int32_t reserved[5];      /* debug / test hooks (INTEGRATION.md has the table): [0] bit 0 disables the sweep's exact empty-space skip,
                                 bits 4-5 pixels per thread, bit 6 second kernel generation, bits 8-9 tile rows (2: 7, 3: 8); [1] = 1 forces
                                 the generic raymarch kernel (2..5: other forms); [2] > 0 caps the tile rows of one sweep launch (forces
                                 banded passes) */
} tbrm_options;

/* Per-op counters written by the *_stats calls (for the metric definitions of SURVEY.md §8d). */
typedef struct tbrm_sweep_stats {
    int32_t passes;          /* axis passes executed (0, 1 or 2; 2..4 for a Change that fell back to Remove+Add) */
    int32_t fell_back;       /* Change: 1 if the major axes differed and Remove+Add ran (LightingShaders.cpp:192-198) */
    int64_t voxels;          /* light-volume voxels x passes */
    int32_t kernel_launches; /* CUDA kernels launched by this op */
    int32_t faces[4];        /* FCubeFace of each pass (0:+X 1:-X 2:+Y 3:-Y 4:+Z 5:-Z), -1 if unused */
    int32_t impl[4];         /* kernel family of each pass: 1 per-slice launches, 2 fused (generic), 3 fused (TMA-staged),
                                4 joined per-slice launches (tbrm_add_dir_lights_joined) */
} tbrm_sweep_stats;

/* One GPU's share of a volume that is sharded over the GPUs of a box as Z-slabs. */
typedef struct tbrm_slab {
    int32_t rank, nranks;    /* position in the chain of slabs; nranks == 1: not sharded */
    int32_t z_begin, z_end;  /* light-volume slices [z_begin, z_end) this GPU owns (tbrm_slab_partition) */
} tbrm_slab;

/* Opaque FBasicRaymarchRenderingResources (RaymarchTypes.h:87-129): data volume, TF texture, light volume,
 * windowing, the 3x4 propagation buffers, and the CUDA stream that plays the render-thread queue. */
typedef struct tbrm_resources tbrm_resources;

/* ---- library ---------------------------------------------------------------------------------------- */
int tbrm_abi_version(void);
const char* tbrm_status_string(int status);
const char* tbrm_last_error(void); /* thread-local message for the last non-OK status */
int tbrm_device_count(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t tbrm_kernel_launch_count(void);

/* ---- resources (ARaymarchVolume::InitializeRaymarchResources, RaymarchVolume.cpp:821-920) ------------ */
/* data_dims: data volume (X,Y,Z). light volume dims = data dims, or ceil(dims/2) if half_res (:850-855).
 * light_fmt: TBRM_FMT_G8 (reference default) or TBRM_FMT_R32F (bLightVolume32Bit, :857-861). */
tbrm_status tbrm_create(int device, const int32_t data_dims[3], tbrm_format data_fmt, tbrm_format light_fmt,
                        int half_res, tbrm_resources** out);
tbrm_status tbrm_destroy(tbrm_resources* res); /* FreeRaymarchResources, RaymarchVolume.cpp:922-949 */
tbrm_status tbrm_set_options(tbrm_resources* res, const tbrm_options* opts);

/* Data volume upload: host or device pointer, x-fastest, data_fmt texels (UVolumeTexture contents). */
tbrm_status tbrm_upload_volume(tbrm_resources* res, const void* src, int src_is_device);
/* Use caller-owned device memory as the data volume without copying (must outlive res). */
tbrm_status tbrm_bind_volume_device(tbrm_resources* res, const void* dptr);

/* Streaming (double-buffered) upload for time-varying volumes: copies the NEXT data volume (host memory, pinned for a true
 * overlap) into a back buffer on a dedicated upload stream while the render queue keeps working on the current one, the way
 * the engine streams texture updates; tbrm_present_volume makes the render queue wait for that copy and swaps the buffers.
 * Both return immediately. src_host must stay valid until the copy has run (tbrm_present_volume + tbrm_flush, or the next
 * tbrm_upload_volume_async). */
tbrm_status tbrm_upload_volume_async(tbrm_resources* res, const void* src_host);
tbrm_status tbrm_present_volume(tbrm_resources* res);

/* Transfer function: RGBA float32, `width` x `height` texels, rounded to fp16 like PF_FloatRGBA
 * (ColorCurveToTexture, RaymarchUtils.cpp:143-174). width must be 256. */
tbrm_status tbrm_set_transfer_function(tbrm_resources* res, const float* rgba, int width, int height);
tbrm_status tbrm_make_default_tf(tbrm_resources* res); /* MakeDefaultTFTexture, RaymarchUtils.cpp:113-141 */
tbrm_status tbrm_set_windowing(tbrm_resources* res, const tbrm_windowing* w);

/* ---- illumination sweep ------------------------------------------------------------------------------- */
/* ClearResourceLightVolumes (RaymarchUtils.cpp:104-111) */
tbrm_status tbrm_clear_light_volume(tbrm_resources* res, float clear_value);
/* AddDirLightToSingleVolume (RaymarchUtils.cpp:35-68). gpu_sync selects the single-launch fused sweep
 * (the intent of the reference's disabled GPUSync shader); it produces the same numbers. */
tbrm_status tbrm_add_dir_light(tbrm_resources* res, const tbrm_dir_light* light, int added, const tbrm_world* world,
                               int* light_added, int gpu_sync);
/* ChangeDirLightInSingleVolume (RaymarchUtils.cpp:70-92) */
tbrm_status tbrm_change_dir_light(tbrm_resources* res, const tbrm_dir_light* old_light, const tbrm_dir_light* new_light,
                                  const tbrm_world* world, int* light_added, int gpu_sync);
/* Same two ops, also reporting what ran. */
tbrm_status tbrm_add_dir_light_stats(tbrm_resources* res, const tbrm_dir_light* light, int added,
                                     const tbrm_world* world, int* light_added, int gpu_sync, tbrm_sweep_stats* stats);
tbrm_status tbrm_change_dir_light_stats(tbrm_resources* res, const tbrm_dir_light* old_light,
                                        const tbrm_dir_light* new_light, const tbrm_world* world, int* light_added,
                                        int gpu_sync, tbrm_sweep_stats* stats);

/* Same-axis light joining (the reference's Readme.md:165-166, 186-187 names it as the optimisation of the paper it
 * does not implement): AddDirLight for n_lights lights at once, the passes of all lights that propagate from the same cube face joined into
 * ONE sweep of the reference's per-slice schedule (one launch per slice for the whole group instead of one per light; at most 8 members per
 * sweep). Per voxel the members are evaluated in light order, each with its own propagation buffers and the arithmetic of
 * AddDirLightShader.usf; a single light gives the bits of tbrm_add_dir_light(gpu_sync = 0), several lights differ from consecutive calls
 * only in the order in which a voxel's contributions are summed (<= a few ulp). *lights_added: lights with a non-zero direction, 0 when a
 * resource is missing. stats: passes = sweeps run, impl = 4. Not available on a slab-sharded volume (TBRM_ERR_UNSUPPORTED). */
tbrm_status tbrm_add_dir_lights_joined(tbrm_resources* res, const tbrm_dir_light* lights, int n_lights, int added, const tbrm_world* world,
                                       int* lights_added, tbrm_sweep_stats* stats);

/* Host parameter math of one light (LightingShaderUtils.cpp:29-265, LightingShaders.cpp:100-130): what the
 * render-thread functions compute on the CPU before dispatching. Pure host function, needs no GPU. */
typedef struct tbrm_pass_plan {
    int32_t face, axis, dirn; /* FCubeFace, face/2, GetAxisDirection */
    int32_t td[3];            /* GetTransposedDimensions */
    int32_t start, stop;      /* GetLoopStartStopIndexes */
    float weight;             /* FMajorAxes::FaceWeight[i].second after the 0.99 rule */
    float light_alpha;        /* GetLightAlpha */
    float border;             /* read-buffer sampler border colour (GetBorderColorIntSingle round trip) */
    float uv_offset[2];       /* GetUVOffset */
    float uvw_offset[3];      /* GetStepSizeAndUVWOffset + longest-voxel-side renormalisation */
    float step_size;
} tbrm_pass_plan;
typedef struct tbrm_light_plan {
    int32_t zero_direction; /* 1: the reference returns without doing anything */
    int32_t add_passes;     /* passes AddDirLight executes (0..2); ChangeDirLight always runs both */
    tbrm_pass_plan pass[2];
    float clip_center[3], clip_dir[3]; /* GetLocalClippingParameters */
    float data_border;                 /* data sampler border colour (LightingShaders.h:82-89) */
    double local_dir[3];               /* normalised local light direction */
} tbrm_light_plan;
tbrm_status tbrm_plan_dir_light(const int32_t light_dims[3], const tbrm_windowing* win, const tbrm_options* opts,
                                const tbrm_dir_light* light, const tbrm_world* world, tbrm_light_plan* out);

/* Light volume access (tests, multi-GPU plumbing). Texels are light_fmt. */
tbrm_status tbrm_light_volume_dims(const tbrm_resources* res, int32_t dims[3]);
tbrm_status tbrm_download_light_volume(tbrm_resources* res, void* dst_host);
tbrm_status tbrm_upload_light_volume(tbrm_resources* res, const void* src_host);
void* tbrm_light_volume_device_ptr(tbrm_resources* res);
void* tbrm_data_volume_device_ptr(tbrm_resources* res);

/* Use caller-owned device memory as the light volume (light_fmt texels, must outlive res): lets the host layer run
 * collectives (NCCL all-gather of the slabs) directly on it. The previous light volume is freed. */
tbrm_status tbrm_bind_light_volume_device(tbrm_resources* res, void* dptr);

/* ---- Z-slab sharding of ONE volume over the GPUs of a box ------------------------------- */
/* Every rank holds the whole data volume (1 B/voxel, replicated) and a full-size light volume of which it owns the slices
 * [z_begin, z_end): after tbrm_slab_configure the sweep ops (clear / add / change) touch only the owned slab, and the
 * propagated light crosses slab boundaries inside the sweep kernel through the neighbours' exchange arenas (NVLink peer
 * stores for sweeps along X / Y, a plane hand-off for sweeps along Z). The result is bit-identical to the unsharded sweep.
 * Gathering the slabs (an all-gather of the light volume, in place) is the caller's collective. The slabs are slices of the
 * LIGHT volume (tbrm_slab_partition(light_dims[2], ...): half of the data slices for a half-resolution light volume).
 * Requires R8 data, an R32F or G8 light volume (full or half resolution), X % 16 == 0, Y % 16 == 0 and Z % 8 == 0 of the data
 * volume, light Z % 8 == 0, and for a G8 light volume its own X % 16 == 0 and Y % 16 == 0; other configurations return
 * TBRM_ERR_UNSUPPORTED. */
void tbrm_slab_partition(int32_t z_slices, int32_t nranks, int32_t rank, int32_t* z_begin, int32_t* z_end);
tbrm_status tbrm_slab_configure(tbrm_resources* res, const tbrm_slab* slab);
/* The exchange arena of this rank: device pointer + size, and its CUDA IPC handle (64 bytes) for the neighbour processes. */
tbrm_status tbrm_slab_arena(tbrm_resources* res, void** dptr, size_t* bytes);
tbrm_status tbrm_slab_ipc_handle(tbrm_resources* res, void* handle64);
/* Connect the neighbour that owns the slab below (side = -1) or above (side = +1): from its IPC handle (another process),
 * or from a device pointer valid in this process (same-process peers, tests). */
tbrm_status tbrm_slab_open_peer(tbrm_resources* res, int side, const void* handle64);
tbrm_status tbrm_slab_set_peer(tbrm_resources* res, int side, void* peer_arena_dptr);
/* Clears the exchange arena and restarts the tag sequence. Call on every rank, between two barriers, with no sweep in
 * flight: after creating the peers' connections is not required, after TBRM_ERR_UNSUPPORTED "sequence used up" it is. */
/* Push-gather (no reference counterpart: the reference is single-GPU). With the light volumes of all ranks mapped into each other
 * (tbrm_slab_light_ipc_handle -> tbrm_slab_open_peer_light across processes, tbrm_slab_set_peer_light inside one) and
 * tbrm_slab_push_light(res, 1), the LAST axis pass of every following tbrm_add_dir_light call stores each finished light brick into
 * every other rank's volume as well (TMA stores over NVLink inside the sweep kernel): after the call's kernels have completed on all
 * ranks, every rank holds the whole light volume and no all-gather of the slabs is needed. The light volume must be the library's
 * own allocation (not tbrm_bind_light_volume_device). Enable it for the last light of a reset only: earlier pushes are wasted traffic. */
tbrm_status tbrm_slab_light_ipc_handle(tbrm_resources* res, void* handle64);
tbrm_status tbrm_slab_open_peer_light(tbrm_resources* res, int peer_rank, const void* handle64);
tbrm_status tbrm_slab_set_peer_light(tbrm_resources* res, int peer_rank, void* peer_light_dptr);
tbrm_status tbrm_slab_push_light(tbrm_resources* res, int enable);
tbrm_status tbrm_slab_reset_comm(tbrm_resources* res);
/* TBRM_ERR_CUDA if a slab exchange timed out since the last call (a neighbour died or ran different passes); synchronises. */
tbrm_status tbrm_slab_check(tbrm_resources* res);
tbrm_status tbrm_slab_set_timeout_ms(tbrm_resources* res, int timeout_ms);
/* One axis pass of AddDirLightToSingleVolume (pass = 0 or 1), and the order in which slabs must run it when they cannot
 * run concurrently (several slabs on one GPU): +1 lower slabs first, -1 higher slabs first, 0 any order / no such pass.
 * Production code calls tbrm_add_dir_light on every rank; these two exist for single-GPU tests of the exchange. */
tbrm_status tbrm_add_dir_light_pass(tbrm_resources* res, const tbrm_dir_light* light, int added, const tbrm_world* world,
                                    int pass, int gpu_sync, tbrm_sweep_stats* stats);
tbrm_status tbrm_slab_pass_order(tbrm_resources* res, const tbrm_dir_light* light, const tbrm_world* world, int pass, int* order);

/* ---- raymarch ----------------------------------------------------------------------------------------- */
/* PerformRaymarchCubeSetup for every pixel: out_entry_thickness[4*(iy*W+ix)] = (entry UVW, thickness). */
tbrm_status tbrm_raymarch_cube_setup(tbrm_resources* res, const tbrm_camera* cam, const tbrm_world* world,
                                     float* out_entry_thickness, int out_is_device);
/* PerformRaymarchCubeSetup + PerformWindowedLitRaymarch for pixel rows [row_begin,row_end) of the image.
 * out_rgba receives premultiplied RGBA float32 for those rows only ((row_end-row_begin)*W*4 floats).
 * out_steps (optional) receives the total number of executed march-loop iterations (incl. clipped ones and
 * the final partial step) — the numerator of Mray-steps/s. */
tbrm_status tbrm_raymarch_lit(tbrm_resources* res, const tbrm_camera* cam, const tbrm_world* world, float step_count,
                              int row_begin, int row_end, float* out_rgba, int out_is_device, uint64_t* out_steps);

/* Whole-frame lit raymarch into a device frame owned by the resource set, followed by its copy to out_host (pinned) on a
 * dedicated download stream: returns immediately, the next ops of the render queue overlap the copy. tbrm_download_wait blocks
 * until every such copy has finished (two frames may be in flight). */
tbrm_status tbrm_raymarch_lit_to_host_async(tbrm_resources* res, const tbrm_camera* cam, const tbrm_world* world, float step_count,
                                            float* out_host);
tbrm_status tbrm_download_wait(tbrm_resources* res);

/* The same for the image rows one GPU renders when a frame is dealt to `block_stride` GPUs in interleaved blocks of
 * `block_rows` rows (a multiple of 8): blocks first_block, first_block + block_stride, ... The rows land compacted in
 * out_rgba (tbrm_raymarch_interleaved_rows(...) rows of W float4), in image order. Interleaving balances early ray
 * termination and ray length between the GPUs ("per-GPU tile compositing"). */
tbrm_status tbrm_raymarch_lit_interleaved(tbrm_resources* res, const tbrm_camera* cam, const tbrm_world* world, float step_count,
                                          int block_rows, int first_block, int block_stride, float* out_rgba, int out_is_device,
                                          uint64_t* out_steps);
int tbrm_raymarch_interleaved_rows(int height, int block_rows, int first_block, int block_stride);

/* PerformMandelbulbRaymarchReturnDistance over the image rows [row_begin,row_end): out[2*pixel] = (x, y).
 * out_iterations (optional): total Mandelbulb_SDF inner-loop iterations executed. Needs no resources. */
tbrm_status tbrm_mandelbulb_march(int device, const tbrm_mandelbulb* params, const tbrm_camera* cam,
                                  const tbrm_world* world, int row_begin, int row_end, float* out_xy,
                                  int out_is_device, uint64_t* out_iterations);

/* ---- the other materials and the octree -------------------------------------------- */
/* URaymarchUtils::GenerateOctree (RaymarchUtils.cpp:94-102 -> GenerateOctreeForVolume_RenderThread, OctreeShaders.cpp:28-54): fills the
 * octree volume of the resource set — sides = the data volume's rounded up to powers of two, 4 mips, UNORM16 (RaymarchVolume.cpp:873-877);
 * mip 0 = data value (0 outside the data volume), mip m = max over 2x2x2 texels of mip m-1. Uploading a new data volume invalidates it. */
tbrm_status tbrm_generate_octree(tbrm_resources* res);
tbrm_status tbrm_octree_mip_dims(const tbrm_resources* res, int mip, int32_t dims[3]);
tbrm_status tbrm_download_octree_mip(tbrm_resources* res, int mip, void* dst_host); /* uint16 texels, x fastest */
/* PerformWindowedIntensityRaymarch (WindowedRaymarchMaterials.usf:187-242): the first unclipped sample of every ray, windowed to grey,
 * alpha 1; (0,0,0,0) where nothing is hit. Same camera / row / output conventions as tbrm_raymarch_lit. */
tbrm_status tbrm_raymarch_intensity(tbrm_resources* res, const tbrm_camera* cam, const tbrm_world* world, float step_count, int row_begin,
                                    int row_end, float* out_rgba, int out_is_device, uint64_t* out_steps);
/* PerformWindowedRaymarchOctree (WindowedRaymarchMaterials.usf:99-183): the lit march's loop with a point load from octree mip
 * `octree_mip` (0..3, ARaymarchVolume::OctreeVolumeMip) instead of the trilinear data sample, no light volume. Needs
 * tbrm_generate_octree (TBRM_ERR_NOT_INITIALIZED otherwise). */
tbrm_status tbrm_raymarch_octree(tbrm_resources* res, const tbrm_camera* cam, const tbrm_world* world, float step_count, int octree_mip,
                                 int row_begin, int row_end, float* out_rgba, int out_is_device, uint64_t* out_steps);

/* ---- Mandelbulb variants -------------------------------------------------------------- */
/* PerformMandelbulbRaymarchReturnNormal (SDFMarcher.usf:117-188): out[4*pixel] = (normal, 1) on a hit — the normalised vector of three
 * SDF evaluations offset backwards by derivation_distance / extent per axis —, (0,0,0,1) for a low-precision hit, (0,0,0,0) for a miss. */
tbrm_status tbrm_mandelbulb_march_normal(int device, const tbrm_mandelbulb* params, float derivation_distance, const tbrm_camera* cam,
                                         const tbrm_world* world, int row_begin, int row_end, float* out_rgba, int out_is_device,
                                         uint64_t* out_iterations);
/* CalculateMandelbulbSDF (CalculateMandelbulbSDF.usf:24-65, FractalShaders.cpp:26-70): the distance estimate (50 iterations, bailout =
 * extent) / extent for every voxel of a dims[0] x dims[1] x dims[2] volume whose voxel (x,y,z) sits at center + ((x,y,z)/dims - 0.5) *
 * extent. out_fmt: TBRM_FMT_G16 (UNORM16, the reference's PF_G16 volume texture) or TBRM_FMT_R32F. extent <= 0: nothing happens. */
tbrm_status tbrm_mandelbulb_sdf(int device, const int32_t dims[3], const float center[3], float extent, float power, tbrm_format out_fmt,
                                void* dst, int dst_is_device, uint64_t* out_iterations);

/* Test hook, host only: the kernels' transcendental-free Power == 8 iteration (csrc/mandelbulb.cu, one __host__ __device__ function) evaluated
 * on the HOST for one position — the value Mandelbulb_SDF returns (not divided by anything) and the iterations it ran. */
float tbrm_debug_mandelbulb_sdf_p8(const float position[3], float bailout, int iterations, uint32_t* out_iterations);
/* Test hook: builds (if it is stale) and downloads a structure the kernels derive from an R8 data volume. which = 0: the lit march's brick
 * grid, one byte per 8^3 brick = the largest data byte over voxels [8b, 8b + 8] on every axis, ceil(dims / 8) bricks per axis, x fastest;
 * which = 1: the (y,z,x)-ordered replica the sweeps along X read, dst[y + Y * (z + Z * x)] = data[x + X * (y + Y * z)];
 * which = 4: four int32 describing the last TMA-staged sweep launch — tile rows (7 or 8), pixels per thread, tiles this GPU launched, bands
 * (2 and 3 belong to the second kernel generation: T-brick flags, section timers of diagnostic builds). */
tbrm_status tbrm_debug_download_derived(tbrm_resources* res, int which, void* dst, size_t capacity);

/* ---- volume ingest ------------------------------------------------------------------ */
/* EVolumeVoxelFormat — Source/VolumeTextureToolkit/Public/VolumeAsset/VolumeInfo.h:12-27 */
typedef enum tbrm_voxel_format {
    TBRM_VOXEL_U8 = 0,
    TBRM_VOXEL_I8 = 1,
    TBRM_VOXEL_U16 = 2,
    TBRM_VOXEL_I16 = 3,
    TBRM_VOXEL_U32 = 4,
    TBRM_VOXEL_I32 = 5,
    TBRM_VOXEL_F32 = 6
} tbrm_voxel_format;
/* FVolumeInfo — VolumeInfo.h:56-141 */
typedef struct tbrm_volume_info {
    int32_t parse_ok;          /* bParseWasSuccessful */
    int32_t dims[3];           /* Dimensions */
    double spacing[3];         /* Spacing (mm) */
    double world_dims[3];      /* WorldDimensions = Spacing * Dimensions */
    int32_t original_format;   /* tbrm_voxel_format of the file */
    int32_t actual_format;     /* after normalisation / float conversion (ConvertData, VolumeLoader.cpp:97-128) */
    int32_t bytes_per_voxel;   /* of the file; 2 once a multi-byte volume has been normalised (ConvertData, VolumeLoader.cpp:106-110) */
    int32_t is_signed;
    int32_t is_normalized;     /* bIsNormalized */
    float min_value, max_value; /* MinValue / MaxValue of the original data (defaults -1000 / 3000) */
    int32_t is_compressed;     /* a CompressedDataSize tag was present: the data file is zlib-compressed */
    int64_t compressed_bytes;
    char data_file[512];       /* DataFileName (relative to the header) */
} tbrm_volume_info;
/* UMHDLoader::ParseVolumeInfoFromHeader (MHDLoader.cpp:18-181) on the header's text. TBRM_ERR_INVALID_ARGUMENT + parse_ok = 0 when a
 * required key (DimSize, ElementSpacing | ElementSize, ElementType, ElementDataFile) is missing or the element type is unknown. Host only. */
tbrm_status tbrm_mhd_parse_header(const char* header_text, tbrm_volume_info* out);
/* FVolumeInfo::NormalizeValue / DenormalizeValue / NormalizeRange / DenormalizeRange (VolumeInfo.cpp:18-55): window centre / width
 * between the original value range and the normalised [0,1] texture. Host only. */
float tbrm_volume_info_normalize_value(const tbrm_volume_info* info, float value);
float tbrm_volume_info_denormalize_value(const tbrm_volume_info* info, float value);
float tbrm_volume_info_normalize_range(const tbrm_volume_info* info, float range);
float tbrm_volume_info_denormalize_range(const tbrm_volume_info* info, float range);
/* IVolumeLoader::ConvertData's decision (VolumeLoader.cpp:97-128) + FVolumeInfo::VoxelFormatToPixelFormat (VolumeInfo.cpp:99-122): which voxel
 * format the loaded volume ends up in (*out_actual_format, a tbrm_voxel_format) and which texture format that is (*out_texture_format, a
 * tbrm_format; -1 for PF_R32_SINT — unnormalised 32-bit integers —, which the path does not sample). Host only. */
tbrm_status tbrm_converted_format(const tbrm_volume_info* info, int normalize, int convert_to_float, int* out_texture_format,
                                  int* out_actual_format);
/* UVolumeTextureToolkit::NormalizeArrayByFormat (TextureUtilities.cpp:304-327) on the GPU: min / max of `count` voxels of
 * voxel_format, then every voxel mapped to the full range of uint8 (1-byte inputs) or uint16 (all others), truncating like the
 * reference. src / dst: host or device memory. out_min / out_max: the original extremes (as floats). */
tbrm_status tbrm_normalize_volume(int device, int voxel_format, const void* src, int src_is_device, uint64_t count, void* dst,
                                  int dst_is_device, float* out_min, float* out_max);
/* UVolumeTextureToolkit::ConvertArrayToFloat (TextureUtilities.cpp:329-350); TBRM_VOXEL_F32 input is rejected like there. */
tbrm_status tbrm_convert_volume_to_float(int device, int voxel_format, const void* src, int src_is_device, uint64_t count, float* dst,
                                         int dst_is_device);
/* UMHDLoader::CreateVolumeFromFile (MHDLoader.cpp:183-227) up to the texture: parses `mhd_path`, loads the raw or zlib data file next
 * to it, converts on the GPU (normalize: to G8 / G16 with min / max recorded in info; else convert_to_float: to R32F; else as stored)
 * and creates a resource set whose data volume it is. Voxel formats with no texture format of the path (unnormalised 32-bit integers)
 * return TBRM_ERR_UNSUPPORTED. */
tbrm_status tbrm_load_mhd_volume(int device, const char* mhd_path, int normalize, int convert_to_float, tbrm_format light_fmt, int half_res,
                                 tbrm_volume_info* info, tbrm_resources** out);
/* UVolumeTextureToolkit::LoadRawIntoNewVolumeTextureAsset / LoadRawFileIntoArray / LoadZLibCompressedFileIntoArray (TextureUtilities.h:67-75,
 * 85-101; TextureUtilities.cpp:262-302): the same without a header — dims[0] x dims[1] x dims[2] voxels of `voxel_format` read from
 * `raw_path` (zlib-compressed when compressed_bytes > 0), converted like above; `info` is filled in as if a header had described the file. */
tbrm_status tbrm_load_raw_volume(int device, const char* raw_path, const int32_t dims[3], int voxel_format, int64_t compressed_bytes, int normalize,
                                 int convert_to_float, tbrm_format light_fmt, int half_res, tbrm_volume_info* info, tbrm_resources** out);

/* ---- queue control ------------------------------------------------------------------------------------ */
tbrm_status tbrm_flush(tbrm_resources* res); /* FlushRenderingCommands(): wait for the resource set's stream */
void* tbrm_stream(tbrm_resources* res);      /* the cudaStream_t ops are enqueued on */
/* Enqueue on a caller-owned cudaStream_t from now on (the engine owns the command list): waits for the work already
 * enqueued, the previous stream is destroyed if the library created it. The stream must outlive res. */
tbrm_status tbrm_set_stream(tbrm_resources* res, void* cuda_stream);
/* Device time in ms between two points of the resource set's stream: call begin, enqueue ops, call end
 * (end synchronises). Used by bench.py so kernels launched on this stream are timed on this stream. */
tbrm_status tbrm_timer_begin(tbrm_resources* res);
tbrm_status tbrm_timer_end(tbrm_resources* res, float* out_ms);

/* ---- synthetic inputs: written into device or host memory, U8 ------------------------ */
typedef enum tbrm_synth_kind { TBRM_SYNTH_SPHERE = 0, TBRM_SYNTH_PERLIN_CT = 1 } tbrm_synth_kind;
tbrm_status tbrm_synth_volume_u8(int device, int kind, const int32_t dims[3], uint32_t seed, void* dst,
                                 int dst_is_device);

#ifdef __cplusplus
}
#endif
#endif /* TBRM_H_ */
