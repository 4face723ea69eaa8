/* tbrm_oracle.h — C interface of the CPU oracle (TEST INFRASTRUCTURE ONLY, see tbrm_oracle.cpp header). */
#ifndef TBRM_ORACLE_H_
#define TBRM_ORACLE_H_
#include <stdint.h>
#include "../include/tbrm.h" /* POD parameter types only */

#ifdef __cplusplus
extern "C" {
#endif

/* One axis pass of one light. Field-for-field comparable with tbrm_pass_plan. */
typedef struct tbo_pass {
    int32_t face, axis, dirn;
    int32_t td[3];
    int32_t start, stop;
    float weight, light_alpha, border;
    float uv_offset[2];
    float uvw_offset[3];
    float step_size;
} tbo_pass;

typedef struct tbo_light_plan {
    int32_t zero_direction;
    int32_t add_passes; /* passes AddDirLight executes (0..2); Change always runs 2 */
    tbo_pass pass[2];
    float clip_center[3], clip_dir[3];
    float data_border;
    double local_dir[3];
} tbo_light_plan;

typedef struct tbo_volume {
    const void* data;
    int32_t ddims[3];
    int32_t data_fmt;
    void* light;
    int32_t ldims[3];
    int32_t light_fmt;
    const float* tf; /* 256 x RGBA, output of tbo_prepare_tf */
    tbrm_windowing win;
    int32_t border_exact;
    int32_t data_addr_wrap;
} tbo_volume;

/* The fp32 uniforms the stand-in camera (tbrm_camera, include/tbrm.h) hands to the pixel shader: ResolvedView.WorldCameraOrigin,
 * the camera basis that yields MaterialParameters.CameraVector per pixel, GetPrimitiveData().WorldToLocal, scene depth.
 * Same layout as the kernels' RayCam (csrc/raymarch.cu). */
typedef struct tbo_camera_uniforms {
    float eye[3], fwd[3], rt[3], ut[3]; /* rt = right * tan(hfov/2), ut = up * tan(hfov/2) * H/W */
    float inv_w2, inv_h2;               /* 2/W, 2/H */
    float m[4][3];                      /* WorldToLocal, row-vector convention: local = [w,1] * M */
    float depth;
    int32_t width, height, frame_mod8, jitter;
} tbo_camera_uniforms;
void tbo_make_camera_uniforms(const tbrm_camera* cam, const tbrm_world* world, tbo_camera_uniforms* out);

void tbo_prepare_tf(const float* rgba, int width, int height, float* out_256x4);
void tbo_default_tf(float* out_256x4);
int tbo_plan_dir_light(const int32_t ldims[3], const tbrm_windowing* win, int border_exact, const tbrm_dir_light* light,
                       const tbrm_world* world, tbo_light_plan* out);
int tbo_clear_light_volume(void* light, const int32_t ldims[3], int light_fmt, float value);
int tbo_add_dir_light(const tbo_volume* vol, const tbrm_dir_light* light, int added, const tbrm_world* world, uint8_t* near_gate);
/* the CPU twin of tbrm_add_dir_lights_joined (same-face passes of several lights in one sweep; not in the reference); returns the sweeps run */
int tbo_add_dir_lights_joined(const tbo_volume* vol, const tbrm_dir_light* lights, int n_lights, int added, const tbrm_world* world);
int tbo_change_dir_light(const tbo_volume* vol, const tbrm_dir_light* old_light, const tbrm_dir_light* new_light,
                         const tbrm_world* world, uint8_t* near_gate);
int tbo_raymarch_cube_setup(const tbrm_camera* cam, const tbrm_world* world, float* out);
int tbo_raymarch_lit(const tbo_volume* vol, const tbrm_camera* cam, const tbrm_world* world, float step_count, int row_begin,
                     int row_end, float* out_rgba, uint64_t* out_steps, uint8_t* near_gate);
int tbo_mandelbulb_march(const tbrm_mandelbulb* mb, const tbrm_camera* cam, const tbrm_world* world, int row_begin, int row_end,
                         float* out_xy, uint64_t* out_iterations);
/* the other materials, the octree, the Mandelbulb variants, volume ingest */
int tbo_raymarch_intensity(const tbo_volume* vol, const tbrm_camera* cam, const tbrm_world* world, float step_count, int row_begin,
                           int row_end, float* out_rgba, uint64_t* out_steps);
int tbo_generate_octree(const void* data, const int32_t ddims[3], int data_fmt, const int32_t odims[3], void* const* mips);
int tbo_raymarch_octree(const tbo_volume* vol, const tbrm_camera* cam, const tbrm_world* world, float step_count, int row_begin,
                        int row_end, const void* const* mips, const int32_t odims[3], int octree_mip, float* out_rgba, uint64_t* out_steps);
int tbo_mandelbulb_march_normal(const tbrm_mandelbulb* mb, float derivation_distance, const tbrm_camera* cam, const tbrm_world* world,
                                int row_begin, int row_end, float* out_rgba, uint64_t* out_iterations);
int tbo_mandelbulb_sdf(const int32_t dims[3], const float center[3], float extent, float power, int out_fmt, void* out,
                       uint64_t* out_iterations);
int tbo_normalize_array(int fmt, const void* in, uint64_t count, void* out, float* out_min, float* out_max);
int tbo_convert_to_float(int fmt, const void* in, uint64_t count, float* out);
int tbo_synth_volume_u8(int kind, const int32_t dims[3], uint32_t seed, uint8_t* out);
/* 0 (default): Mandelbulb_SDF as the reference writes it; 1: the CPU twin of the kernels' transcendental-free Power == 8 iteration */
void tbo_set_mandelbulb_variant(int variant);
float tbo_mandelbulb_sdf_at(const float pos[3], float bailout, float power, int iterations, int variant, uint64_t* out_iterations);
float tbo_det_pow(float x, float y);
float tbo_round_to_half(float x);
void tbo_sample_windowed_tf(float value, float step, const float* tf, const tbrm_windowing* w, float out[4]);
float tbo_sample_data(const void* data, const int32_t dims[3], int fmt, float u, float v, float w, int mode, float border);
uint32_t tbo_pcg16_x(int x, int y, int z);
int tbo_max_threads(void);
void tbo_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
