// host_plan.hpp — host-side parameter math of the illumination sweep and the camera, in fp64 like UE5's FVector.
//
// Mirrors Source/Raymarcher/Private/Rendering/LightingShaderUtils.cpp (FMajorAxes::GetMajorAxes :29-46,
// GetTransposedDimensions :48-64, GetAxisDirection :66-70, GetUVOffset :82-129, GetStepSizeAndUVWOffset :132-158,
// GetLocalLightParamsAndAxes :160-188, GetBorderColorIntSingle :197-203, GetLocalClippingParameters :205-220,
// GetLightAlpha :222-225, GetLoopStartStopIndexes :251-265) and the per-axis set-up of
// LightingShaders.cpp:100-130. Pure C++ (no CUDA), so the CPU test-suite can exercise it through the C ABI.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>

#include "../../include/tbrm.h"

namespace tbrm {
namespace host {

struct Vec3d {
    double x = 0, y = 0, z = 0;
    Vec3d() = default;
    Vec3d(double a, double b, double c) : x(a), y(b), z(c) {}
    explicit Vec3d(const double* p) : x(p[0]), y(p[1]), z(p[2]) {}
    Vec3d operator-() const { return {-x, -y, -z}; }
    Vec3d operator-(const Vec3d& o) const { return {x - o.x, y - o.y, z - o.z}; }
    Vec3d operator+(const Vec3d& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Vec3d operator*(double s) const { return {x * s, y * s, z * s}; }
    Vec3d operator*(const Vec3d& o) const { return {x * o.x, y * o.y, z * o.z}; }
    // FVector::operator/=(scalar) multiplies by the reciprocal
    Vec3d divided_by(double s) const {
        const double r = 1.0 / s;
        return {x * r, y * r, z * r};
    }
    double size() const { return std::sqrt(x * x + y * y + z * z); }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    // FVector::Normalize(SMALL_NUMBER)
    void normalize() {
        const double sq = x * x + y * y + z * z;
        if (sq > 1e-8) {
            const double s = 1.0 / std::sqrt(sq);
            x *= s, y *= s, z *= s;
        }
    }
    static Vec3d cross(const Vec3d& a, const Vec3d& b) {
        return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    }
};

// FTransform subset
struct TransformD {
    Vec3d translation, scale;
    double q[4];  // x, y, z, w
    explicit TransformD(const tbrm_world& w) : translation(w.translation), scale(w.scale) {
        std::memcpy(q, w.rotation, sizeof(q));
    }
    Vec3d unrotate(const Vec3d& v) const {  // FQuat::UnrotateVector
        const Vec3d qn(-q[0], -q[1], -q[2]);
        const Vec3d t = Vec3d::cross(qn, v) * 2.0;
        return v + t * q[3] + Vec3d::cross(qn, t);
    }
    static double safe_reciprocal(double s) { return std::fabs(s) <= 1e-8 ? 0.0 : 1.0 / s; }
    Vec3d inv_scale() const { return {safe_reciprocal(scale.x), safe_reciprocal(scale.y), safe_reciprocal(scale.z)}; }
    Vec3d inverse_transform_vector(const Vec3d& v) const { return unrotate(v) * inv_scale(); }
    Vec3d inverse_transform_vector_no_scale(const Vec3d& v) const { return unrotate(v); }
    Vec3d inverse_transform_position(const Vec3d& p) const { return unrotate(p - translation) * inv_scale(); }
};

inline double srgb_decode(double c) { return c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4); }
inline double srgb_encode(double c) { return c <= 0.0031308 ? c * 12.92 : 1.055 * std::pow(c, 1.0 / 2.4) - 0.055; }
inline double clamp_unit(double v) { return std::min(1.0, std::max(0.0, v)); }
inline double quantize8(double v) { return std::floor(v * 255.0 + 0.5) / 255.0; }

// FMajorAxes (LightingShaderUtils.h:46-55): faces sorted by weight
struct MajorAxes {
    std::array<std::pair<int, float>, 6> face_weight;
    static MajorAxes from_light_position(const Vec3d& light_pos) {
        static const double normals[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
        MajorAxes r;
        for (int i = 0; i < 6; ++i) {
            float w = (float) (normals[i][0] * light_pos.x + normals[i][1] * light_pos.y + normals[i][2] * light_pos.z);
            w = (w > 0 ? w * w : 0);
            r.face_weight[i] = {i, w};
        }
        // ties: stable, lower face index first (the reference's std::sort leaves this unspecified)
        std::stable_sort(r.face_weight.begin(), r.face_weight.end(),
                         [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second > b.second; });
        return r;
    }
};

inline void plan_clip(const tbrm_world& world, float center[3], float dir[3]) {
    const TransformD t(world);
    const Vec3d c = t.inverse_transform_position(Vec3d(world.clip.center)) + Vec3d(0.5, 0.5, 0.5);
    Vec3d d = t.inverse_transform_vector_no_scale(Vec3d(world.clip.direction)) * t.scale;
    d.normalize();
    center[0] = (float) c.x, center[1] = (float) c.y, center[2] = (float) c.z;
    dir[0] = (float) d.x, dir[1] = (float) d.y, dir[2] = (float) d.z;
}

inline float plan_data_border(const tbrm_windowing& win, bool exact) {
    const float zero_tf = win.center - 0.5f * win.width;  // LightingShaders.h:82
    if (exact) return zero_tf;
    return (float) srgb_decode(quantize8(clamp_unit((double) zero_tf)));
}

inline void plan_dir_light(const int32_t dims[3], const tbrm_windowing& win, bool border_exact, const tbrm_dir_light& light,
                           const tbrm_world& world, tbrm_light_plan& out) {
    std::memset(&out, 0, sizeof(out));
    if (light.direction[0] == 0.0 && light.direction[1] == 0.0 && light.direction[2] == 0.0) {
        out.zero_direction = 1;
        return;
    }
    const TransformD xf(world);
    Vec3d local_dir = xf.inverse_transform_vector(Vec3d(light.direction));
    local_dir.normalize();
    const Vec3d light_pos = -local_dir;
    MajorAxes axes = MajorAxes::from_light_position(light_pos);
    if (axes.face_weight[0].second > 0.99f) axes.face_weight[0].second = 1.0f;
    axes.face_weight[1].second = 1 - axes.face_weight[0].second;

    plan_clip(world, out.clip_center, out.clip_dir);
    out.data_border = plan_data_border(win, border_exact);
    out.local_dir[0] = local_dir.x, out.local_dir[1] = local_dir.y, out.local_dir[2] = local_dir.z;

    for (int i = 0; i < 2; ++i) {
        tbrm_pass_plan& p = out.pass[i];
        p.face = axes.face_weight[i].first;
        p.weight = axes.face_weight[i].second;
        p.axis = p.face / 2;
        static const int perm[3][3] = {{1, 2, 0}, {0, 2, 1}, {0, 1, 2}};  // (Y,Z,X) (X,Z,Y) (X,Y,Z)
        for (int k = 0; k < 3; ++k) p.td[k] = dims[perm[p.axis][k]];
        p.dirn = (p.face & 1) ? 1 : -1;
        p.start = p.dirn < 0 ? p.td[2] - 1 : 0;
        p.stop = p.dirn < 0 ? -1 : p.td[2];
        p.light_alpha = light.intensity * p.weight;
        p.border = border_exact ? p.light_alpha : (float) srgb_decode(quantize8(srgb_encode(clamp_unit((double) p.light_alpha))));

        const double major = light_pos[p.axis];
        if (major == 0.0) {
            // exactly axis-aligned light: this (weight-0) pass would divide by zero in the reference; it carries light
            // alpha 0 and contributes nothing. Zero offsets / step keep everything finite (policy Q11).
            continue;  // offsets and step size stay 0 from the memset
        }
        // GetUVOffset: divide by +major for positive faces, -major for negative ones, keep the two minor components
        const Vec3d q = light_pos.divided_by((p.face & 1) ? -major : major);
        const double minor_u = q[perm[p.axis][0]], minor_v = q[perm[p.axis][1]];
        const double inv_slices = 1.0 / (double) p.td[2];
        p.uv_offset[0] = (float) (minor_u * inv_slices);
        p.uv_offset[1] = (float) (minor_v * inv_slices);

        // GetStepSizeAndUVWOffset, then renormalise to the longest voxel side (LightingShaders.cpp:119-124)
        Vec3d off = light_pos.divided_by(std::fabs(major) * (double) p.td[2]);
        p.step_size = (float) off.size();
        const int lowest_voxel_count = std::min(p.td[0], std::min(p.td[1], p.td[2]));
        const float longest_voxel_side = 1.0f / lowest_voxel_count;
        off.normalize();
        off = off * (double) longest_voxel_side;
        p.uvw_offset[0] = (float) off.x, p.uvw_offset[1] = (float) off.y, p.uvw_offset[2] = (float) off.z;
    }
    out.add_passes = out.pass[0].weight == 0 ? 0 : (out.pass[1].weight == 0 ? 1 : 2);
}

// Camera uniforms for the raymarch kernels: basis in world space, WorldToLocal of the
// unit cube in UE's row-vector convention, all rounded once from fp64.
struct CameraUniforms {
    float eye[3], fwd[3], rt[3], ut[3];
    float inv_w2, inv_h2;
    float m[4][3];
    float depth;
    int width, height, frame_mod8, jitter;
};

inline void plan_camera(const tbrm_camera& cam, const tbrm_world& world, CameraUniforms& u) {
    const Vec3d eye(cam.eye);
    Vec3d f = Vec3d(cam.look_at) - eye;
    f = Vec3d(f.x / f.size(), f.y / f.size(), f.z / f.size());
    Vec3d r = Vec3d::cross(f, Vec3d(cam.up));
    r = Vec3d(r.x / r.size(), r.y / r.size(), r.z / r.size());
    const Vec3d up = Vec3d::cross(r, f);
    const double tan_x = std::tan(cam.hfov_deg * 3.14159265358979323846 / 360.0);
    const double tan_y = tan_x * (double) cam.height / (double) cam.width;
    const Vec3d rt = r * tan_x, ut = up * tan_y;
    for (int k = 0; k < 3; ++k) {
        u.eye[k] = (float) eye[k];
        u.fwd[k] = (float) f[k];
        u.rt[k] = (float) rt[k];
        u.ut[k] = (float) ut[k];
    }
    u.inv_w2 = 2.0f / (float) cam.width;
    u.inv_h2 = 2.0f / (float) cam.height;
    const TransformD xf(world);
    for (int i = 0; i < 3; ++i) {
        const Vec3d row = xf.inverse_transform_vector(Vec3d(i == 0, i == 1, i == 2));
        u.m[i][0] = (float) row.x, u.m[i][1] = (float) row.y, u.m[i][2] = (float) row.z;
    }
    const Vec3d t = xf.inverse_transform_position(Vec3d(0, 0, 0));
    u.m[3][0] = (float) t.x, u.m[3][1] = (float) t.y, u.m[3][2] = (float) t.z;
    u.depth = cam.scene_depth > 0.0f ? cam.scene_depth : 1e8f;
    u.width = cam.width, u.height = cam.height;
    u.frame_mod8 = cam.frame_index % 8;
    u.jitter = cam.jitter;
}

}  // namespace host
}  // namespace tbrm
