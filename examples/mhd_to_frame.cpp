// mhd_to_frame.cpp — the drop-in boundary end to end in plain C++ (no CUDA headers, no torch): MetaImage file -> GPU resources ->
// illumination sweep -> octree -> three materials -> frames on the host. Build (after `python -m tbraymarcherplugin_b200.build`):
//     g++ -std=c++17 -I. examples/mhd_to_frame.cpp -Ltbraymarcherplugin_b200 -ltbrm -Wl,-rpath,$PWD/tbraymarcherplugin_b200 -o mhd_to_frame
//     ./mhd_to_frame ct_head.mhd
// What the calls stand for in the plugin: UMHDLoader::CreateVolumeFromFile + ARaymarchVolume::InitializeRaymarchResources,
// URaymarchUtils::MakeDefaultTFTexture / ClearResourceLightVolumes / AddDirLightToSingleVolume / GenerateOctree, and the Custom nodes of
// M_Raymarch / M_Intensity_Raymarch / M_Octree_Raymarch.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "include/tbrm.h"

#define CHECK(call)                                                                                  \
    do {                                                                                             \
        const tbrm_status s_ = (call);                                                               \
        if (s_ != TBRM_OK) {                                                                         \
            std::fprintf(stderr, "%s: %s (%s)\n", #call, tbrm_status_string(s_), tbrm_last_error()); \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s volume.mhd [width height steps]   (default 1280 720 256)\n", argv[0]);
        return 2;
    }
    tbrm_volume_info info;
    tbrm_resources* res = nullptr;
    CHECK(tbrm_load_mhd_volume(/*device*/ 0, argv[1], /*normalize*/ 1, /*convert_to_float*/ 0, TBRM_FMT_R32F, /*half_res*/ 0, &info, &res));
    std::printf("%d x %d x %d voxels, original range [%g, %g], normalised to %s\n", info.dims[0], info.dims[1], info.dims[2], info.min_value,
                info.max_value, info.bytes_per_voxel > 1 ? "G16" : "G8");

    CHECK(tbrm_make_default_tf(res));
    // a CT-like window given in original units (e.g. Hounsfield): centre 300, width 1500, cut off below the window
    const tbrm_windowing win = {tbrm_volume_info_normalize_value(&info, 300.0f), tbrm_volume_info_normalize_range(&info, 1500.0f), 1, 0};
    CHECK(tbrm_set_windowing(res, &win));

    tbrm_world world = {};
    world.rotation[3] = 1.0;
    world.scale[0] = world.scale[1] = world.scale[2] = 1.0;
    world.clip.center[2] = 100000.0;  // "no clipping" sentinel of ARaymarchVolume (RaymarchVolume.cpp:640-641)
    world.clip.direction[2] = -1.0;

    const tbrm_dir_light lights[2] = {{{1.0, 0.4, -0.3}, 1.0f}, {{-0.2, -1.0, -0.5}, 0.6f}};
    CHECK(tbrm_clear_light_volume(res, 0.0f));
    for (const tbrm_dir_light& l : lights) {
        int added = 0;
        CHECK(tbrm_add_dir_light(res, &l, /*added*/ 1, &world, &added, /*gpu_sync: the fused single-launch sweep*/ 1));
        if (!added) return 1;
    }
    CHECK(tbrm_generate_octree(res));

    tbrm_camera cam = {};
    cam.eye[0] = -0.9, cam.eye[1] = -0.5, cam.eye[2] = 0.7;  // the unit cube sits at the origin, [-0.5, 0.5]^3
    cam.up[2] = 1.0;
    cam.hfov_deg = 60.0;
    cam.width = argc > 3 ? std::atoi(argv[2]) : 1280, cam.height = argc > 3 ? std::atoi(argv[3]) : 720;
    const float march_steps = argc > 4 ? (float) std::atof(argv[4]) : 256.0f;
    cam.jitter = 1;
    std::vector<float> frame((size_t) cam.width * cam.height * 4);
    uint64_t steps = 0;
    CHECK(tbrm_raymarch_lit(res, &cam, &world, march_steps, 0, cam.height, frame.data(), 0, &steps));
    std::printf("lit march:       %llu ray-steps\n", (unsigned long long) steps);
    CHECK(tbrm_raymarch_intensity(res, &cam, &world, march_steps, 0, cam.height, frame.data(), 0, &steps));
    std::printf("intensity march: %llu ray-steps\n", (unsigned long long) steps);
    CHECK(tbrm_raymarch_octree(res, &cam, &world, march_steps, /*OctreeVolumeMip*/ 1, 0, cam.height, frame.data(), 0, &steps));
    std::printf("octree march:    %llu ray-steps\n", (unsigned long long) steps);
    CHECK(tbrm_flush(res));
    CHECK(tbrm_destroy(res));
    return 0;
}
