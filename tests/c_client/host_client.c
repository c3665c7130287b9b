/* host_client.c - a plain C99 client of include/rt_b200.h (test infrastructure).
 *
 * The drop-in boundary is a C ABI that a Rust `extern "C"` block or any other FFI binds. This
 * program is that kind of caller, written in C because no Rust toolchain exists in the build image: it includes the
 * header as C (so the header must be valid C99, not only C++), links librt_b200.so, and drives every host-side entry
 * point a front end needs - scene loading, flattened scene description, construction on a host-only handle
 * (RT_DEVICE_NONE), camera moves, octree introspection, Stats / BenchMark - and checks that render calls on such a
 * handle fail loudly with RT_ERR_CUDA instead of falling back to a CPU path. With a device argument it also renders.
 *
 * usage: host_client <file.dae> [device ordinal]
 * prints "key value" lines that tests/test_abi.py compares with the Python binding's results. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rt_b200.h"

#define CHECK(call)                                                          \
    do {                                                                     \
        int rc_ = (call);                                                    \
        if (rc_ != RT_OK) {                                                  \
            fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, err);      \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(int argc, char** argv) {
    char err[512] = "";
    char text[512];
    rt_scene* scene = NULL;
    rt_scene_desc desc;
    rt_config cfg;
    rt_raytracer* rt = NULL;
    uint64_t oct[6];
    float cam[34];
    uint32_t n_rays = 0;
    int rc;
    rt_stats* stats;
    rt_benchmark* bm;

    if (argc < 2) {
        fprintf(stderr, "usage: %s <file.dae> [device]\n", argv[0]);
        return 2;
    }
    printf("version %s\n", rt_version());
    CHECK(rt_scene_load_file(argv[1], &scene, err, sizeof err));
    CHECK(rt_scene_get_desc(scene, &desc));
    printf("triangles %u\ngeometries %u\nlights %u\ntextures %u\nfov %.9g\n", desc.num_triangles, desc.num_geometries, desc.num_lights,
           desc.num_textures, (double)desc.camera_fov_deg);
    /* reserved per-vertex attributes: the loader keeps none (the reference reads past them, colladaloader.rs:587-593) */
    printf("reserved_arrays %s\n", (desc.normals == NULL && desc.uvs == NULL) ? "null" : "set");

    rt_config_default(&cfg, 96, 54);
    cfg.recursions = 0;
    cfg.jitter_mode = RT_JITTER_FIXED_HALF;
    cfg.accel = RT_ACCEL_OCTREE;
    cfg.device = argc > 2 ? atoi(argv[2]) : RT_DEVICE_NONE;
    CHECK(rt_create(&desc, &cfg, &rt, err, sizeof err));
    rt_scene_free(scene); /* rt_create copied everything it needs */

    CHECK(rt_octree_stats(rt, oct));
    printf("octree_nodes %llu\noctree_refs %llu\noctree_depth %llu\n", (unsigned long long)oct[0], (unsigned long long)oct[4],
           (unsigned long long)oct[5]);
    CHECK(rt_camera_move_rel(rt, 0.25f, 0.0f, -0.5f));
    CHECK(rt_camera_add_x_angle(rt, 0.125f));
    CHECK(rt_camera_add_y_angle(rt, -0.0625f));
    CHECK(rt_camera_get(rt, cam));
    printf("camera_max_x %.9g\ncamera_rot0 %.9g\ncamera_pos %.9g %.9g %.9g\n", (double)cam[32], (double)cam[0], (double)cam[28], (double)cam[29],
           (double)cam[30]);

    rc = rt_trace_frame_additive(rt, &n_rays);
    if (cfg.device == RT_DEVICE_NONE) {
        /* no CPU fallback: a host-only handle must refuse to render and say why */
        printf("trace_rc %d\ntrace_error %s\n", rc, rt_last_error(rt));
        if (rc != RT_ERR_CUDA) return 3;
    } else {
        uint32_t* frame = (uint32_t*)malloc((size_t)cfg.width * cfg.height * 4);
        unsigned long long sum = 0;
        uint32_t i;
        if (rc != RT_OK || !frame) return 4;
        CHECK(rt_get_tonemapped_pixels(rt, frame));
        for (i = 0; i < cfg.width * cfg.height; ++i) sum += frame[i];
        printf("trace_rc %d\nprimary_rays %u\nframe_checksum %llu\n", rc, n_rays, sum);
        {
            /* the band loop with one frame buffer the host keeps (rt_get_tonemapped_pixels_delta), then Film::get_estimated_variances */
            uint32_t* keep = (uint32_t*)calloc((size_t)cfg.width * cfg.height, 4);
            float* var = (float*)malloc((size_t)cfg.width * cfg.height * 3 * sizeof(float));
            unsigned long long sum_keep = 0, sum_full = 0, finite = 0;
            int call;
            if (!keep || !var) return 5;
            for (call = 0; call < 3; ++call) {
                CHECK(rt_trace_frame_additive(rt, &n_rays));
                CHECK(rt_get_tonemapped_pixels_delta(rt, keep));
            }
            CHECK(rt_get_tonemapped_pixels(rt, frame));
            for (i = 0; i < cfg.width * cfg.height; ++i) {
                sum_keep += keep[i];
                sum_full += frame[i];
            }
            CHECK(rt_get_estimated_variances(rt, var));
            for (i = 0; i < cfg.width * cfg.height * 3; ++i) finite += (var[i] == var[i]) ? 1u : 0u;
            printf("delta_matches_full %d\nvariance_finite_values %llu\n", sum_keep == sum_full, finite);
            free(keep);
            free(var);
        }
        free(frame);
    }
    rt_destroy(rt);

    stats = rt_stats_new();
    CHECK(rt_stats_stats(stats, 96u * 50u, text, sizeof text));
    printf("stats_prefix %.5s\n", text);
    CHECK(rt_stats_mean_stats(stats, text, sizeof text));
    printf("mean_stats_prefix %.9s\n", text);
    rt_stats_free(stats);

    bm = rt_benchmark_new();
    CHECK(rt_benchmark_start(bm, "frame"));
    CHECK(rt_benchmark_stop(bm, "frame"));
    printf("benchmark_unknown_rc %d\n", rt_benchmark_stop(bm, "never started"));
    CHECK(rt_benchmark_report(bm, text, sizeof text));
    printf("benchmark_report %.12s\n", text);
    rt_benchmark_free(bm);
    return 0;
}
