// raytracer.cu — device-resident RayTracer and the C ABI declared in include/rt_b200.h.
//
// Host-side mirror of raytracer_lib's public surface (lib.rs:15-44, raytracer/mod.rs:32-128): construction wires
// loader -> acceleration structure -> renderer, the render calls launch the fused kernel of kernels.cu on the
// handle's stream. There is deliberately no CPU rendering path in this library.
#include <cuda.h>  // types of the two driver entry points fetched at run time (no link dependency on libcuda)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "accel_build.h"
#include "device_types.h"
#include "host_scene.h"
#include "kernels.h"
#include "pgrid_build.h"

using namespace rtb;

struct rt_scene {
    HostScene scene;
};

namespace {

struct CudaFail {
    std::string what;
};
#define RT_CUDA_RET(expr)                      \
    do {                                       \
        cudaError_t e__ = (expr);              \
        if (e__ != cudaSuccess) return e__;    \
    } while (0)
#define RT_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) throw CudaFail{std::string(#expr) + ": " + cudaGetErrorString(e__)};     \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        RT_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
        n = count;
    }
    void upload(const std::vector<T>& h, cudaStream_t s) {
        alloc(h.size());
        if (!h.empty()) RT_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    }
};

// No C++ exception may cross the extern "C" boundary (std::bad_alloc from a huge texture, a parser fault): it becomes a load error.
template <typename F>
bool load_guarded(F&& f, std::string* err) {
    try {
        return f();
    } catch (std::exception& ex) {
        *err = std::string("scene loading failed: ") + ex.what();
    } catch (...) {
        *err = "scene loading failed";
    }
    return false;
}

void copy_err(const std::string& msg, char* err, size_t err_len) {
    if (err && err_len) {
        std::snprintf(err, err_len, "%s", msg.c_str());
    }
}

}  // namespace

struct rt_raytracer {
    rt_config cfg{};
    HostScene scene;
    HostCamera camera;
    FlatOctree octree;
    bool octree_built = false;
    FlatBvh bvh;
    bool bvh_built = false;
    FlatCwbvh cwbvh;
    bool cwbvh_built = false;
    FlatBvh4 bvh4;
    bool bvh4_built = false;
    // GPU-built binary BVH (RT_ACCEL_LBVH)
    DevBuf<float4> d_lbvh_nodes, d_lbvh_tris;
    DevBuf<uint32_t> d_lbvh_order, d_lbvh_depth;
    DevBuf<float> d_verts;
    uint32_t lbvh_depth = 0, lbvh_nodes = 0;
    float lbvh_build_ms = 0.f;
    std::string last_error;
    int device = 0;
    cudaStream_t stream = nullptr;

    // device data
    DevBuf<float4> d_oct_nodes, d_oct_tris, d_bvh_nodes, d_bvh_tris, d_cw_tris, d_bvh4_nodes, d_bvh4_tris, d_tri_shade, d_materials, d_lights;
    DevBuf<CwWord> d_cw_nodes;
    std::vector<std::unique_ptr<DevBuf<float>>> d_tex_data;
    DevBuf<DevTexture> d_textures;
    DevBuf<float4> d_film_sum, d_film_sq, d_planes;
    DevBuf<float> d_variance;  // rt_get_estimated_variances staging (allocated on first use)
    int multi_sample_launch = 1;      // RT_TUNE_MULTI_SAMPLE_LAUNCH: 0 one launch per sample, 1 sample lanes where they apply, else planes, 2 planes
    bool bounce_wavefront = true;     // RT_TUNE_BOUNCE_WAVEFRONT
    bool bounce_stream = true;        // RT_TUNE_BOUNCE_STREAM: wavefront levels as a ray stream (binary BVH) instead of lockstep warps
    bool stream_chain = true;         // RT_TUNE_STREAM_CHAIN
    int stream_blocks = 4;            // RT_TUNE_STREAM_BLOCKS: resident blocks per SM the ray-stream kernel is compiled for (3, 4, 5)
    int wf_blocks_cap = 0;            // RT_TUNE_WF_BLOCKS: cap on the resident blocks per SM of the lockstep wavefront kernels (0 = as many as fit)
    int wf_occupancy[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    int stream_refill = 16;           // RT_TUNE_STREAM_REFILL: lanes without a ray in flight that trigger a service round
    int stream_min_inner = 8;         // RT_TUNE_STREAM_MIN_INNER: the inner-node loop yields to waiting leaves below this many descending lanes
    DevBuf<float4> d_wf_rec;
    DevBuf<float> d_wf_child;
    DevBuf<unsigned int> d_wf_counts;
    DevBuf<uint32_t> d_ldr, d_ids, d_row_list, d_owned_rows;
    DevBuf<unsigned long long> d_counters;
    DevBuf<uint32_t> d_sync_timeouts;
    unsigned long long* h_counters = nullptr;  // pinned
    uint32_t* ldr_remote = nullptr;
    uint32_t* host_frame = nullptr;       // registered zero-copy frame (host address)
    uint32_t* host_frame_dev = nullptr;   // its device-visible address
    bool host_frame_stale = true;         // rows traced before registration / film clear are not in it yet
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    // pipelined readback (rt_get_tonemapped_pixels_async): snapshot on the render stream, device -> host on a copy stream
    // (two snapshots: the copy of frame k may still be running while frame k+1 is snapshotted and queued behind it, so the copy engine
    // never waits for the host)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_snapshot = nullptr, ev_copied[2] = {nullptr, nullptr};
    DevBuf<uint32_t> d_ldr_snapshot[2];
    uint64_t copies_issued = 0, copies_waited = 0;  // copy c uses snapshot / event c & 1

    // host state
    uint32_t current_row = 0;
    std::vector<uint32_t> owned_rows;       // all rows of this shard, ascending
    std::vector<uint32_t> row_list_cache;   // rows of the last sharded / wrapped launch
    std::vector<uint32_t> lap_start;        // offsets into row_list_cache where each lap over the image begins
    uint32_t cached_first = ~0u, cached_n = ~0u;
    uint64_t total_kernels = 0, total_primary = 0;
    int variant = 1;            // RT_TUNE_KERNEL_VARIANT
    int split_quarters = 4;     // RT_TUNE_SPLIT_QUARTERS
    int queue_batch = 4, queue_batch_from_pct = 33;  // RT_TUNE_QUEUE_BATCH, RT_TUNE_QUEUE_BATCH_FROM
    int pool_refill = 16;       // RT_TUNE_POOL_REFILL
    int pool_min_inner = 8;     // RT_TUNE_POOL_MIN_INNER
    int pool_blocks = 0;        // resident blocks per SM of the ray-pool kernel
    int time_launches = 1;      // RT_TUNE_TIME_LAUNCHES: record the two CUDA events behind rt_launch_stats.trace_kernel_ms
    bool last_timed = false;    // the last trace call recorded them
    int lpt_schedule = 1;       // RT_TUNE_TILE_SCHEDULE: 1 = heaviest tiles first (cost feedback), 0 = image order
    // cost-feedback schedules, one per launch geometry (first row, rows, item shape): a host that walks over the image in 50-row
    // bands (trace_frame_additive, mod.rs:87) retraces the same 22 geometries frame after frame and finds each band's schedule again
    struct TileSchedule {
        uint32_t first = ~0u, rows = ~0u, tiles = 0, samples_log2 = 0;
        int kernel = -1;            // instantiation the costs were recorded with (a tile costs the grid kernels and the tree walkers differently)
        uint32_t launches = 0;      // launches recorded since the schedule was created / the view last changed
        uint32_t recorded = 0;      // launches recorded since the schedule was created
        bool have_order = false;    // `order` holds a queue
        bool restart_costs = true;  // the next launch starts from zeroed costs (new schedule, or the view changed)
        uint32_t max_level = 0;     // finest split the order buffer has room for (kernels.cu, tile_sort_kernel)
        uint64_t last_use = 0;
        DevBuf<uint32_t> cost, order;  // order: tiles << (max_level + 1) items, then the item count
    };
    std::vector<std::unique_ptr<TileSchedule>> schedules;
    uint64_t schedule_clock = 0;
    int min_schedule_tiles = 4096;   // RT_TUNE_MIN_SCHEDULE_TILES: launches with fewer tiles run in image order, whole tiles
    int max_split_level = 3;         // RT_TUNE_MAX_SPLIT_LEVEL: 0 never split, 1 / 2 / 3 = up to 4 / 8 / 16 items per tile
    int call_parity = 0;             // which per-call counter set the current trace call counts into (device_types.h)
    bool set_dirty[2] = {false, false};  // the set holds counts of an earlier call (a launch with warp_checkout cleans the OTHER set)
    // start of a trace call: switch to the other per-call counter set; it is zero already unless the previous call launched no kernel
    // that checks out (a pure commit of the band look-ahead), then it is zeroed here
    void begin_call_counters() {
        call_parity ^= 1;
        if (set_dirty[call_parity])
            RT_CUDA(cudaMemsetAsync(d_counters.p + (call_parity ? CNT_SET_B : CNT_SET_A), 0, 4 * sizeof(unsigned long long), stream));
        set_dirty[call_parity] = true;
    }
    uint32_t* done_flag = nullptr;   // multi-GPU: the trace kernel's last warp publishes done_value here (rt_set_done_signal)
    uint32_t done_value = 0;
    bool arm_done = false;           // the launch being issued is the last one of its trace call
    bool done_published = false;     // ... and carried the signal
    int blocks_per_sm[5][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};  // [kernel accel][bounce]; 4 = binary BVH + camera grid
    // perspective grid of the camera rays (pgrid_build.cu): rebuilt when the camera, the resolution or the triangle array changes
    int camera_grid_log2 = 3;  // RT_TUNE_CAMERA_GRID: 0 = camera rays walk the BVH, 2..5 = grid cells of 4..32 pixels
    DevBuf<uint32_t> d_pg_count, d_pg_start, d_pg_cursor, d_pg_total, d_pg_big;
    DevBuf<uint2> d_pg_entries;
    struct PGridKey {
        float cam[21];  // rotation[16], ray origin[3], max_x, max_y
        uint32_t w, h, shift, n_slots;
        const float4* tris;
        PGridKey() { std::memset(this, 0, sizeof(*this)); }  // compared with memcmp: padding included
    } pg_key;
    // cube of grids around every point light, for the shadow rays of ACCEL = 4 kernels: built once per (triangle array, lights)
    int light_grid_min_tris = 256;
    int light_grid_log2 = 8;  // RT_TUNE_LIGHT_GRID: 0 = shadow rays walk the BVH, 6..9 = 64..512 cells per cube-face edge (8 grid units each)
    DevBuf<uint32_t> d_lg_count, d_lg_start, d_lg_cursor, d_lg_total, d_lg_big;
    DevBuf<uint2> d_lg_entries;
    DevBuf<float> d_lg_dmin2;
    const float4* lg_tris_key = nullptr;
    uint32_t lg_slots_key = 0, lg_n = 0, lg_entries = 0, lg_cells = 0;
    int lg_log2_key = -1;
    bool lg_valid = false, lg_off = false;
    static constexpr uint32_t kGridMaxEntries = 1u << 26;  // 512 MB of (slot, key) pairs
    float lg_far2[kGridLights] = {0.f, 0.f, 0.f, 0.f};
    PGridKey pg_seen;                 // the view of the last launch that had no grid
    uint32_t pg_seen_launches = 0;
    int camera_grid_after = 1;        // RT_TUNE_CAMERA_GRID_AFTER
    bool pg_valid = false;
    uint32_t pg_nx = 0, pg_shift = 0, pg_entries = 0, pg_cells = 0;
    uint64_t pg_builds = 0;
    int num_sms = 0;
    rt_launch_stats last{};
    bool stats_pending = false;

    ~rt_raytracer() {
        if (h_counters) cudaFreeHost(h_counters);
        if (ev_start) cudaEventDestroy(ev_start);
        if (ev_stop) cudaEventDestroy(ev_stop);
        if (ev_snapshot) cudaEventDestroy(ev_snapshot);
        for (cudaEvent_t e : ev_copied)
            if (e) cudaEventDestroy(e);
        if (copy_stream) cudaStreamDestroy(copy_stream);
    }

    bool host_only = false;  // cfg.device == RT_DEVICE_NONE: construction, camera and accel introspection only
    void bind_device() {
        if (host_only) throw CudaFail{"this handle was created with RT_DEVICE_NONE (host-side introspection only); rendering needs a B200"};
        RT_CUDA(cudaSetDevice(device));
    }
    uint32_t npix() const { return cfg.width * cfg.height; }
    bool sharded() const { return cfg.shard_count > 1; }
    bool owns_row(uint32_t r) const {
        if (!sharded()) return true;
        return ((r / cfg.band_rows) % cfg.shard_count) == cfg.shard_index;
    }

    void init(const HostScene& sc, const rt_config& c) {
        cfg = c;
        if (cfg.width == 0 || cfg.height == 0) throw CudaFail{"width and height must be positive"};
        if (cfg.triangles_per_leaf == 0) cfg.triangles_per_leaf = RT_DEFAULT_TRIANGLES_PER_LEAF;
        if (cfg.rows_per_call == 0) cfg.rows_per_call = 50;
        if (cfg.band_rows == 0) cfg.band_rows = 8;
        if (cfg.shard_count == 0) cfg.shard_count = 1;
        if (cfg.shard_index >= cfg.shard_count) throw CudaFail{"shard_index out of range"};
        scene = sc;
        dirty_rows.assign(cfg.height, 1);
        camera.init(cfg.width, cfg.height, scene.camera_orientation, scene.camera_fov_deg);
        for (uint32_t r = 0; r < cfg.height; ++r)
            if (owns_row(r)) owned_rows.push_back(r);
        if (cfg.device == RT_DEVICE_NONE) {
            host_only = true;
            return;
        }
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw CudaFail{std::string("no CUDA device available (") + cudaGetErrorString(e) + "); rt_b200 has no CPU fallback"};
        if (cfg.device >= 0) {
            device = cfg.device;
        } else {
            RT_CUDA(cudaGetDevice(&device));
        }
        bind_device();
        cudaDeviceProp prop;
        RT_CUDA(cudaGetDeviceProperties(&prop, device));
        num_sms = prop.multiProcessorCount;
        if (prop.major != 10) throw CudaFail{std::string("device '") + prop.name + "' is not sm_100 (Blackwell B200); rt_b200 ships sm_100a code only"};
        RT_CUDA(cudaEventCreate(&ev_start));
        RT_CUDA(cudaEventCreate(&ev_stop));
        RT_CUDA(cudaHostAlloc((void**)&h_counters, CNT_SLOTS * sizeof(unsigned long long), cudaHostAllocDefault));
        d_counters.alloc(CNT_SLOTS);
        RT_CUDA(cudaMemsetAsync(d_counters.p, 0, CNT_SLOTS * sizeof(unsigned long long), stream));
        upload_scene();
        ensure_accel(cfg.accel);
        d_film_sum.alloc(npix());
        d_film_sq.alloc(npix());
        d_ldr.alloc(npix());
        d_ids.alloc(npix());
        d_owned_rows.upload(owned_rows, stream);
        film_clear();
        RT_CUDA(cudaStreamSynchronize(stream));
    }

    float root_lo[3], root_hi[3];  // scene AABB (= octree root cube, calc_extents :315-330), computed once

    void upload_scene() {
        const uint32_t nt = scene.num_triangles();
        for (int a = 0; a < 3; ++a) {
            root_lo[a] = 3.402823466e+38f;
            root_hi[a] = -3.402823466e+38f;
        }
        for (size_t k = 0; k < scene.vertices.size(); ++k) {
            root_lo[k % 3] = std::fmin(root_lo[k % 3], scene.vertices[k]);
            root_hi[k % 3] = std::fmax(root_hi[k % 3], scene.vertices[k]);
        }
        std::vector<float4> shade(nt);
        for (uint32_t t = 0; t < nt; ++t) {
            const float* v = &scene.vertices[9 * (size_t)t];
            const f3 v0{v[0], v[1], v[2]}, v1{v[3], v[4], v[5]}, v2{v[6], v[7], v[8]};
            const f3 n = unit3(cross3(v1 - v0, v2 - v0));  // calc_normal, mod.rs:198-205
            uint32_t g = scene.tri_geom[t];
            float gw;
            std::memcpy(&gw, &g, 4);
            shade[t] = make_float4(n.x, n.y, n.z, gw);
        }
        d_tri_shade.upload(shade, stream);
        std::vector<float4> mats(scene.materials.size());
        for (size_t g = 0; g < mats.size(); ++g) {
            const rt_material& m = scene.materials[g];
            int32_t tex = m.kind == RT_DIFFUSE_TEXTURE ? (int32_t)m.texture_id : -1;
            if (tex >= (int32_t)scene.textures.size()) throw CudaFail{"material references a texture that does not exist"};
            float tw;
            std::memcpy(&tw, &tex, 4);
            mats[g] = make_float4(m.rgb[0], m.rgb[1], m.rgb[2], tw);
        }
        d_materials.upload(mats, stream);
        std::vector<float4> lights;
        for (const rt_light& l : scene.lights) {
            lights.push_back(make_float4(l.pos[0], l.pos[1], l.pos[2], 0.f));
            lights.push_back(make_float4(l.color[0], l.color[1], l.color[2], 0.f));
        }
        d_lights.upload(lights, stream);
        std::vector<DevTexture> tex;
        for (const HostTexture& t : scene.textures) {
            auto buf = std::make_unique<DevBuf<float>>();
            buf->upload(t.rgb, stream);
            tex.push_back(DevTexture{buf->p, t.width, t.height});
            d_tex_data.push_back(std::move(buf));
        }
        d_textures.upload(tex, stream);
    }

    static float4 tri_word0(const float* v) { return make_float4(v[0], v[1], v[2], v[3] - v[0]); }
    static void pack_triangle(const float* v, uint32_t id, float4* out) {
        // e1 = v1 - v0, e2 = v2 - v0: the subtractions intersect.rs:66-67 performs per ray
        const float e1x = v[3] - v[0], e1y = v[4] - v[1], e1z = v[5] - v[2];
        const float e2x = v[6] - v[0], e2y = v[7] - v[1], e2z = v[8] - v[2];
        float idw;
        std::memcpy(&idw, &id, 4);
        out[0] = make_float4(v[0], v[1], v[2], e1x);
        out[1] = make_float4(e1y, e1z, e2x, e2y);
        out[2] = make_float4(e2z, idw, 0.f, 0.f);
    }

    // SampleGenerator::new (sample_generator.rs:15-24, 35-52): 65 536 unit vectors, rejection sampled in the unit
    // ball. The reference draws them from rand::rng(); here they come from the shared counter-based hash so that
    // oracle and GPU hold the same table for a given seed.
    DevBuf<float> d_sample_table;
    uint32_t sample_table_seed = 0;
    bool sample_table_ready = false;
    static uint32_t mix32(uint32_t h) {
        h ^= h >> 16;
        h *= 0x7feb352dU;
        h ^= h >> 15;
        h *= 0x846ca68bU;
        h ^= h >> 16;
        return h;
    }
    static uint32_t hash4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        uint32_t h = mix32(a + 0x9e3779b9U);
        h = mix32(h ^ (b + 0x85ebca6bU));
        h = mix32(h ^ (c + 0xc2b2ae35U));
        h = mix32(h ^ (d + 0x27d4eb2fU));
        return h;
    }
    void ensure_sample_table() {
        if (sample_table_ready && sample_table_seed == cfg.seed) return;
        std::vector<float> table(65536 * 3);
        const uint32_t key = cfg.seed ^ 0x5a17ab1eU;
        for (uint32_t i = 0; i < 65536; ++i) {
            for (uint32_t attempt = 0;; ++attempt) {
                f3 v;
                v.x = (float)(hash4(key, i, attempt, 0) >> 8) * (1.0f / 16777216.0f) * 2.0f - 1.0f;
                v.y = (float)(hash4(key, i, attempt, 1) >> 8) * (1.0f / 16777216.0f) * 2.0f - 1.0f;
                v.z = (float)(hash4(key, i, attempt, 2) >> 8) * (1.0f / 16777216.0f) * 2.0f - 1.0f;
                const float len2 = dot3(v, v);
                if (len2 < 1.0f && len2 > 0.0f) {
                    const f3 u = unit3(v);
                    table[3 * i] = u.x;
                    table[3 * i + 1] = u.y;
                    table[3 * i + 2] = u.z;
                    break;
                }
            }
        }
        d_sample_table.upload(table, stream);
        RT_CUDA(cudaStreamSynchronize(stream));  // `table` goes out of scope
        sample_table_seed = cfg.seed;
        sample_table_ready = true;
    }

    void ensure_octree_host() {
        if (octree_built) return;
        octree = build_octree(scene, cfg.triangles_per_leaf);
        octree_built = true;
    }

    void ensure_bvh_host() {
        if (bvh_built) return;
        uint32_t max_leaf = 4;
        if (const char* e = std::getenv("RT_BVH_MAX_LEAF")) max_leaf = (uint32_t)std::min(15, std::max(1, std::atoi(e)));  // developer override
        bvh = build_bvh(scene, max_leaf);
        bvh_built = true;
    }

    void ensure_bvh4_host() {
        if (bvh4_built) return;
        bvh4 = build_bvh4(scene);
        bvh4_built = true;
    }

    void ensure_cwbvh_host() {
        if (cwbvh_built) return;
        cwbvh = build_cwbvh(scene);
        cwbvh_built = true;
    }

    void ensure_accel(int accel) {
        if (accel == RT_ACCEL_OCTREE) {
            ensure_octree_host();
            if (d_oct_nodes.p) return;
            if (octree.max_stack > (uint32_t)kOctStack) throw CudaFail{"octree deeper than the traversal stack"};
            const size_t nn = octree.num_nodes();
            std::vector<float4> nodes(2 * nn);
            for (size_t i = 0; i < nn; ++i) {
                const float* c = &octree.cubes[6 * i];
                uint32_t a_w, b_w;
                if (octree.first_child[i] < 0) {
                    a_w = octree.leaf_offset[i];
                    b_w = kOctLeafFlag | (octree.leaf_offset[i + 1] - octree.leaf_offset[i]);
                } else {
                    a_w = (uint32_t)octree.first_child[i];
                    b_w = 0;
                    for (int k = 0; k < 8; ++k) {
                        const size_t ch = (size_t)octree.first_child[i] + k;
                        const bool empty_leaf = octree.first_child[ch] < 0 && octree.leaf_offset[ch + 1] == octree.leaf_offset[ch];
                        if (!empty_leaf) b_w |= 1u << k;
                    }
                }
                float aw, bw;
                std::memcpy(&aw, &a_w, 4);
                std::memcpy(&bw, &b_w, 4);
                nodes[2 * i] = make_float4(c[0], c[1], c[2], aw);
                nodes[2 * i + 1] = make_float4(c[3], c[4], c[5], bw);
            }
            std::vector<float4> tris(3 * octree.leaf_tris.size());
            for (size_t r = 0; r < octree.leaf_tris.size(); ++r) {
                const uint32_t t = octree.leaf_tris[r];
                pack_triangle(&scene.vertices[9 * (size_t)t], t, &tris[3 * r]);
            }
            d_oct_nodes.upload(nodes, stream);
            d_oct_tris.upload(tris, stream);
        } else if (accel == RT_ACCEL_BVH) {
            if (d_bvh_nodes.p) return;
            ensure_bvh_host();
            if (bvh.depth + 2 > (uint32_t)kBvhStack) throw CudaFail{"BVH deeper than the traversal stack"};
            // The device array holds the nodes in BREADTH-FIRST order (the builder numbers them depth first): the top levels, which
            // every ray visits, are then one contiguous run of cache lines (and the prefix an RT_TOP_SMEM build stages in shared
            // memory). Only the numbering changes; rt_bvh_export keeps the builder's.
            std::vector<int32_t> bfs_of(bvh.nodes.size(), -1), order;
            order.reserve(bvh.nodes.size());
            order.push_back(0);
            bfs_of[0] = 0;
            for (size_t q = 0; q < order.size(); ++q)
                for (int k = 0; k < 2; ++k) {
                    const int32_t c = bvh.nodes[(size_t)order[q]].child[k];
                    if (c >= 0) {
                        bfs_of[(size_t)c] = (int32_t)order.size();
                        order.push_back(c);
                    }
                }
            if (order.size() != bvh.nodes.size()) throw CudaFail{"BVH has unreachable nodes"};
            std::vector<float4> nodes(4 * bvh.nodes.size());
            for (size_t i = 0; i < bvh.nodes.size(); ++i) {
                const FlatBvh::Node& n = bvh.nodes[(size_t)order[i]];
                int32_t ref[2];
                for (int k = 0; k < 2; ++k) {
                    if (n.child[k] >= 0)
                        ref[k] = bfs_of[(size_t)n.child[k]];
                    else
                        ref[k] = ~(int32_t)(((uint32_t)(~n.child[k]) << 4) | (uint32_t)n.count[k]);
                }
                float r0, r1;
                std::memcpy(&r0, &ref[0], 4);
                std::memcpy(&r1, &ref[1], 4);
                nodes[4 * i + 0] = make_float4(n.lo[0][0], n.lo[0][1], n.lo[0][2], n.hi[0][0]);
                nodes[4 * i + 1] = make_float4(n.hi[0][1], n.hi[0][2], n.lo[1][0], n.lo[1][1]);
                nodes[4 * i + 2] = make_float4(n.lo[1][2], n.hi[1][0], n.hi[1][1], n.hi[1][2]);
                nodes[4 * i + 3] = make_float4(r0, r1, 0.f, 0.f);
            }
            std::vector<float4> tris(3 * bvh.tri_order.size());
            for (size_t s = 0; s < bvh.tri_order.size(); ++s) {
                const uint32_t t = bvh.tri_order[s];
                pack_triangle(&scene.vertices[9 * (size_t)t], t, &tris[3 * s]);
            }
            d_bvh_nodes.upload(nodes, stream);
            d_bvh_tris.upload(tris, stream);
        } else if (accel == RT_ACCEL_LBVH) {
            if (d_lbvh_nodes.p) return;
            build_lbvh();
        } else if (accel == RT_ACCEL_BVH4) {
            if (d_bvh4_nodes.p) return;
            ensure_bvh4_host();
            if (3 * bvh4.depth + 5 > (uint32_t)kBvh4Stack) throw CudaFail{"4-wide BVH deeper than the traversal stack"};
            std::vector<float4> nodes(8 * bvh4.nodes.size());
            for (size_t i = 0; i < bvh4.nodes.size(); ++i) {
                const FlatBvh4::Node& n = bvh4.nodes[i];
                float ref[4];
                for (int k = 0; k < 4; ++k) {
                    const int32_t r = n.child[k] >= 0 ? n.child[k] : ~(int32_t)(((uint32_t)(~n.child[k]) << 4) | (uint32_t)n.count[k]);
                    std::memcpy(&ref[k], &r, 4);
                }
                for (int a = 0; a < 3; ++a) {
                    nodes[8 * i + a] = make_float4(n.lo[0][a], n.lo[1][a], n.lo[2][a], n.lo[3][a]);
                    nodes[8 * i + 3 + a] = make_float4(n.hi[0][a], n.hi[1][a], n.hi[2][a], n.hi[3][a]);
                }
                nodes[8 * i + 6] = make_float4(ref[0], ref[1], ref[2], ref[3]);
                nodes[8 * i + 7] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            std::vector<float4> tris(3 * bvh4.tri_order.size());
            for (size_t s = 0; s < bvh4.tri_order.size(); ++s) {
                const uint32_t t = bvh4.tri_order[s];
                pack_triangle(&scene.vertices[9 * (size_t)t], t, &tris[3 * s]);
            }
            d_bvh4_nodes.upload(nodes, stream);
            d_bvh4_tris.upload(tris, stream);
        } else if (accel == RT_ACCEL_CWBVH) {
            if (d_cw_nodes.p) return;
            ensure_cwbvh_host();
            if (cwbvh.depth + 3 > (uint32_t)kCwStack) throw CudaFail{"wide BVH deeper than the traversal stack"};
            std::vector<float4> tris(3 * cwbvh.tri_order.size());
            for (size_t s = 0; s < cwbvh.tri_order.size(); ++s) {
                const uint32_t t = cwbvh.tri_order[s];
                pack_triangle(&scene.vertices[9 * (size_t)t], t, &tris[3 * s]);
            }
            d_cw_nodes.upload(cwbvh.nodes, stream);
            d_cw_tris.upload(tris, stream);
        } else {
            throw CudaFail{"unknown accel"};
        }
    }

    // the binary BVH built on the device (lbvh_build.cu); timed with CUDA events on the handle's stream
    void build_lbvh() {
        const uint32_t n = scene.num_triangles();
        if (!d_verts.p) d_verts.upload(scene.vertices, stream);
        DevBuf<char> scratch;
        scratch.alloc(lbvh_scratch_bytes(n));
        lbvh_nodes = n > 1 ? n - 1 : 1;
        d_lbvh_nodes.alloc(4 * (size_t)lbvh_nodes);
        d_lbvh_tris.alloc(3 * (size_t)std::max(n, 1u));
        d_lbvh_order.alloc(std::max(n, 1u));
        d_lbvh_depth.alloc(1);
        RT_CUDA(cudaEventRecord(ev_start, stream));
        RT_CUDA(build_lbvh_device(d_verts.p, n, root_lo, root_hi, scratch.p, d_lbvh_nodes.p, d_lbvh_tris.p, d_lbvh_order.p, d_lbvh_depth.p, stream));
        RT_CUDA(cudaEventRecord(ev_stop, stream));
        RT_CUDA(cudaMemcpyAsync(&lbvh_depth, d_lbvh_depth.p, 4, cudaMemcpyDeviceToHost, stream));
        RT_CUDA(cudaStreamSynchronize(stream));
        RT_CUDA(cudaEventElapsedTime(&lbvh_build_ms, ev_start, ev_stop));
        total_kernels += n > 1 ? 16 : 3;
        if (lbvh_depth + 2 > (uint32_t)kBvhStack) {
            d_lbvh_nodes.release();
            throw CudaFail{"GPU-built BVH is deeper than the traversal stack (degenerate triangle distribution); use RT_ACCEL_BVH"};
        }
    }

    // rows of the packed frame that changed since the last rt_get_tonemapped_pixels_delta (see there)
    std::vector<uint8_t> dirty_rows;
    const void* delta_host = nullptr;  // host buffer that equals the device frame except for the dirty rows
    void mark_frame_dirty_all() { std::fill(dirty_rows.begin(), dirty_rows.end(), (uint8_t)1); }
    void mark_rows_dirty(uint32_t first_row, uint32_t n_rows) {
        if (n_rows >= cfg.height) return mark_frame_dirty_all();
        for (uint32_t k = 0; k < n_rows; ++k) dirty_rows[(first_row + k) % cfg.height] = 1;
    }

    void film_clear() {
        host_frame_stale = true;
        mark_frame_dirty_all();
        drop_lap();
        RT_CUDA(launch_film_clear(d_film_sum.p, d_film_sq.p, d_ldr.p, d_ids.p, npix(), stream));
        ++total_kernels;
    }

    void fill_params(TraceParams* p) {
        std::memset(p, 0, sizeof(*p));
        std::memcpy(p->cam.rot, camera.rotation.data(), 64);
        const f3 o = camera.ray_origin();
        p->cam.pos[0] = o.x;
        p->cam.pos[1] = o.y;
        p->cam.pos[2] = o.z;
        p->cam.max_x = camera.max_x;
        p->cam.max_y = camera.max_y;
        p->cam.width = cfg.width;
        p->cam.height = cfg.height;
        p->oct_nodes = d_oct_nodes.p;
        p->oct_tris = d_oct_tris.p;
        p->bvh_nodes = cfg.accel == RT_ACCEL_LBVH ? d_lbvh_nodes.p : d_bvh_nodes.p;
        p->bvh_tris = cfg.accel == RT_ACCEL_LBVH ? d_lbvh_tris.p : d_bvh_tris.p;
        p->bvh_top_count = cfg.accel == RT_ACCEL_LBVH ? 0u : (uint32_t)bvh.nodes.size();
        p->bvh_node_lines = (uint32_t)(((cfg.accel == RT_ACCEL_LBVH ? d_lbvh_nodes.n : d_bvh_nodes.n) * sizeof(float4)) / 128u);
        p->bvh_tri_lines = (uint32_t)(((cfg.accel == RT_ACCEL_LBVH ? d_lbvh_tris.n : d_bvh_tris.n) * sizeof(float4)) / 128u);
        p->bvh4_nodes = d_bvh4_nodes.p;
        p->bvh4_tris = d_bvh4_tris.p;
        p->cw_nodes = reinterpret_cast<const uint4*>(d_cw_nodes.p);
        p->cw_tris = d_cw_tris.p;
        p->tri_shade = d_tri_shade.p;
        p->tri_shade_lines = (uint32_t)(d_tri_shade.n * sizeof(float4) / 128u);
        p->materials = d_materials.p;
        p->lights = d_lights.p;
        p->textures = d_textures.p;
        p->num_lights = (uint32_t)scene.lights.size();
        p->film_sum = d_film_sum.p;
        p->film_sq = d_film_sq.p;
        p->ldr = d_ldr.p;
        p->ldr_remote = ldr_remote;
        p->primary_ids = d_ids.p;
        p->counters = d_counters.p;
        p->jitter_mode = (uint32_t)cfg.jitter_mode;
        p->seed = cfg.seed;
        p->recursions = cfg.recursions;
        p->sub_spread = cfg.sub_spread;
        p->sample_table = d_sample_table.p;
        p->magic_w = udiv_magic_of(cfg.width);
        p->magic_h = udiv_magic_of(cfg.height);
        p->magic_tiles_x = udiv_magic_of((cfg.width + 7u) / 8u);
        p->lane_samples_log2 = 0;  // item geometry: set_item_geometry() once n_rows is known
        p->item_cols_log2 = 3;
        p->item_rows = 4;
        p->items_x = (cfg.width + 7u) / 8u;
        p->items_y = 0;
        std::memcpy(p->root_lo, root_lo, 12);
        std::memcpy(p->root_hi, root_hi, 12);
    }

    // The camera moved: the recorded costs no longer describe the view, but they still predict it far better than image order
    // does (a key press moves the camera by a fraction of the scene, main.rs:124-162), so every schedule keeps its order for the
    // next launch, restarts its cost record from zero and re-sorts after that launch and the one after it.
    void invalidate_schedule() {
        drop_lap();  // (the camera moved)
        for (auto& sc : schedules) {
            sc->launches = 0;
            sc->restart_costs = true;
        }
    }
    // The kernel variant / structure / split policy changed: an order written for one kernel must not reach another (the ray-pool
    // kernel takes whole tiles, the persistent kernel takes the split items), and its costs are in different units.
    void reset_schedules() {
        if (!schedules.empty() && !host_only) {  // a launch in flight may still read an order
            cudaSetDevice(device);
            cudaStreamSynchronize(stream);
        }
        schedules.clear();
    }

    static constexpr uint32_t kResortEvery = 32;
    static uint32_t udiv_magic_of(uint32_t d) { return d <= 1u ? 0xffffffffu : (uint32_t)((1ull << 32) / d); }

    // 32 lanes of a warp item = columns x rows x samples (kernels.cu, trace_shade_persistent_kernel)
    static void set_item_geometry(TraceParams* p, uint32_t samples_log2) {
        const uint32_t S = 1u << samples_log2, px = 32u / S;  // pixels per item
        p->lane_samples_log2 = samples_log2;
        p->item_cols_log2 = px >= 8u ? 3u : 2u;
        p->item_rows = px >> p->item_cols_log2;
        p->items_x = (p->cam.width + (1u << p->item_cols_log2) - 1u) >> p->item_cols_log2;
        p->items_y = (p->n_rows + p->item_rows - 1u) / p->item_rows;
        p->magic_tiles_x = udiv_magic_of(p->items_x);
    }

    TileSchedule* find_schedule(const TraceParams& p, uint32_t tiles, int kernel) {
        for (auto& sc : schedules)
            if (sc->first == p.first_row && sc->rows == p.n_rows && sc->tiles == tiles && sc->samples_log2 == p.lane_samples_log2 && sc->kernel == kernel)
                return sc.get();
        if (schedules.size() >= 64) {  // drop the least recently used geometry
            size_t lru = 0;
            for (size_t i = 1; i < schedules.size(); ++i)
                if (schedules[i]->last_use < schedules[lru]->last_use) lru = i;
            RT_CUDA(cudaStreamSynchronize(stream));  // a launch may still read its buffers
            schedules.erase(schedules.begin() + (long)lru);
        }
        auto sc = std::make_unique<TileSchedule>();
        sc->first = p.first_row;
        sc->rows = p.n_rows;
        sc->tiles = tiles;
        sc->samples_log2 = p.lane_samples_log2;
        sc->kernel = kernel;
        // a part must hold whole pixels (32 >> (level + 1) lanes >= the samples of a pixel); big launches fill the GPU with
        // whole tiles and four-way splits, so their order buffer is not sized for the finer levels
        uint32_t lv = (uint32_t)std::max(0, std::min(max_split_level, 3));
        while (lv > 0 && (32u >> (lv + 1u)) < (1u << p.lane_samples_log2)) --lv;
        if (tiles > (1u << 16)) lv = std::min(lv, 1u);
        sc->max_level = lv;
        sc->cost.alloc(tiles);
        sc->order.alloc(((size_t)tiles << (lv ? lv + 1u : 0u)) + 1);
        schedules.push_back(std::move(sc));
        return schedules.back().get();
    }

    // The inverse of Camera::get_ray (camera.rs:80-90) as a matrix: dir = a e0 + b e1 + e2 with a = dir_x, b = -dir_y and e0, e1, e2 = rows 0, 1,
    // 2 + 3 of the rotation matrix, so a point p = origin + t dir has (a t, b t, t) = B^-1 (p - origin), B = [e0 e1 e2] as columns (inverted in
    // binary64). Folded with U = W/2 + a W / (2 max_x), V = H/2 - b H / (2 max_y): (X, Y, Z) = A (p - origin), sample-plane position
    // (U, V) = (X / Z, Y / Z) in pixels, and Z = t. False when the camera matrix cannot be inverted.
    bool camera_plane_matrix(double A[9]) const {
        const float* R = camera.rotation.data();
        const double e[3][3] = {{R[0], R[1], R[2]}, {R[4], R[5], R[6]}, {(double)R[8] + R[12], (double)R[9] + R[13], (double)R[10] + R[14]}};
        // rows of the inverse are cross products / det
        const double c0[3] = {e[1][1] * e[2][2] - e[1][2] * e[2][1], e[1][2] * e[2][0] - e[1][0] * e[2][2], e[1][0] * e[2][1] - e[1][1] * e[2][0]};
        const double c1[3] = {e[2][1] * e[0][2] - e[2][2] * e[0][1], e[2][2] * e[0][0] - e[2][0] * e[0][2], e[2][0] * e[0][1] - e[2][1] * e[0][0]};
        const double c2[3] = {e[0][1] * e[1][2] - e[0][2] * e[1][1], e[0][2] * e[1][0] - e[0][0] * e[1][2], e[0][0] * e[1][1] - e[0][1] * e[1][0]};
        const double det = e[0][0] * c0[0] + e[0][1] * c0[1] + e[0][2] * c0[2];
        const double scale = std::fabs(e[0][0]) + std::fabs(e[0][1]) + std::fabs(e[0][2]) + std::fabs(e[1][0]) + std::fabs(e[1][1]) + std::fabs(e[1][2]) +
                             std::fabs(e[2][0]) + std::fabs(e[2][1]) + std::fabs(e[2][2]);
        if (!(std::fabs(det) > 1e-9 * scale * scale * scale) || !(camera.max_x > 0.f) || !(camera.max_y > 0.f)) return false;
        const double W = cfg.width, H = cfg.height, sx = W / (2.0 * camera.max_x), sy = H / (2.0 * camera.max_y);
        for (int k = 0; k < 3; ++k) {
            const double i0 = c0[k] / det, i1 = c1[k] / det, i2 = c2[k] / det;
            A[k] = sx * i0 + 0.5 * W * i2;
            A[3 + k] = -sy * i1 + 0.5 * H * i2;  // (dir_y = -b)
            A[6 + k] = i2;
        }
        return true;
    }

    // (Re)builds the camera grid for the current view over `tris` and fills p's grid fields; false = no grid for this launch (switched off,
    // or the camera matrix cannot be inverted): the camera rays walk the BVH.
    bool ensure_pgrid(TraceParams* p) {
        if (camera_grid_log2 <= 0 || !p->bvh_tris) return false;
        PGridKey key;
        std::memcpy(key.cam, camera.rotation.data(), 64);
        const f3 o = camera.ray_origin();
        key.cam[16] = o.x, key.cam[17] = o.y, key.cam[18] = o.z, key.cam[19] = camera.max_x, key.cam[20] = camera.max_y;
        key.w = cfg.width, key.h = cfg.height, key.shift = (uint32_t)camera_grid_log2;
        key.tris = p->bvh_tris;
        key.n_slots = (uint32_t)((cfg.accel == RT_ACCEL_LBVH ? d_lbvh_tris.n : d_bvh_tris.n) / 3);
        if (!pg_valid || std::memcmp(&key, &pg_key, sizeof(key)) != 0) {
            pg_valid = false;
            // A build costs about as much as a frame (five small kernels and one 4-byte readback the host waits for): a view is given its grid
            // once it has been launched `camera_grid_after` times without changing (a camera that moves every frame walks the BVH, as before;
            // a view that stays — the reference accumulates samples until the next key press — gets the grid from its second frame on)
            if (std::memcmp(&key, &pg_seen, sizeof(key)) != 0) {
                pg_seen = key;
                pg_seen_launches = 0;
            }
            if (pg_seen_launches++ < (uint32_t)camera_grid_after) return false;
            PGridParams g{};
            if (!camera_plane_matrix(g.A[0])) return false;
            g.origin[0] = o.x, g.origin[1] = o.y, g.origin[2] = o.z;
            double extent = 0.0;
            for (int a = 0; a < 3; ++a) extent = std::max(extent, (double)root_hi[a] - (double)root_lo[a]);
            g.z_eps = std::max(1e-9 * extent * std::sqrt(g.A[0][6] * g.A[0][6] + g.A[0][7] * g.A[0][7] + g.A[0][8] * g.A[0][8]), 1e-30);  // Z is in units of |row 2 of the inverse|
            uint32_t sh = key.shift;
            const uint32_t pv_max = (uint32_t)(((uint64_t)cfg.width * cfg.height - 1u) / cfg.height);  // v = idx / height (mod.rs:96)
            while ((uint64_t)(((cfg.width - 1u) >> sh) + 1u) * ((pv_max >> sh) + 1u) >= (1u << 20)) ++sh;  // (the scan handles < 2^20 cells)
            g.nx = ((cfg.width - 1u) >> sh) + 1u;
            g.ny = (pv_max >> sh) + 1u;
            g.cell = (double)(1u << sh);
            const uint32_t n_cells = g.nx * g.ny;
            if (d_pg_count.n < n_cells) {
                RT_CUDA(cudaStreamSynchronize(stream));  // a launch in flight may still read the old arrays
                d_pg_count.alloc(n_cells);
                d_pg_start.alloc((size_t)n_cells + 1);
                d_pg_cursor.alloc(n_cells);
            }
            if (!d_pg_total.p) d_pg_total.alloc(1 + 1024);  // entry count, then the scan's block sums
            if (d_pg_big.n < (size_t)key.n_slots + 1) d_pg_big.alloc((size_t)key.n_slots + 1);  // count, then the queue of big footprints
            g.tris = key.tris;
            g.n_slots = key.n_slots;
            g.count = d_pg_count.p;
            g.cursor = d_pg_cursor.p;
            g.entries = d_pg_entries.p;
            g.n_frusta = 1;
            g.cell_base = 0;
            g.big_count = d_pg_big.p;
            g.big_queue = d_pg_big.p + 1;
            g.key_mode = 0;
            g.dmin2 = nullptr;
            RT_CUDA(cudaMemsetAsync(d_pg_count.p, 0, (size_t)n_cells * 4, stream));
            RT_CUDA(pgrid_bin_count(g, num_sms, stream));
            RT_CUDA(pgrid_scan(d_pg_count.p, d_pg_start.p, d_pg_cursor.p, n_cells, d_pg_total.p, d_pg_total.p + 1, stream));
            uint32_t total = 0;
            RT_CUDA(cudaMemcpyAsync(&total, d_pg_total.p, 4, cudaMemcpyDeviceToHost, stream));
            RT_CUDA(cudaStreamSynchronize(stream));  // (also: no launch still reads the old lists)
            if (total > kGridMaxEntries) {  // a view from inside a dense mesh can list every triangle in most cells: the tree then serves better
                pg_seen_launches = 0;       // (asked again after `camera_grid_after` more launches of this view; a new view asks at once)
                return false;
            }
            if (d_pg_entries.n < total) d_pg_entries.alloc((size_t)total + total / 2 + 1024);
            g.entries = d_pg_entries.p;
            RT_CUDA(pgrid_bin_fill(g, num_sms, stream));
            RT_CUDA(pgrid_sort_lists(d_pg_start.p, d_pg_entries.p, n_cells, stream));
            total_kernels += 8;
            last.kernels_launched += 8;
            pg_key = key;
            pg_nx = g.nx;
            pg_cells = n_cells;
            pg_shift = sh;
            pg_entries = total;
            pg_valid = true;
            ++pg_builds;
        }
        p->pg_start = d_pg_start.p;
        p->pg_tris = d_pg_entries.p;
        p->grid_lines[0] = (uint32_t)(((size_t)pg_cells + 1) * 4 / 128);
        p->grid_lines[1] = (uint32_t)((size_t)pg_entries * 8 / 128);
        p->pg_nx = pg_nx;
        p->pg_shift = pg_shift;
        return true;
    }

    // Shadow-ray grids (pgrid_build.cu with six frusta per light); lights never move, so this runs once per triangle array.
    void ensure_lgrid(TraceParams* p) {
        const uint32_t n_lights = (uint32_t)std::min<size_t>(scene.lights.size(), (size_t)kGridLights);
        if (light_grid_log2 <= 0 || n_lights == 0 || !p->bvh_tris) return;
        if (lg_off && lg_tris_key == p->bvh_tris && lg_log2_key == light_grid_log2) return;  // measured too large for this tree and resolution
        const uint32_t n_slots = (uint32_t)((cfg.accel == RT_ACCEL_LBVH ? d_lbvh_tris.n : d_bvh_tris.n) / 3);
        if (n_slots < (uint32_t)light_grid_min_tris) return;  // a tree of a few dozen nodes is walked faster than a list is read (4boxes: 0.057 against 0.061 ms)
        if (!lg_valid || lg_tris_key != p->bvh_tris || lg_slots_key != n_slots || lg_log2_key != light_grid_log2) {
            lg_valid = false;
            lg_off = false;
            uint32_t n = 1u << light_grid_log2;
            while (n > 32u && (uint64_t)n_lights * 6u * n * n >= (1u << 20)) n >>= 1;  // (the scan handles < 2^20 cells)
            const uint32_t shift = 3u, face_cells = n * n, n_cells = n_lights * 6u * face_cells;
            RT_CUDA(cudaStreamSynchronize(stream));  // a launch in flight may still read the old arrays
            d_lg_count.alloc(n_cells);
            d_lg_start.alloc((size_t)n_cells + 1);
            d_lg_cursor.alloc(n_cells);
            if (!d_lg_total.p) d_lg_total.alloc(1 + 1024);
            if (!d_lg_dmin2.p) d_lg_dmin2.alloc(kGridLights);
            const size_t big_words = (size_t)n_slots * 6 + 1;  // per light: count, then the queue
            d_lg_big.alloc(big_words * n_lights);
            RT_CUDA(cudaMemsetAsync(d_lg_count.p, 0, (size_t)n_cells * 4, stream));
            RT_CUDA(cudaMemsetAsync(d_lg_dmin2.p, 0x7f, kGridLights * sizeof(float), stream));  // 3.39e38
            double extent = 0.0;
            for (int a = 0; a < 3; ++a) extent = std::max(extent, (double)root_hi[a] - (double)root_lo[a]);
            PGridParams g[kGridLights];
            for (uint32_t li = 0; li < n_lights; ++li) {
                PGridParams& q = g[li];
                q = PGridParams{};
                q.tris = p->bvh_tris;
                q.n_slots = n_slots;
                q.n_frusta = 6;
                const double half = 0.5 * (double)(n << shift);
                for (int m = 0; m < 3; ++m)
                    for (int neg = 0; neg < 2; ++neg) {
                        double* A = q.A[2 * m + neg];
                        const double sgn = neg ? -1.0 : 1.0;
                        for (int k = 0; k < 9; ++k) A[k] = 0.0;
                        A[0 + (m + 1) % 3] = half, A[0 + m] = half * sgn;  // X = half (w[m+1] + s w[m])
                        A[3 + (m + 2) % 3] = half, A[3 + m] = half * sgn;  // Y = half (w[m+2] + s w[m])
                        A[6 + m] = sgn;                                    // Z = s w[m]
                    }
                q.origin[0] = scene.lights[li].pos[0], q.origin[1] = scene.lights[li].pos[1], q.origin[2] = scene.lights[li].pos[2];
                q.z_eps = std::max(1e-9 * extent, 1e-30);
                q.nx = q.ny = n;
                q.cell = (double)(1u << shift);
                q.cell_base = li * 6u * face_cells;
                q.count = d_lg_count.p;
                q.cursor = d_lg_cursor.p;
                q.entries = nullptr;
                q.dmin2 = d_lg_dmin2.p + li;
                q.big_count = d_lg_big.p + big_words * li;
                q.big_queue = q.big_count + 1;
                q.key_mode = 1;
                RT_CUDA(pgrid_bin_count(q, num_sms, stream));
            }
            RT_CUDA(pgrid_scan(d_lg_count.p, d_lg_start.p, d_lg_cursor.p, n_cells, d_lg_total.p, d_lg_total.p + 1, stream));
            uint32_t total = 0;
            float dmin2[kGridLights];
            RT_CUDA(cudaMemcpyAsync(&total, d_lg_total.p, 4, cudaMemcpyDeviceToHost, stream));
            RT_CUDA(cudaMemcpyAsync(dmin2, d_lg_dmin2.p, sizeof(dmin2), cudaMemcpyDeviceToHost, stream));
            RT_CUDA(cudaStreamSynchronize(stream));
            if (total > kGridMaxEntries) {  // a light buried in a dense mesh: shadow rays keep walking the tree
                lg_off = true;
                lg_tris_key = p->bvh_tris;
                lg_log2_key = light_grid_log2;
                return;
            }
            if (d_lg_entries.n < total) d_lg_entries.alloc((size_t)total + 1024);
            for (uint32_t li = 0; li < n_lights; ++li) {
                g[li].entries = d_lg_entries.p;
                RT_CUDA(pgrid_bin_fill(g[li], num_sms, stream));
                // the ray ends 0.01 |L| past the light: it stays short of every surface while 0.01 |L| < d_min, i.e. |L|^2 < 1e4 d_min^2
                // (2 % kept in hand for the rounding of the f32 ray and of the bound)
                lg_far2[li] = dmin2[li] * 1e4f * 0.98f;
            }
            RT_CUDA(pgrid_sort_lists(d_lg_start.p, d_lg_entries.p, n_cells, stream));
            total_kernels += 4 * n_lights + 4;
            last.kernels_launched += 4 * n_lights + 4;
            lg_tris_key = p->bvh_tris;
            lg_slots_key = n_slots;
            lg_log2_key = light_grid_log2;
            lg_n = n;
            lg_cells = n_cells;
            lg_entries = total;
            lg_valid = true;
        }
        p->lg_start = d_lg_start.p;
        p->lg_tris = d_lg_entries.p;
        p->grid_lines[2] = (uint32_t)(((size_t)lg_cells + 1) * 4 / 128);
        p->grid_lines[3] = (uint32_t)((size_t)lg_entries * 8 / 128);
        p->lg_n = lg_n;
        p->lg_shift = 3u;
        p->lg_half = 0.5f * (float)(lg_n << 3u);
        for (int li = 0; li < kGridLights; ++li) p->lg_far2[li] = lg_far2[li];
    }

    cudaError_t launch_one(const TraceParams& p_in) {
        TraceParams p = p_in;
        set_item_geometry(&p, p.lane_samples_log2);
        const int a = cfg.accel == RT_ACCEL_OCTREE ? 0 : (cfg.accel == RT_ACCEL_CWBVH ? 2 : (cfg.accel == RT_ACCEL_BVH4 ? 3 : 1));  // LBVH: same traversal as the SAH binary BVH
        const int b = cfg.recursions > 0 ? 1 : 0;
        // the ray-pool kernel covers the headline configuration; everything else runs the persistent tile kernel
        const bool use_pool = variant == 2 && a == 1 && b == 0 && scene.lights.size() == 1 && !p.planes && p.lane_samples_log2 == 0;
        if (use_pool && pool_blocks == 0) pool_blocks = pool_blocks_per_sm();
        // kernel instantiation: the binary-BVH kernels exist a second time with the camera rays sent through the perspective grid
        int ka = a;
        const uint64_t pg_builds_before = pg_builds;
        if (a == 1 && variant == 1 && !use_pool) {
            // the grids are an accelerator of the accelerator: a build that fails (device memory) switches them off, the launch walks the tree
            try {
                if (ensure_pgrid(&p)) ka = 4;
            } catch (CudaFail&) {
                cudaGetLastError();
                camera_grid_log2 = 0;
                pg_valid = false;
                ka = a;
            }
            if (ka == 4) {
                try {
                    ensure_lgrid(&p);
                } catch (CudaFail&) {
                    cudaGetLastError();
                    light_grid_log2 = 0;
                    lg_valid = false;
                    p.lg_start = nullptr;
                    p.lg_tris = nullptr;
                    p.grid_lines[2] = p.grid_lines[3] = 0u;
                }
            }
        }
        if (!use_pool && variant != 0 && blocks_per_sm[ka][b] == 0) blocks_per_sm[ka][b] = persistent_blocks_per_sm(ka, b);
        p.queue_batch = (uint32_t)queue_batch;
        p.queue_batch_from_pct = (uint32_t)queue_batch_from_pct;
        p.pool_refill = (uint32_t)pool_refill;
        p.pool_min_inner = (uint32_t)pool_min_inner;
        p.counter_set = call_parity ? CNT_SET_B : CNT_SET_A;
        p.film_prefetch = (uint32_t)film_prefetch;
        p.film_prefetch_rows = film_prefetch >= 3 && (size_t)p.n_rows * cfg.width * 16u <= ((size_t)film_prefetch_rows_mb << 20) ? 1u : 0u;
        // the frame-done signal of a multi-GPU run rides on the last launch of the call when that launch is the one that finishes
        // the pixels (persistent / ray-pool kernel writing the film itself); otherwise trace_rows appends a signal launch
        const bool finishes_pixels = variant != 0 && !p.planes && !wavefront_applies(p);
        p.done_flag = (arm_done && finishes_pixels) ? done_flag : nullptr;
        p.done_value = done_value;
        if (p.done_flag) done_published = true;
        const uint32_t tiles = p.items_x * p.items_y;
        if (variant != 0 && lpt_schedule && tiles >= (uint32_t)min_schedule_tiles) {  // tiny launches are latency bound: image order
            TileSchedule* sc = nullptr;
            try {
                sc = find_schedule(p, tiles, ka);
            } catch (CudaFail&) {
                return cudaErrorMemoryAllocation;
            }
            sc->last_use = ++schedule_clock;
            // A schedule that is only ever used for the first frame of a view (the tree-walking kernel, while views that stay get the grid
            // kernel) is invalidated before it reaches its first sort: the costs of its last launch describe an earlier view, which still
            // predicts this one far better than image order does — sort them before they are zeroed.
            const bool late_first_sort = sc->restart_costs && !sc->have_order && sc->recorded > 0;
            if (sc->restart_costs && !late_first_sort) {
                RT_CUDA_RET(cudaMemsetAsync(sc->cost.p, 0, tiles * sizeof(uint32_t), stream));
                sc->restart_costs = false;
            }
            // re-sort after the 1st and 2nd recorded launch of a view, then every 32nd (the one-block sort costs
            // ~45 us for a 1080p frame; per-tile costs of an unchanged view move little between frames)
            if (late_first_sort || sc->launches == 1 || sc->launches == 2 || (sc->launches > 2 && sc->launches % kResortEvery == 0)) {
                const uint32_t warps = (uint32_t)((use_pool ? pool_blocks : blocks_per_sm[ka][b]) * num_sms * 8);
                const uint32_t level = (a != 0 && !use_pool && split_quarters > 0) ? sc->max_level : 0u;
                // a launch that cannot fill the resident warps even once has issue slots to spare: split from 5 us of work on
                const uint32_t min_cycles = tiles < warps ? 10000u : 40000u;
                cudaError_t e = launch_tile_sort(sc->cost.p, sc->order.p, tiles, warps, (uint32_t)split_quarters, level, min_cycles,
                                                 sc->order.p + (sc->order.n - 1), stream);
                if (e != cudaSuccess) return e;
                ++total_kernels;
                ++last.kernels_launched;
                sc->have_order = true;
                if (late_first_sort) {
                    RT_CUDA_RET(cudaMemsetAsync(sc->cost.p, 0, tiles * sizeof(uint32_t), stream));
                    sc->restart_costs = false;
                }
            }
            if (use_pool) {
                // the pool kernel ADDS every camera ray's steps to its tile's cost: record only the launches a sort
                // will read (the one before each re-sort), starting from zero
                const bool record = sc->launches <= 1 || sc->launches % kResortEvery == kResortEvery - 1;
                if (record && sc->launches > 0) RT_CUDA_RET(cudaMemsetAsync(sc->cost.p, 0, tiles * sizeof(uint32_t), stream));
                p.tile_cost = record ? sc->cost.p : nullptr;
            } else {
                p.tile_cost = sc->cost.p;
            }
            p.tile_order = sc->have_order ? sc->order.p : nullptr;
            p.queue_items = sc->order.p + (sc->order.n - 1);
            ++sc->launches;
            ++sc->recorded;
            // The view just got its grid, so it stays: its first frame walked the tree and left that kernel's schedule one recorded launch that
            // nothing would ever sort (the tree-walking kernel runs again only after the next camera move, which restarts the record). Sort it
            // now, on this frame, which pays for the build anyway: the first frame after the NEXT key press then runs with an order that is one
            // view old — what it had before the grids existed — instead of the order of the handle's very first view.
            if (ka == 4 && pg_builds != pg_builds_before && blocks_per_sm[1][b] > 0) {
                for (auto& s1 : schedules)
                    if (s1->kernel == 1 && s1->first == sc->first && s1->rows == sc->rows && s1->tiles == tiles && s1->samples_log2 == sc->samples_log2 &&
                        s1->launches >= 1 && !s1->restart_costs) {
                        const uint32_t warps1 = (uint32_t)(blocks_per_sm[1][b] * num_sms * 8);
                        const uint32_t level1 = split_quarters > 0 ? s1->max_level : 0u;
                        cudaError_t e1 = launch_tile_sort(s1->cost.p, s1->order.p, tiles, warps1, (uint32_t)split_quarters, level1, tiles < warps1 ? 10000u : 40000u,
                                                          s1->order.p + (s1->order.n - 1), stream);
                        if (e1 != cudaSuccess) return e1;
                        ++total_kernels;
                        ++last.kernels_launched;
                        s1->have_order = true;
                    }
            }
        }
        cudaError_t e;
        if (use_pool) e = launch_trace(p, a, 2, pool_blocks * num_sms, stream);
        else if (wavefront_applies(p)) e = launch_wavefront(p, a, ka);
        else e = launch_trace(p, ka, variant == 0 ? 0 : 1, blocks_per_sm[ka][b] * num_sms, stream);
        if (e != cudaSuccess) return e;
        if (variant == 0) {
            // the one-thread-per-pixel kernel has no last warp out (warp_checkout): the host zeroes the other per-call counter set
            const int other = call_parity ? CNT_SET_A : CNT_SET_B;
            e = cudaMemsetAsync(d_counters.p + other, 0, 4 * sizeof(unsigned long long), stream);
        }
        set_dirty[call_parity ^ 1] = false;
        return e;
    }

    // ---- bounce wavefront (kernels.cu: trace_pixel BOUNCE = 2, wf_bounce_kernel, wf_combine_kernel) ----
    static constexpr size_t kWfMaxBytes = size_t(6) << 30;
    // nodes of level l are at most pixels * prod_{i<l} n_i with n_i = sub_spread * (recursions - i) bounce rays per hit
    bool wf_layout(const TraceParams& p, size_t cap[kWfLevels], uint32_t nch[kWfLevels], size_t* rec_total, size_t* child_total) const {
        const int R = cfg.recursions;
        if (R < 1 || R > kWfLevels - 1) return false;
        size_t c = (size_t)p.n_rows * p.cam.width, recs = 0, childs = 0;
        for (int l = 0; l <= R; ++l) {
            cap[l] = c;
            nch[l] = cfg.sub_spread * (uint32_t)(R - l);
            if (cap[l] >= (size_t(1) << 28)) return false;  // parent index has 28 bits
            recs += kWfRecWords * cap[l];
            childs += 3 * cap[l] * nch[l];
            if (nch[l] > 15u) return false;  // child number has 4 bits
            c *= std::max<size_t>(nch[l], 1);
            if (recs * 16 + childs * 4 > kWfMaxBytes) return false;
        }
        *rec_total = recs;
        *child_total = childs;
        return true;
    }
    bool wavefront_applies(const TraceParams& p) const {
        if (!bounce_wavefront || cfg.recursions < 1 || variant == 0 || p.planes || cfg.sub_spread == 0) return false;
        size_t cap[kWfLevels], rt_, ct_;
        uint32_t nch[kWfLevels];
        return wf_layout(p, cap, nch, &rt_, &ct_);
    }
    cudaError_t launch_wavefront(const TraceParams& p_in, int a, int ka /* instantiation of the trace kernel: a, or 4 = camera grid */) {
        TraceParams p = p_in;
        size_t cap[kWfLevels], rec_total = 0, child_total = 0;
        uint32_t nch[kWfLevels];
        wf_layout(p, cap, nch, &rec_total, &child_total);
        if (d_wf_rec.n < rec_total) d_wf_rec.alloc(rec_total);
        if (d_wf_child.n < std::max<size_t>(child_total, 1)) d_wf_child.alloc(std::max<size_t>(child_total, 1));
        if (!d_wf_counts.p) d_wf_counts.alloc(3 * kWfLevels);  // nodes per level | ray-queue head | shade-queue head
        const int R = cfg.recursions;
        size_t ro = 0, co = 0;
        for (int l = 0; l <= R; ++l) {
            p.wf[l].rec = d_wf_rec.p + ro;
            p.wf[l].child_r = d_wf_child.p + co;
            p.wf[l].cap = (uint32_t)cap[l];
            p.wf[l].n_children = nch[l];
            ro += kWfRecWords * cap[l];
            co += 3 * cap[l] * nch[l];
        }
        p.wf_counts = d_wf_counts.p;
        RT_CUDA_RET(cudaMemsetAsync(d_wf_counts.p, 0, 3 * kWfLevels * sizeof(unsigned int), stream));
        if (blocks_per_sm[ka][0] == 0) blocks_per_sm[ka][0] = persistent_blocks_per_sm(ka, 0);
        RT_CUDA_RET(launch_trace(p, ka, 1, blocks_per_sm[ka][0] * num_sms, stream));
        const int wf_blocks = num_sms * 3;  // 80 registers: three 256-thread blocks per SM
        const bool stream_levels = bounce_stream && a == 1;  // ray-stream kernel (binary BVH): bounce + shadow rays share the lanes
        // every level from the second on sends at most one bounce ray per hit (the reference's RECURSIONS = 2, SUB_SPREAD = 1): the ray-stream
        // kernel chains them in place and ONE launch walks the whole bounce tree
        bool chain = stream_levels && stream_chain;
        for (int l = 1; l <= R; ++l) chain = chain && nch[l] <= 1;
        p.wf_chain = chain ? 1u : 0u;
        int stream_launches = 0;
        for (int l = 0; l < (chain ? 1 : R); ++l) {  // level l -> l + 1
            p.wf_level = (uint32_t)l;
            ++stream_launches;
            if (stream_levels) {
                p.pool_refill = (uint32_t)stream_refill;
                p.pool_min_inner = (uint32_t)stream_min_inner;
                RT_CUDA_RET(launch_wf_stream(p, stream_blocks, num_sms, stream));
            } else {  // lockstep form: trace the bounce rays, then shade the compacted hits
                // these kernels need far fewer registers than the trace kernel: as many blocks as fit (the rays are latency bound)
                if (wf_occupancy[a][0] == 0) {
                    wf_occupancy[a][0] = wf_blocks_per_sm(0, a);
                    wf_occupancy[a][1] = wf_blocks_per_sm(1, a);
                }
                const int b0 = wf_blocks_cap > 0 ? std::min(wf_blocks_cap, wf_occupancy[a][0]) : wf_occupancy[a][0];
                const int b1 = wf_blocks_cap > 0 ? std::min(wf_blocks_cap, wf_occupancy[a][1]) : wf_occupancy[a][1];
                RT_CUDA_RET(launch_wf_bounce(p, a, num_sms * b0, stream));
                p.wf_level = (uint32_t)(l + 1);
                RT_CUDA_RET(launch_wf_shade(p, a, num_sms * b1, stream));
            }
        }
        for (int l = R; l >= 0; --l) {  // bottom up; level 0 adds the radiance to the film
            p.wf_level = (uint32_t)l;
            RT_CUDA_RET(launch_wf_combine(p, wf_blocks, stream));
        }
        const int wf_kernels = (stream_levels ? stream_launches : 2 * R) + R + 1;
        total_kernels += (uint64_t)wf_kernels;
        last.kernels_launched += (uint32_t)wf_kernels;
        return cudaSuccess;
    }

    // ---- the band loop, traced a lap ahead ------------------------------------------------------------------------------------
    // The reference's render loop asks for 50 rows at a time (trace_frame_additive, mod.rs:87; main.rs:200). A launch over 50 rows
    // cannot fill a B200: at 1080p it holds 3 120 tiles for 3 552 resident warps, every warp starts with a cold L1, and the launch
    // lasts as long as its slowest chain of dependent node fetches from L2 — 20 to 130 us per band, 1.5 ms for the 22 bands of a frame
    // that one full-frame launch traces in 0.15 ms. So the first band call of a lap traces the NEXT SAMPLE OF EVERY ROW from its first
    // row to the bottom of the image in ONE launch, into a frame-aligned sample plane (radiance + primitive id per pixel, ray counts per
    // pixel), and every band call — this one included — COMMITS its rows from the plane: PixelData::add_sample, mean, tonemap, pack, the
    // call's ray counts. A sample's number is the film count of its pixel, which only a commit changes, so the committed samples are
    // exactly the ones band-by-band tracing would have produced: same film, ids, frame and per-call counters after every call.
    // Anything that makes the plane stale drops it (camera, film, configuration, an explicit rt_trace_rows); the rows not yet
    // committed are then simply traced again when their band comes up.
    DevBuf<float4> d_lap;
    DevBuf<uint16_t> d_lap_rays;
    bool lap_valid = false;
    uint32_t lap_next = 0;   // rows [lap_next, height) of the plane are traced and not yet committed
    int band_lookahead = 1;  // RT_TUNE_BAND_LOOKAHEAD
    int film_prefetch = 3;   // RT_TUNE_FILM_PREFETCH
    int film_prefetch_rows_mb = 48;  // RT_TUNE_FILM_PREFETCH_ROWS_MB
    uint64_t lap_builds = 0, lap_drops = 0;
    void drop_lap() {
        if (lap_valid) ++lap_drops;
        lap_valid = false;
    }
    bool lookahead_applies(uint32_t rows) const {
        return band_lookahead && !sharded() && variant == 1 && rows < cfg.height && cfg.recursions <= 4 && scene.lights.size() < 255;
    }
    void lap_build(uint32_t from) {
        const uint32_t rows = cfg.height - from, padded = (rows + 3u) & ~3u;
        if (!d_lap.p) {
            d_lap.alloc(npix());
            d_lap_rays.alloc(npix());
        }
        TraceParams q;
        fill_params(&q);
        q.first_row = from;
        q.planes = d_lap.p + (size_t)from * cfg.width;  // compact row c of the launch = image row from + c: the plane is frame aligned
        q.n_planes = 1;
        q.plane_rows = rows;
        q.plane_rows_padded = padded;
        q.magic_plane_rows = udiv_magic_of(padded);
        q.n_rows = padded;
        q.lap_rays = d_lap_rays.p;
        RT_CUDA(launch_one(q));
        ++total_kernels;
        ++last.kernels_launched;
        ++lap_builds;
        lap_valid = true;
        lap_next = from;
    }
    void lap_commit(uint32_t r0, uint32_t r1) {
        TraceParams q;
        fill_params(&q);
        q.first_row = r0;
        q.planes = d_lap.p + (size_t)r0 * cfg.width;
        q.n_planes = 1;
        q.plane_rows = r1 - r0;
        q.plane_rows_padded = r1 - r0;
        q.lap_rays = d_lap_rays.p;
        q.counter_set = call_parity ? CNT_SET_B : CNT_SET_A;
        RT_CUDA(launch_film_accumulate(q, stream));
        ++total_kernels;
        ++last.kernels_launched;
        lap_next = r1;
        if (lap_next >= cfg.height) lap_valid = false;  // used up (not a drop)
    }
    // trace_frame_additive through the lap plane: rows [first, first + rows) modulo height, one sample
    void trace_band_lookahead(uint32_t first, uint32_t rows) {
        if (cfg.recursions > 0) ensure_sample_table();
        ensure_accel(cfg.accel);
        begin_call_counters();  // (most band calls are pure commits: their counter set is zeroed by a 32-byte memset)
        mark_rows_dirty(first, rows);
        last = rt_launch_stats{};
        last_timed = time_launches != 0;
        if (last_timed) RT_CUDA(cudaEventRecord(ev_start, stream));
        uint32_t r = first % cfg.height, remaining = rows;
        while (remaining) {
            if (!lap_valid || lap_next != r) {
                drop_lap();
                lap_build(r);
            }
            const uint32_t take = std::min(remaining, cfg.height - r);
            lap_commit(r, r + take);
            r = (r + take) % cfg.height;
            remaining -= take;
        }
        if (last_timed) RT_CUDA(cudaEventRecord(ev_stop, stream));
        last.n_primary = (uint64_t)rows * cfg.width;
        total_primary += last.n_primary;
        stats_pending = true;
    }

    // rows [first_row, first_row + n_rows) modulo height, `spp` passes
    void trace_rows(uint32_t first_row, uint32_t n_rows, uint32_t spp) {
        drop_lap();  // the film counts of these rows move: samples traced ahead for them would no longer be the next ones
        if (cfg.recursions > 4) throw CudaFail{"recursions > 4 is not supported (the reference uses 2)"};
        if (cfg.recursions > 0) ensure_sample_table();
        ensure_accel(cfg.accel);
        TraceParams p;
        fill_params(&p);
        uint32_t launch_rows = n_rows;
        first_row %= cfg.height;
        const bool wraps_twice = n_rows > cfg.height;
        if (sharded() || wraps_twice) {
            if (cached_first != first_row || cached_n != n_rows) {
                row_list_cache.clear();
                lap_start.clear();
                for (uint32_t k = 0; k < n_rows; ++k) {
                    if (k % cfg.height == 0) lap_start.push_back((uint32_t)row_list_cache.size());  // a new lap over the image begins
                    const uint32_t r = (first_row + k) % cfg.height;
                    if (owns_row(r)) row_list_cache.push_back(r);
                }
                d_row_list.upload(row_list_cache, stream);
                cached_first = first_row;
                cached_n = n_rows;
            }
            p.row_list = d_row_list.p;
            launch_rows = (uint32_t)row_list_cache.size();
        }
        p.first_row = first_row;
        // no memset: this call counts into the counter set the previous call's last warp out left zeroed (kernels.cu, warp_checkout)
        begin_call_counters();
        mark_rows_dirty(first_row, n_rows);
        last = rt_launch_stats{};
        last_timed = time_launches != 0;
        if (last_timed) RT_CUDA(cudaEventRecord(ev_start, stream));
        uint32_t launches = 0;
        if (wraps_twice) {
            // the same pixel appears more than once: keep the reference's sequential order, one launch per lap (a launch
            // must never hold a pixel twice — two warps would update its film record concurrently); a sharded handle's laps
            // are the owned rows of each lap
            for (uint32_t s = 0; s < spp; ++s)
                for (size_t lap = 0; lap < lap_start.size(); ++lap) {
                    const uint32_t off = lap_start[lap];
                    const uint32_t end = lap + 1 < lap_start.size() ? lap_start[lap + 1] : launch_rows;
                    if (end == off) continue;
                    TraceParams q = p;
                    q.row_list = d_row_list.p + off;
                    q.n_rows = end - off;
                    RT_CUDA(launch_one(q));
                    ++launches;
                }
        } else if (spp > 1 && multi_sample_launch == 1 && variant == 1 && (spp & 1u) == 0u && cfg.recursions == 0) {
            // sample lanes: the 32 lanes of a warp item hold S = 8, 4 or 2 samples of 4, 8 or 16 pixels and add them to the
            // film in sample order themselves (finish_sample_lanes): spp / S launches, no sample planes, no second pass
            const uint32_t s_log2 = (spp & 7u) == 0u ? 3u : ((spp & 3u) == 0u ? 2u : 1u);
            p.n_rows = launch_rows;
            p.lane_samples_log2 = s_log2;
            for (uint32_t s = 0; s < spp; s += 1u << s_log2) {
                arm_done = done_flag && s + (1u << s_log2) >= spp;
                RT_CUDA(launch_one(p));
                ++launches;
            }
        } else if (spp > 1 && multi_sample_launch && !(cfg.recursions > 0 && bounce_wavefront && variant != 0)) {
            // all samples of a pass in ONE launch (sample planes) + one ordered accumulation: same film as `spp`
            // consecutive launches, but the GPU sees spp times as many work items (matters for small row ranges)
            const uint32_t padded = (launch_rows + 3u) & ~3u;
            const size_t plane_px = (size_t)padded * cfg.width;
            const uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(spp, (size_t(1) << 30) / std::max<size_t>(plane_px * sizeof(float4), 1)));
            if (d_planes.n < plane_px * chunk) d_planes.alloc(plane_px * chunk);
            for (uint32_t s0 = 0; s0 < spp; s0 += chunk) {
                TraceParams q = p;
                q.planes = d_planes.p;
                q.n_planes = std::min(chunk, spp - s0);
                q.plane_rows = launch_rows;
                q.plane_rows_padded = padded;
                q.magic_plane_rows = udiv_magic_of(padded);
                q.n_rows = padded * q.n_planes;
                RT_CUDA(launch_one(q));
                RT_CUDA(launch_film_accumulate(q, stream));  // adds the planes to the film in sample order
                launches += 2;
            }
        } else {
            p.n_rows = launch_rows;
            for (uint32_t s = 0; s < spp; ++s) {
                arm_done = done_flag && s + 1u == spp;
                RT_CUDA(launch_one(p));
                ++launches;
            }
        }
        arm_done = false;
        if (done_flag) {  // one-shot: rt_set_done_signal arms it for one trace call
            if (!done_published) {
                RT_CUDA(launch_flag_signal(done_flag, done_value, stream));
                ++launches;
            }
            done_flag = nullptr;
            done_published = false;
        }
        if (last_timed) RT_CUDA(cudaEventRecord(ev_stop, stream));
        // the ray counters are fetched only if somebody asks for them before the next trace call (finish_stats)
        total_kernels += launches;
        last.kernels_launched += launches;
        last.n_primary = (uint64_t)launch_rows * cfg.width * spp;
        total_primary += last.n_primary;
        stats_pending = true;
    }

    void finish_stats() {
        if (!stats_pending) return;
        RT_CUDA(cudaMemcpyAsync(h_counters, d_counters.p, CNT_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        RT_CUDA(cudaStreamSynchronize(stream));
        float ms = 0.f;
        if (last_timed) RT_CUDA(cudaEventElapsedTime(&ms, ev_start, ev_stop));
        last.trace_kernel_ms = ms;
        const int set = call_parity ? CNT_SET_B : CNT_SET_A;
        last.n_shadow = h_counters[set + CNT_SHADOW];
        last.n_bounce = h_counters[set + CNT_BOUNCE];
        stats_pending = false;
    }
};

// =====================================================================================================
// C ABI
// =====================================================================================================
#define RT_GUARD(rt, ...)                      \
    if (!(rt)) return RT_ERR_INVALID;          \
    try {                                      \
        (rt)->bind_device();                   \
        __VA_ARGS__;                           \
        return RT_OK;                          \
    } catch (CudaFail & f) {                   \
        (rt)->last_error = f.what;             \
        return RT_ERR_CUDA;                    \
    } catch (std::exception & e) {             \
        (rt)->last_error = e.what();           \
        return RT_ERR_INVALID;                 \
    }

#define RT_GUARD_HOST(rt, ...)                 \
    if (!(rt)) return RT_ERR_INVALID;          \
    try {                                      \
        __VA_ARGS__;                           \
        return RT_OK;                          \
    } catch (CudaFail & f) {                   \
        (rt)->last_error = f.what;             \
        return RT_ERR_CUDA;                    \
    } catch (std::exception & e) {             \
        (rt)->last_error = e.what();           \
        return RT_ERR_INVALID;                 \
    }

extern "C" {

const char* rt_version(void) { return "rt_b200 0.2.0 sm_100a"; }
#ifndef RT_KERNELS_HASH
#define RT_KERNELS_HASH "unknown"
#endif
const char* rt_kernels_hash(void) { return RT_KERNELS_HASH; }

void rt_config_default(rt_config* cfg, uint32_t width, uint32_t height) {
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->width = width;
    cfg->height = height;
    cfg->triangles_per_leaf = RT_DEFAULT_TRIANGLES_PER_LEAF;
    cfg->rows_per_call = 50;
    cfg->recursions = 2;
    cfg->sub_spread = 1;
    cfg->jitter_mode = RT_JITTER_HASHED;
    cfg->seed = 0;
    cfg->accel = RT_ACCEL_BVH;
    cfg->device = -1;
    cfg->shard_index = 0;
    cfg->shard_count = 1;
    cfg->band_rows = 8;
}

int rt_scene_load_file(const char* collada_filename, rt_scene** out, char* err, size_t err_len) {
    if (!collada_filename || !out) return RT_ERR_INVALID;
    auto s = std::make_unique<rt_scene>();
    std::string e;
    if (!load_guarded([&] { return load_collada_file(collada_filename, &s->scene, &e); }, &e)) {
        copy_err(e, err, err_len);
        return RT_ERR_LOAD;
    }
    *out = s.release();
    return RT_OK;
}

int rt_scene_load_str(const char* collada_doc, const char* data_dir, rt_scene** out, char* err, size_t err_len) {
    if (!collada_doc || !out) return RT_ERR_INVALID;
    auto s = std::make_unique<rt_scene>();
    std::string e;
    if (!load_guarded([&] { return load_collada_str(collada_doc, data_dir, &s->scene, &e); }, &e)) {
        copy_err(e, err, err_len);
        return RT_ERR_LOAD;
    }
    *out = s.release();
    return RT_OK;
}

int rt_scene_get_desc(const rt_scene* scene, rt_scene_desc* desc) {
    if (!scene || !desc) return RT_ERR_INVALID;
    const_cast<rt_scene*>(scene)->scene.make_desc(desc);
    return RT_OK;
}

void rt_scene_free(rt_scene* scene) { delete scene; }

static int create_from_host_scene(const HostScene& sc, const rt_config& cfg, rt_raytracer** out, char* err, size_t err_len) {
    if (!sc.has_camera) {
        copy_err("scene has no camera (the reference indexes scene.cameras[0], lib.rs:39)", err, err_len);
        return RT_ERR_LOAD;
    }
    auto rt = std::make_unique<rt_raytracer>();
    try {
        rt->init(sc, cfg);
    } catch (CudaFail& f) {
        copy_err(f.what, err, err_len);
        return RT_ERR_CUDA;
    } catch (std::exception& e) {
        copy_err(e.what(), err, err_len);
        return RT_ERR_INVALID;
    }
    *out = rt.release();
    return RT_OK;
}

int rt_create(const rt_scene_desc* scene, const rt_config* cfg, rt_raytracer** out, char* err, size_t err_len) {
    if (!scene || !cfg || !out) return RT_ERR_INVALID;
    // a null array with a non-zero count would be read out of bounds below and in HostScene::from_desc
    if ((scene->num_triangles && (!scene->vertices || !scene->tri_geom)) || (scene->num_geometries && !scene->materials) ||
        (scene->num_lights && !scene->lights) || (scene->num_textures && !scene->textures)) {
        copy_err("scene description has a null array with a non-zero count", err, err_len);
        return RT_ERR_INVALID;
    }
    for (uint32_t t = 0; t < scene->num_triangles; ++t)
        if (scene->tri_geom[t] >= scene->num_geometries) {
            copy_err("tri_geom entry out of range", err, err_len);
            return RT_ERR_INVALID;
        }
    for (uint32_t k = 0; k < scene->num_textures; ++k)
        if (!scene->textures[k].rgb || scene->textures[k].width == 0 || scene->textures[k].height == 0) {
            copy_err("texture without texels", err, err_len);
            return RT_ERR_INVALID;
        }
    return create_from_host_scene(HostScene::from_desc(*scene), *cfg, out, err, err_len);
}

int rt_create_raytracer(const char* collada_doc, size_t triangles_per_leaf, size_t width, size_t height, rt_raytracer** out, char* err,
                        size_t err_len) {
    if (!collada_doc || !out) return RT_ERR_INVALID;
    HostScene sc;
    std::string e;
    if (!load_guarded([&] { return load_collada_str(collada_doc, nullptr, &sc, &e); }, &e)) {
        copy_err(e, err, err_len);
        return RT_ERR_LOAD;
    }
    rt_config cfg;
    rt_config_default(&cfg, (uint32_t)width, (uint32_t)height);
    cfg.triangles_per_leaf = (uint32_t)triangles_per_leaf;
    return create_from_host_scene(sc, cfg, out, err, err_len);
}

int rt_create_raytracer_from_file(const char* collada_filename, size_t triangles_per_leaf, size_t width, size_t height, rt_raytracer** out,
                                  char* err, size_t err_len) {
    if (!collada_filename || !out) return RT_ERR_INVALID;
    HostScene sc;
    std::string e;
    if (!load_guarded([&] { return load_collada_file(collada_filename, &sc, &e); }, &e)) {
        copy_err(e, err, err_len);
        return RT_ERR_LOAD;
    }
    rt_config cfg;
    rt_config_default(&cfg, (uint32_t)width, (uint32_t)height);
    cfg.triangles_per_leaf = (uint32_t)triangles_per_leaf;
    return create_from_host_scene(sc, cfg, out, err, err_len);
}

void rt_destroy(rt_raytracer* rt) {
    if (!rt) return;
    if (!rt->host_only) {
        cudaSetDevice(rt->device);
        cudaStreamSynchronize(rt->stream);
    }
    delete rt;
}

const char* rt_last_error(const rt_raytracer* rt) { return rt ? rt->last_error.c_str() : "null handle"; }

int rt_configure(rt_raytracer* rt, int32_t recursions, uint32_t sub_spread, int32_t jitter_mode, uint32_t seed, int32_t accel) {
    RT_GUARD_HOST(rt, {
        if (accel != RT_ACCEL_OCTREE && accel != RT_ACCEL_BVH && accel != RT_ACCEL_CWBVH && accel != RT_ACCEL_BVH4 && accel != RT_ACCEL_LBVH) throw std::invalid_argument("unknown accel");
        if (jitter_mode != RT_JITTER_FIXED_HALF && jitter_mode != RT_JITTER_HASHED) throw std::invalid_argument("unknown jitter mode");
        if (recursions < 0) throw std::invalid_argument("negative recursions");
        rt->drop_lap();
        rt->cfg.recursions = recursions;
        rt->cfg.sub_spread = sub_spread;
        rt->cfg.jitter_mode = jitter_mode;
        rt->cfg.seed = seed;
        if (rt->cfg.accel != accel) rt->reset_schedules();  // tile costs of one structure say little about another
        rt->cfg.accel = accel;
        if (!rt->host_only) {
            rt->bind_device();
            rt->ensure_accel(accel);
        }
    });
}

int rt_set_rows_per_call(rt_raytracer* rt, uint32_t rows) {
    RT_GUARD_HOST(rt, {
        if (rows == 0) throw std::invalid_argument("rows_per_call must be positive");
        rt->cfg.rows_per_call = rows;
    });
}

int rt_trace_frame_additive(rt_raytracer* rt, uint32_t* num_primary_rays) {
    RT_GUARD(rt, {
        const uint32_t rows = rt->cfg.rows_per_call;
        if (rt->lookahead_applies(rows)) rt->trace_band_lookahead(rt->current_row, rows);
        else rt->trace_rows(rt->current_row, rows, 1);
        rt->current_row = (rt->current_row + rows) % rt->cfg.height;
        // mod.rs:113-116 returns rows * width; a sharded handle traces (and reports) only the rows it owns
        if (num_primary_rays) *num_primary_rays = (uint32_t)rt->last.n_primary;
    });
}

int rt_trace_rows(rt_raytracer* rt, uint32_t first_row, uint32_t n_rows, uint32_t spp, uint64_t* n_primary, uint64_t* n_shadow) {
    RT_GUARD(rt, {
        rt->trace_rows(first_row, n_rows, spp);
        if (n_primary) *n_primary = rt->last.n_primary;
        if (n_shadow) {
            rt->finish_stats();
            *n_shadow = rt->last.n_shadow;
        }
    });
}

int rt_get_tonemapped_pixels(rt_raytracer* rt, uint32_t* out) {
    RT_GUARD(rt, {
        if (!out) throw std::invalid_argument("null output");
        if (out == rt->host_frame && !rt->host_frame_stale) {
            RT_CUDA(cudaStreamSynchronize(rt->stream));  // the kernel already stored the pixels into this buffer
        } else {
            RT_CUDA(cudaMemcpyAsync(out, rt->d_ldr.p, (size_t)rt->npix() * 4, cudaMemcpyDeviceToHost, rt->stream));
            RT_CUDA(cudaStreamSynchronize(rt->stream));
            if (out == rt->host_frame) rt->host_frame_stale = false;
        }
    });
}

int rt_get_tonemapped_pixels_delta(rt_raytracer* rt, uint32_t* out) {
    RT_GUARD(rt, {
        if (!out) throw std::invalid_argument("null output");
        const uint32_t W = rt->cfg.width, H = rt->cfg.height;
        if (out != rt->delta_host) rt->mark_frame_dirty_all();  // another buffer: nothing is known about its contents
        // one copy per run of consecutive changed rows (rows are contiguous in the frame)
        uint32_t r = 0;
        while (r < H) {
            if (!rt->dirty_rows[r]) {
                ++r;
                continue;
            }
            uint32_t e = r;
            while (e < H && rt->dirty_rows[e]) ++e;
            RT_CUDA(cudaMemcpyAsync(out + (size_t)r * W, rt->d_ldr.p + (size_t)r * W, (size_t)(e - r) * W * 4, cudaMemcpyDeviceToHost, rt->stream));
            r = e;
        }
        RT_CUDA(cudaStreamSynchronize(rt->stream));
        std::fill(rt->dirty_rows.begin(), rt->dirty_rows.end(), (uint8_t)0);
        rt->delta_host = out;
    });
}

int rt_get_tonemapped_pixels_async(rt_raytracer* rt, uint32_t* pinned_out) {
    RT_GUARD(rt, {
        if (!pinned_out) throw std::invalid_argument("null output");
        if (!rt->copy_stream) {
            RT_CUDA(cudaStreamCreateWithFlags(&rt->copy_stream, cudaStreamNonBlocking));
            RT_CUDA(cudaEventCreateWithFlags(&rt->ev_snapshot, cudaEventDisableTiming));
            for (int k = 0; k < 2; ++k) {
                RT_CUDA(cudaEventCreateWithFlags(&rt->ev_copied[k], cudaEventDisableTiming));
                rt->d_ldr_snapshot[k].alloc(rt->npix());
            }
        }
        const size_t bytes = (size_t)rt->npix() * 4;
        const int slot = (int)(rt->copies_issued & 1u);
        // the snapshot buffer is free again once the copy that used it two calls ago has left it
        if (rt->copies_issued >= 2) RT_CUDA(cudaStreamWaitEvent(rt->stream, rt->ev_copied[slot], 0));
        RT_CUDA(cudaMemcpyAsync(rt->d_ldr_snapshot[slot].p, rt->d_ldr.p, bytes, cudaMemcpyDeviceToDevice, rt->stream));
        RT_CUDA(cudaEventRecord(rt->ev_snapshot, rt->stream));
        RT_CUDA(cudaStreamWaitEvent(rt->copy_stream, rt->ev_snapshot, 0));
        RT_CUDA(cudaMemcpyAsync(pinned_out, rt->d_ldr_snapshot[slot].p, bytes, cudaMemcpyDeviceToHost, rt->copy_stream));
        RT_CUDA(cudaEventRecord(rt->ev_copied[slot], rt->copy_stream));
        // an event is about to be reused: the host must not still be counting on its previous recording
        if (rt->copies_issued >= 2 && rt->copies_waited + 2 <= rt->copies_issued) rt->copies_waited = rt->copies_issued - 1;
        ++rt->copies_issued;
    });
}
int rt_wait_pixels_keep(rt_raytracer* rt, uint32_t keep) {
    RT_GUARD(rt, {
        // copies complete in order on the copy stream: waiting for copy c implies every earlier one
        while (rt->copies_issued - rt->copies_waited > (uint64_t)keep) {
            RT_CUDA(cudaEventSynchronize(rt->ev_copied[rt->copies_waited & 1u]));
            ++rt->copies_waited;
        }
    });
}
int rt_wait_pixels(rt_raytracer* rt) { return rt_wait_pixels_keep(rt, 0); }
int rt_set_host_frame(rt_raytracer* rt, uint32_t* pinned_host_frame) {
    RT_GUARD(rt, {
        if (rt->ldr_remote && rt->ldr_remote != rt->host_frame_dev) throw std::invalid_argument("an LDR target is already set (rt_set_ldr_target)");
        if (!pinned_host_frame) {
            rt->host_frame = nullptr;
            rt->host_frame_dev = nullptr;
            rt->ldr_remote = nullptr;
            return RT_OK;
        }
        void* dev = nullptr;
        RT_CUDA(cudaHostGetDevicePointer(&dev, pinned_host_frame, 0));  // fails unless the buffer is page-locked and mapped
        rt->host_frame = pinned_host_frame;
        rt->host_frame_dev = (uint32_t*)dev;
        rt->ldr_remote = (uint32_t*)dev;
        rt->host_frame_stale = true;  // the first readback copies the whole frame, later ones only synchronise
    });
}

int rt_film_clear(rt_raytracer* rt) { RT_GUARD(rt, { rt->film_clear(); }); }

int rt_get_film(rt_raytracer* rt, float* out) {
    RT_GUARD(rt, {
        if (!out) throw std::invalid_argument("null output");
        const size_t n = rt->npix();
        std::vector<float4> sum(n), sq(n);
        RT_CUDA(cudaMemcpyAsync(sum.data(), rt->d_film_sum.p, n * sizeof(float4), cudaMemcpyDeviceToHost, rt->stream));
        RT_CUDA(cudaMemcpyAsync(sq.data(), rt->d_film_sq.p, n * sizeof(float4), cudaMemcpyDeviceToHost, rt->stream));
        RT_CUDA(cudaStreamSynchronize(rt->stream));
        for (size_t i = 0; i < n; ++i) {
            uint32_t cnt;
            std::memcpy(&cnt, &sum[i].w, 4);
            float* o = out + 7 * i;
            o[0] = sum[i].x;
            o[1] = sum[i].y;
            o[2] = sum[i].z;
            o[3] = sq[i].x;
            o[4] = sq[i].y;
            o[5] = sq[i].z;
            o[6] = (float)cnt;
        }
    });
}

int rt_set_film(rt_raytracer* rt, const float* in) {
    RT_GUARD(rt, {
        if (!in) throw std::invalid_argument("null input");
        const size_t n = rt->npix();
        std::vector<float4> sum(n), sq(n);
        for (size_t i = 0; i < n; ++i) {
            const float* o = in + 7 * i;
            if (!(o[6] >= 0.0f && o[6] < 4294967296.0f)) throw std::invalid_argument("num_samples out of range");
            const uint32_t cnt = (uint32_t)o[6];
            float w;
            std::memcpy(&w, &cnt, 4);
            sum[i] = make_float4(o[0], o[1], o[2], w);
            sq[i] = make_float4(o[3], o[4], o[5], 0.f);
        }
        rt->drop_lap();
        RT_CUDA(cudaMemcpyAsync(rt->d_film_sum.p, sum.data(), n * sizeof(float4), cudaMemcpyHostToDevice, rt->stream));
        RT_CUDA(cudaMemcpyAsync(rt->d_film_sq.p, sq.data(), n * sizeof(float4), cudaMemcpyHostToDevice, rt->stream));
        // the packed frame follows the film (get_tonemapped_pixels is a function of the film alone, mod.rs:120-128)
        RT_CUDA(launch_tonemap(rt->d_film_sum.p, rt->d_ldr.p, (uint32_t)n, rt->stream));
        ++rt->total_kernels;
        rt->mark_frame_dirty_all();
        RT_CUDA(cudaStreamSynchronize(rt->stream));  // `sum` / `sq` go out of scope
    });
}

int rt_get_estimated_variances(rt_raytracer* rt, float* out) {
    RT_GUARD(rt, {
        if (!out) throw std::invalid_argument("null output");
        const size_t n = rt->npix();
        if (rt->d_variance.n < 3 * n) rt->d_variance.alloc(3 * n);
        RT_CUDA(launch_film_variance(rt->d_film_sum.p, rt->d_film_sq.p, rt->d_variance.p, (uint32_t)n, rt->stream));
        ++rt->total_kernels;
        RT_CUDA(cudaMemcpyAsync(out, rt->d_variance.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, rt->stream));
        RT_CUDA(cudaStreamSynchronize(rt->stream));
    });
}

int rt_get_primary_ids(rt_raytracer* rt, uint32_t* out) {
    RT_GUARD(rt, {
        if (!out) throw std::invalid_argument("null output");
        RT_CUDA(cudaMemcpyAsync(out, rt->d_ids.p, (size_t)rt->npix() * 4, cudaMemcpyDeviceToHost, rt->stream));
        RT_CUDA(cudaStreamSynchronize(rt->stream));
    });
}

int rt_camera_move_rel(rt_raytracer* rt, float x, float y, float z) {
    if (!rt) return RT_ERR_INVALID;
    rt->camera.move_rel(x, y, z);
    rt->invalidate_schedule();
    return RT_OK;
}
int rt_camera_add_x_angle(rt_raytracer* rt, float radians) {
    if (!rt) return RT_ERR_INVALID;
    rt->camera.add_x_angle(radians);
    rt->invalidate_schedule();
    return RT_OK;
}
int rt_camera_add_y_angle(rt_raytracer* rt, float radians) {
    if (!rt) return RT_ERR_INVALID;
    rt->camera.add_y_angle(radians);
    rt->invalidate_schedule();
    return RT_OK;
}
int rt_camera_get(const rt_raytracer* rt, float* out34) {
    if (!rt || !out34) return RT_ERR_INVALID;
    std::memcpy(out34, rt->camera.rotation.data(), 64);
    std::memcpy(out34 + 16, rt->camera.orientation.data(), 64);
    out34[32] = rt->camera.max_x;
    out34[33] = rt->camera.max_y;
    return RT_OK;
}
int rt_camera_set_state(rt_raytracer* rt, float x_angle, float y_angle, const float pos[3]) {
    if (!rt || !pos) return RT_ERR_INVALID;
    const bool changed = rt->camera.x_angle != x_angle || rt->camera.y_angle != y_angle || rt->camera.pos.x != pos[0] ||
                         rt->camera.pos.y != pos[1] || rt->camera.pos.z != pos[2];
    rt->camera.x_angle = x_angle;
    rt->camera.y_angle = y_angle;
    rt->camera.pos = f3{pos[0], pos[1], pos[2]};
    rt->camera.update_matrices();
    if (changed) rt->invalidate_schedule();
    return RT_OK;
}

int rt_set_stream(rt_raytracer* rt, void* cuda_stream) {
    RT_GUARD(rt, {
        RT_CUDA(cudaStreamSynchronize(rt->stream));
        rt->stream = (cudaStream_t)cuda_stream;
    });
}
int rt_get_ldr_device_ptr(rt_raytracer* rt, void** dev_ptr) {
    if (!rt || !dev_ptr) return RT_ERR_INVALID;
    *dev_ptr = rt->d_ldr.p;
    return RT_OK;
}
int rt_set_ldr_target(rt_raytracer* rt, void* dev_ptr) {
    if (!rt) return RT_ERR_INVALID;
    if (rt->host_only) {
        rt->last_error = "this handle was created with RT_DEVICE_NONE";
        return RT_ERR_CUDA;
    }
    if (rt->host_frame) {  // the second store target is taken by the registered host frame (rt_set_host_frame): refuse rather than
                           // silently stop updating a frame whose readback only synchronises
        rt->last_error = "a host frame is registered (rt_set_host_frame); unregister it before setting an LDR target";
        return RT_ERR_INVALID;
    }
    rt->ldr_remote = (uint32_t*)dev_ptr;
    return RT_OK;
}
int rt_set_done_signal(rt_raytracer* rt, void* dev_flag, uint32_t value) {
    if (!rt) return RT_ERR_INVALID;
    if (rt->host_only) {
        rt->last_error = "this handle was created with RT_DEVICE_NONE";
        return RT_ERR_CUDA;
    }
    rt->done_flag = (uint32_t*)dev_flag;
    rt->done_value = value;
    return RT_OK;
}
int rt_get_owned_ldr_rows_device(rt_raytracer* rt, void* dev_out, uint32_t* n_rows) {
    RT_GUARD(rt, {
        if (n_rows) *n_rows = (uint32_t)rt->owned_rows.size();
        if (dev_out) {
            RT_CUDA(launch_gather_rows(rt->d_ldr.p, rt->d_owned_rows.p, (uint32_t)rt->owned_rows.size(), rt->cfg.width, (uint32_t*)dev_out,
                                       rt->stream));
            ++rt->total_kernels;
        }
    });
}
int rt_device_alloc(rt_raytracer* rt, size_t bytes, void** dev_ptr) {
    RT_GUARD(rt, {
        if (!dev_ptr || bytes == 0) throw std::invalid_argument("bad allocation request");
        RT_CUDA(cudaMalloc(dev_ptr, bytes));
        RT_CUDA(cudaMemsetAsync(*dev_ptr, 0xFF, bytes, rt->stream));
    });
}
int rt_device_free(rt_raytracer* rt, void* dev_ptr) {
    RT_GUARD(rt, {
        RT_CUDA(cudaStreamSynchronize(rt->stream));
        RT_CUDA(cudaFree(dev_ptr));
    });
}
int rt_ipc_export(rt_raytracer* rt, void* dev_ptr, uint8_t* handle64) {
    RT_GUARD(rt, {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        if (!dev_ptr || !handle64) throw std::invalid_argument("null argument");
        cudaIpcMemHandle_t h;
        RT_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
        std::memcpy(handle64, &h, 64);
    });
}
int rt_ipc_open(rt_raytracer* rt, const uint8_t* handle64, void** dev_ptr) {
    RT_GUARD(rt, {
        if (!dev_ptr || !handle64) throw std::invalid_argument("null argument");
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle64, 64);
        RT_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    });
}
int rt_ipc_close(rt_raytracer* rt, void* dev_ptr) {
    RT_GUARD(rt, {
        RT_CUDA(cudaStreamSynchronize(rt->stream));
        if (rt->ldr_remote == dev_ptr) rt->ldr_remote = nullptr;
        RT_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    });
}
int rt_stream_signal_flag(rt_raytracer* rt, void* dev_flag, uint32_t value) {
    RT_GUARD(rt, {
        if (!dev_flag) throw std::invalid_argument("null flag");
        RT_CUDA(launch_flag_signal((uint32_t*)dev_flag, value, rt->stream));
        ++rt->total_kernels;
    });
}
int rt_stream_wait_flags(rt_raytracer* rt, void* dev_flags, uint32_t n_flags, uint32_t target, int32_t signal_slot, int32_t release_slot) {
    RT_GUARD(rt, {
        if (!dev_flags) throw std::invalid_argument("null flags");
        if (!rt->d_sync_timeouts.p) {
            rt->d_sync_timeouts.alloc(1);
            RT_CUDA(cudaMemsetAsync(rt->d_sync_timeouts.p, 0, 4, rt->stream));
        }
        RT_CUDA(launch_flag_wait((uint32_t*)dev_flags, n_flags, target, signal_slot, release_slot, rt->d_sync_timeouts.p, rt->stream));
        ++rt->total_kernels;
    });
}
int rt_stream_signal_then_wait(rt_raytracer* rt, void* dev_signal_flag, uint32_t value, void* dev_wait_flag, uint32_t target) {
    RT_GUARD(rt, {
        if (!dev_signal_flag || !dev_wait_flag) throw std::invalid_argument("null flag");
        if (!rt->d_sync_timeouts.p) {
            rt->d_sync_timeouts.alloc(1);
            RT_CUDA(cudaMemsetAsync(rt->d_sync_timeouts.p, 0, 4, rt->stream));
        }
        RT_CUDA(launch_flag_signal_wait((uint32_t*)dev_signal_flag, value, (uint32_t*)dev_wait_flag, target, rt->d_sync_timeouts.p, rt->stream));
        ++rt->total_kernels;
    });
}
// Stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32), fetched from the driver at run time: a frame fence without
// a kernel launch — the stream's front end stores / polls the flag itself.
namespace {
typedef CUresult (*StreamValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamValueFn driver_fn(const char* name) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return (StreamValueFn)fn;
}
}  // namespace
int rt_stream_write_value(rt_raytracer* rt, void* dev_flag, uint32_t value) {
    RT_GUARD(rt, {
        static StreamValueFn fn = driver_fn("cuStreamWriteValue32");
        if (!fn || !dev_flag) {
            rt->last_error = "cuStreamWriteValue32 is not available";
            return RT_ERR_UNSUPPORTED;
        }
        // default flags: a system-wide memory barrier precedes the write, so everything the stream did before is visible first
        const CUresult r = fn((CUstream)rt->stream, (CUdeviceptr)(uintptr_t)dev_flag, value, CU_STREAM_WRITE_VALUE_DEFAULT);
        if (r != CUDA_SUCCESS) {
            rt->last_error = "cuStreamWriteValue32 failed (" + std::to_string((int)r) + ")";
            return RT_ERR_UNSUPPORTED;
        }
    });
}
int rt_stream_wait_value(rt_raytracer* rt, void* dev_flag, uint32_t value) {
    RT_GUARD(rt, {
        static StreamValueFn fn = driver_fn("cuStreamWaitValue32");
        if (!fn || !dev_flag) {
            rt->last_error = "cuStreamWaitValue32 is not available";
            return RT_ERR_UNSUPPORTED;
        }
        // GEQ is cyclic: (int32)(*flag - value) >= 0, the comparison the flag kernels use
        const CUresult r = fn((CUstream)rt->stream, (CUdeviceptr)(uintptr_t)dev_flag, value, CU_STREAM_WAIT_VALUE_GEQ);
        if (r != CUDA_SUCCESS) {
            rt->last_error = "cuStreamWaitValue32 failed (" + std::to_string((int)r) + ")";
            return RT_ERR_UNSUPPORTED;
        }
    });
}
int rt_sync_timeouts(rt_raytracer* rt, uint32_t* count) {
    RT_GUARD(rt, {
        if (!count) throw std::invalid_argument("null output");
        *count = 0;
        if (rt->d_sync_timeouts.p) {
            RT_CUDA(cudaMemcpyAsync(count, rt->d_sync_timeouts.p, 4, cudaMemcpyDeviceToHost, rt->stream));
            RT_CUDA(cudaStreamSynchronize(rt->stream));
        }
    });
}
int rt_host_register(rt_raytracer* rt, void* host_ptr, size_t bytes, void** dev_ptr) {
    RT_GUARD(rt, {
        if (!host_ptr || bytes == 0) throw std::invalid_argument("bad registration request");
        RT_CUDA(cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
        if (dev_ptr) {
            void* d = nullptr;
            RT_CUDA(cudaHostGetDevicePointer(&d, host_ptr, 0));
            *dev_ptr = d;
        }
    });
}
int rt_host_unregister(rt_raytracer* rt, void* host_ptr) {
    RT_GUARD(rt, {
        RT_CUDA(cudaStreamSynchronize(rt->stream));
        RT_CUDA(cudaHostUnregister(host_ptr));
    });
}
int rt_copy_owned_rows(rt_raytracer* rt, const void* src_frame, void* dst_frame, void* cuda_stream) {
    RT_GUARD(rt, {
        if (!src_frame || !dst_frame) throw std::invalid_argument("null frame");
        const size_t row_bytes = (size_t)rt->cfg.width * 4;
        const uint32_t H = rt->cfg.height, B = rt->cfg.band_rows, N = rt->cfg.shard_count, r = rt->cfg.shard_index;
        cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : rt->stream;
        // owned bands r, r + N, r + 2N, ...: one strided 2-D copy for the full bands (a band is B consecutive rows = one contiguous
        // run of the frame), one more for a partial band at the bottom of the image
        const uint32_t full_bands = H / B;  // bands 0 .. full_bands - 1 have B rows
        const uint32_t mine_full = full_bands > r ? (full_bands - r + N - 1) / N : 0;
        const size_t band_bytes = (size_t)B * row_bytes, pitch = (size_t)N * band_bytes, off = (size_t)r * band_bytes;
        if (mine_full)
            RT_CUDA(cudaMemcpy2DAsync((char*)dst_frame + off, pitch, (const char*)src_frame + off, pitch, band_bytes, mine_full, cudaMemcpyDefault, st));
        if (H % B != 0 && full_bands % N == r) {
            const size_t tail = (size_t)full_bands * band_bytes;
            RT_CUDA(cudaMemcpyAsync((char*)dst_frame + tail, (const char*)src_frame + tail, (size_t)(H % B) * row_bytes, cudaMemcpyDefault, st));
        }
    });
}
int rt_signal_flag_on_stream(rt_raytracer* rt, void* dev_flag, uint32_t value, void* cuda_stream) {
    RT_GUARD(rt, {
        if (!dev_flag) throw std::invalid_argument("null flag");
        RT_CUDA(launch_flag_signal((uint32_t*)dev_flag, value, cuda_stream ? (cudaStream_t)cuda_stream : rt->stream));
        ++rt->total_kernels;
    });
}
int rt_get_counters_device_ptr(rt_raytracer* rt, void** dev_ptr) {
    if (!rt || !dev_ptr || rt->host_only) return RT_ERR_INVALID;
    *dev_ptr = rt->d_counters.p + (rt->call_parity ? CNT_SET_B : CNT_SET_A);  // the set the last trace call counted into
    return RT_OK;
}
uint32_t rt_launch_param_bytes(void) { return (uint32_t)sizeof(TraceParams); }

int rt_get_camera_plane_matrix(rt_raytracer* rt, double* out12) {
    if (!out12) return RT_ERR_INVALID;
    RT_GUARD_HOST(rt, {
        if (!rt->camera_plane_matrix(out12)) throw std::invalid_argument("the camera matrix cannot be inverted");
        const f3 o = rt->camera.ray_origin();
        out12[9] = o.x, out12[10] = o.y, out12[11] = o.z;
    });
}
int rt_set_tuning(rt_raytracer* rt, int32_t key, int32_t value) {
    if (!rt) return RT_ERR_INVALID;
    if (key == RT_TUNE_KERNEL_VARIANT && value >= 0 && value <= 2) {
        if (rt->variant != value) rt->reset_schedules();  // an order holds items in the units and granularity of one kernel
        rt->variant = value;
        rt->drop_lap();
        return RT_OK;
    }
    if (key == RT_TUNE_MIN_SCHEDULE_TILES && value >= 1) {
        rt->min_schedule_tiles = value;
        rt->reset_schedules();
        return RT_OK;
    }
    if (key == RT_TUNE_MAX_SPLIT_LEVEL && value >= 0 && value <= 3) {
        rt->max_split_level = value;
        rt->reset_schedules();
        return RT_OK;
    }
    if (key == RT_TUNE_POOL_REFILL && value >= 1 && value <= 32) {
        rt->pool_refill = value;
        return RT_OK;
    }
    if (key == RT_TUNE_QUEUE_BATCH && value >= 1 && value <= 64) {
        rt->queue_batch = value;
        return RT_OK;
    }
    if (key == RT_TUNE_QUEUE_BATCH_FROM && value >= 0 && value <= 100) {
        rt->queue_batch_from_pct = value;
        return RT_OK;
    }
    if (key == RT_TUNE_SPLIT_QUARTERS && value >= 0 && value <= 64) {
        rt->split_quarters = value;
        rt->reset_schedules();
        return RT_OK;
    }
    if (key == RT_TUNE_BOUNCE_WAVEFRONT && (value == 0 || value == 1)) {
        rt->bounce_wavefront = value != 0;
        return RT_OK;
    }
    if (key == RT_TUNE_BOUNCE_STREAM && (value == 0 || value == 1)) {
        rt->bounce_stream = value != 0;
        return RT_OK;
    }
    if (key == RT_TUNE_STREAM_REFILL && value >= 1 && value <= 32) {
        rt->stream_refill = value;
        return RT_OK;
    }
    if (key == RT_TUNE_LIGHT_GRID && (value == 0 || (value >= 6 && value <= 9))) {
        rt->light_grid_log2 = value;
        return RT_OK;
    }
    if (key == RT_TUNE_CAMERA_GRID_AFTER && value >= 0 && value <= 1000) {
        rt->camera_grid_after = value;
        return RT_OK;
    }
    if (key == RT_TUNE_CAMERA_GRID && (value == 0 || (value >= 2 && value <= 5))) {
        rt->camera_grid_log2 = value;
        rt->pg_valid = false;
        return RT_OK;
    }
    if (key == RT_TUNE_FILM_PREFETCH_ROWS_MB && value >= 0 && value <= 1024) {
        rt->film_prefetch_rows_mb = value;
        return RT_OK;
    }
    if (key == RT_TUNE_FILM_PREFETCH && value >= 0 && value <= 3) {
        rt->film_prefetch = value;
        return RT_OK;
    }
    if (key == RT_TUNE_BAND_LOOKAHEAD && (value == 0 || value == 1)) {
        rt->band_lookahead = value;
        rt->drop_lap();
        return RT_OK;
    }
    if (key == RT_TUNE_STREAM_CHAIN && (value == 0 || value == 1)) {
        rt->stream_chain = value != 0;
        return RT_OK;
    }
    if (key == RT_TUNE_STREAM_BLOCKS && value >= 3 && value <= 5) {
        rt->stream_blocks = value;
        return RT_OK;
    }
    if (key == RT_TUNE_WF_BLOCKS && value >= 0 && value <= 8) {
        rt->wf_blocks_cap = value;
        return RT_OK;
    }
    if (key == RT_TUNE_STREAM_MIN_INNER && value >= 0 && value <= 32) {
        rt->stream_min_inner = value;
        return RT_OK;
    }
    if (key == RT_TUNE_MULTI_SAMPLE_LAUNCH && value >= 0 && value <= 2) {
        rt->multi_sample_launch = value;
        return RT_OK;
    }
    if (key == RT_TUNE_POOL_MIN_INNER && value >= 0 && value <= 32) {
        rt->pool_min_inner = value;
        return RT_OK;
    }
    if (key == RT_TUNE_POOL_BLOCKS && value >= 0 && value <= 8) {
        rt->pool_blocks = value;  // 0 = as many as fit
        return RT_OK;
    }
    if (key == RT_TUNE_TIME_LAUNCHES && (value == 0 || value == 1)) {
        rt->time_launches = value;
        return RT_OK;
    }
    if (key == RT_TUNE_TILE_SCHEDULE && (value == 0 || value == 1)) {
        rt->lpt_schedule = value;
        rt->reset_schedules();
        return RT_OK;
    }
    rt->last_error = "unknown tuning key or value";
    return RT_ERR_INVALID;
}

int rt_get_tile_costs(rt_raytracer* rt, uint32_t* out, uint32_t capacity, uint32_t* n_tiles, uint32_t* n_items) {
    RT_GUARD(rt, {
        const rt_raytracer::TileSchedule* sc = nullptr;
        for (auto& s : rt->schedules)
            if (!sc || s->last_use > sc->last_use) sc = s.get();
        if (n_tiles) *n_tiles = sc ? sc->tiles : 0;
        if (n_items) *n_items = 0;
        if (!sc) return RT_OK;
        if (out && capacity) RT_CUDA(cudaMemcpyAsync(out, sc->cost.p, sizeof(uint32_t) * std::min(capacity, sc->tiles), cudaMemcpyDeviceToHost, rt->stream));
        if (n_items && sc->have_order) RT_CUDA(cudaMemcpyAsync(n_items, sc->order.p + (sc->order.n - 1), 4, cudaMemcpyDeviceToHost, rt->stream));
        RT_CUDA(cudaStreamSynchronize(rt->stream));
    });
}

int rt_get_launch_stats(const rt_raytracer* rt_c, rt_launch_stats* out) {
    rt_raytracer* rt = const_cast<rt_raytracer*>(rt_c);
    RT_GUARD(rt, {
        if (!out) throw std::invalid_argument("null output");
        rt->finish_stats();
        *out = rt->last;
    });
}
uint64_t rt_kernels_launched(const rt_raytracer* rt) { return rt ? rt->total_kernels : 0; }

int rt_get_ray_totals(rt_raytracer* rt, uint64_t* out3) {
    RT_GUARD(rt, {
        if (!out3) throw std::invalid_argument("null output");
        unsigned long long tot[2];
        RT_CUDA(cudaMemcpyAsync(tot, rt->d_counters.p + CNT_SHADOW_TOTAL, sizeof(tot), cudaMemcpyDeviceToHost, rt->stream));
        RT_CUDA(cudaStreamSynchronize(rt->stream));
        out3[0] = rt->total_primary;
        out3[1] = tot[0];
        out3[2] = tot[1];
    });
}

int rt_octree_stats(const rt_raytracer* rt_c, uint64_t* out) {
    rt_raytracer* rt = const_cast<rt_raytracer*>(rt_c);
    if (!rt || !out) return RT_ERR_INVALID;
    rt->ensure_octree_host();
    const FlatOctree& o = rt->octree;
    uint64_t inner = 0, leaves = 0, empty = 0;
    for (size_t i = 0; i < o.num_nodes(); ++i) {
        if (o.first_child[i] < 0) {
            ++leaves;
            if (o.leaf_offset[i + 1] == o.leaf_offset[i]) ++empty;
        } else
            ++inner;
    }
    out[0] = o.num_nodes();
    out[1] = inner;
    out[2] = leaves;
    out[3] = empty;
    out[4] = o.leaf_tris.size();
    out[5] = o.depth;
    return RT_OK;
}
int rt_octree_export(const rt_raytracer* rt_c, float* cubes, int32_t* first_child, uint32_t* leaf_offset, uint32_t* leaf_tris,
                     uint64_t* n_refs) {
    rt_raytracer* rt = const_cast<rt_raytracer*>(rt_c);
    if (!rt) return RT_ERR_INVALID;
    rt->ensure_octree_host();
    const FlatOctree& o = rt->octree;
    if (cubes) std::memcpy(cubes, o.cubes.data(), o.cubes.size() * 4);
    if (first_child) std::memcpy(first_child, o.first_child.data(), o.first_child.size() * 4);
    if (leaf_offset) std::memcpy(leaf_offset, o.leaf_offset.data(), o.leaf_offset.size() * 4);
    if (leaf_tris) std::memcpy(leaf_tris, o.leaf_tris.data(), o.leaf_tris.size() * 4);
    if (n_refs) *n_refs = o.leaf_tris.size();
    return RT_OK;
}
int rt_bvh_stats(const rt_raytracer* rt_c, uint64_t* out) {
    rt_raytracer* rt = const_cast<rt_raytracer*>(rt_c);
    if (!rt || !out) return RT_ERR_INVALID;
    rt->ensure_bvh_host();
    out[0] = rt->bvh.nodes.size();
    out[1] = rt->bvh.num_leaves;
    out[2] = rt->bvh.max_leaf;
    out[3] = rt->bvh.depth;
    return RT_OK;
}
int rt_bvh_export(const rt_raytracer* rt_c, float* boxes, int32_t* children, int32_t* counts, uint32_t* tri_order) {
    rt_raytracer* rt = const_cast<rt_raytracer*>(rt_c);
    if (!rt) return RT_ERR_INVALID;
    rt->ensure_bvh_host();
    for (size_t i = 0; i < rt->bvh.nodes.size(); ++i) {
        const FlatBvh::Node& n = rt->bvh.nodes[i];
        for (int k = 0; k < 2; ++k) {
            if (boxes) {
                std::memcpy(boxes + 12 * i + 6 * k, n.lo[k], 12);
                std::memcpy(boxes + 12 * i + 6 * k + 3, n.hi[k], 12);
            }
            if (children) children[2 * i + k] = n.child[k];
            if (counts) counts[2 * i + k] = n.count[k];
        }
    }
    if (tri_order && !rt->bvh.tri_order.empty()) std::memcpy(tri_order, rt->bvh.tri_order.data(), rt->bvh.tri_order.size() * 4);
    return RT_OK;
}
int rt_lbvh_build(rt_raytracer* rt, uint64_t* out3, float* build_ms) {
    RT_GUARD(rt, {
        rt->d_lbvh_nodes.release();  // rebuild on every call: this entry point is also the build benchmark
        rt->build_lbvh();
        if (out3) {
            out3[0] = rt->lbvh_nodes;
            out3[1] = rt->lbvh_depth;
            out3[2] = rt->scene.num_triangles();
        }
        if (build_ms) *build_ms = rt->lbvh_build_ms;
    });
}
int rt_lbvh_export(rt_raytracer* rt, float* boxes, int32_t* children, int32_t* counts, uint32_t* tri_order) {
    RT_GUARD(rt, {
        if (!rt->d_lbvh_nodes.p) rt->build_lbvh();
        const size_t nn = rt->lbvh_nodes;
        std::vector<float4> nodes(4 * nn);
        RT_CUDA(cudaMemcpyAsync(nodes.data(), rt->d_lbvh_nodes.p, nodes.size() * sizeof(float4), cudaMemcpyDeviceToHost, rt->stream));
        if (tri_order && rt->scene.num_triangles())
            RT_CUDA(cudaMemcpyAsync(tri_order, rt->d_lbvh_order.p, (size_t)rt->scene.num_triangles() * 4, cudaMemcpyDeviceToHost, rt->stream));
        RT_CUDA(cudaStreamSynchronize(rt->stream));
        for (size_t i = 0; i < nn; ++i) {
            const float* f = reinterpret_cast<const float*>(&nodes[4 * i]);
            if (boxes) std::memcpy(boxes + 12 * i, f, 48);
            for (int k = 0; k < 2; ++k) {
                int32_t ref;
                std::memcpy(&ref, f + 12 + k, 4);
                int32_t child = ref, count = 0;
                if (ref < 0) {
                    const uint32_t r = (uint32_t)~ref;
                    child = ~(int32_t)(r >> 4);
                    count = (int32_t)(r & 15u);
                }
                if (children) children[2 * i + k] = child;
                if (counts) counts[2 * i + k] = count;
            }
        }
    });
}
int rt_bvh4_stats(const rt_raytracer* rt_c, uint64_t* out) {
    rt_raytracer* rt = const_cast<rt_raytracer*>(rt_c);
    if (!rt || !out) return RT_ERR_INVALID;
    rt->ensure_bvh4_host();
    out[0] = rt->bvh4.nodes.size();
    out[1] = rt->bvh4.num_leaves;
    out[2] = rt->bvh4.max_leaf;
    out[3] = rt->bvh4.depth;
    return RT_OK;
}
int rt_bvh4_export(const rt_raytracer* rt_c, float* boxes, int32_t* children, int32_t* counts, uint32_t* tri_order) {
    rt_raytracer* rt = const_cast<rt_raytracer*>(rt_c);
    if (!rt) return RT_ERR_INVALID;
    rt->ensure_bvh4_host();
    for (size_t i = 0; i < rt->bvh4.nodes.size(); ++i) {
        const FlatBvh4::Node& n = rt->bvh4.nodes[i];
        for (int k = 0; k < 4; ++k) {
            if (boxes) {
                std::memcpy(boxes + 24 * i + 6 * k, n.lo[k], 12);
                std::memcpy(boxes + 24 * i + 6 * k + 3, n.hi[k], 12);
            }
            if (children) children[4 * i + k] = n.child[k];
            if (counts) counts[4 * i + k] = n.count[k];
        }
    }
    if (tri_order && !rt->bvh4.tri_order.empty()) std::memcpy(tri_order, rt->bvh4.tri_order.data(), rt->bvh4.tri_order.size() * 4);
    return RT_OK;
}
int rt_cwbvh_stats(const rt_raytracer* rt_c, uint64_t* out) {
    rt_raytracer* rt = const_cast<rt_raytracer*>(rt_c);
    if (!rt || !out) return RT_ERR_INVALID;
    rt->ensure_cwbvh_host();
    out[0] = rt->cwbvh.num_nodes();
    out[1] = rt->cwbvh.num_leaves;
    out[2] = rt->cwbvh.tri_order.size();
    out[3] = rt->cwbvh.depth;
    return RT_OK;
}
int rt_cwbvh_export(const rt_raytracer* rt_c, uint32_t* node_words, uint32_t* tri_order) {
    rt_raytracer* rt = const_cast<rt_raytracer*>(rt_c);
    if (!rt) return RT_ERR_INVALID;
    rt->ensure_cwbvh_host();
    if (node_words) std::memcpy(node_words, rt->cwbvh.nodes.data(), rt->cwbvh.nodes.size() * sizeof(CwWord));
    if (tri_order && !rt->cwbvh.tri_order.empty()) std::memcpy(tri_order, rt->cwbvh.tri_order.data(), rt->cwbvh.tri_order.size() * 4);
    return RT_OK;
}

}  // extern "C"
