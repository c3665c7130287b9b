// device_types.h — layouts shared by the host upload code and the sm_100a kernels.
//
// Everything a ray touches is packed for 16-byte (LDG.128) loads:
//   octree node   2 x float4 : (min.xyz | first_child or first triangle slot) (max.xyz | meta)
//   bvh node      4 x float4 : both child boxes + child references
//   triangle      3 x float4 : v0.xyz e1.x | e1.yz e2.xy | e2.z, global id, -, -        (e1 = v1-v0, e2 = v2-v0)
//   shade record  1 x float4 : unit normal xyz | geometry index
// e1/e2/normal are computed on the host with the same f32 operations the reference performs per ray
// (intersect.rs:66-67, mod.rs:198-205), so precomputing them is bit-identical.
#pragma once
#include <cstdint>

namespace rtb {

struct DevCamera {
    float rot[16];  // Camera::rotation_matrix
    float pos[3];   // orientation_matrix * (0,0,0,1)
    float max_x, max_y;
    uint32_t width, height;
};

struct DevTexture {
    const float* rgb;  // width*height*3
    uint32_t width, height;
};

constexpr uint32_t kOctLeafFlag = 0x80000000u;
constexpr uint32_t kNoHit = 0xFFFFFFFFu;
constexpr int kOctStack = 64;  // >= 7 * depth + 1 with depth <= 9 (oct_tree_intersector.rs:108)
constexpr int kBvhStack = 48;
constexpr int kGridLights = 4;  // point lights that can have a shadow-ray grid (more: their shadow rays walk the BVH)
constexpr int kBvh4Stack = 64;  // up to three pushes per level
constexpr int kCwStack = 32;  // node groups only: at most one per level plus slack

// Ray counters (u64 slots). Per-call counters exist twice (sets A and B): a trace call counts into the set its launch
// parameters name (TraceParams::counter_set) while the last warp out of every queue-driven launch zeroes the other set for the
// next call and puts the tile queue back to zero (warp_checkout, kernels.cu), so a call needs no memset. The *_TOTAL slots run
// since the handle was created (exact ray totals over many asynchronous calls, rt_get_ray_totals).
enum CounterSlot { CNT_SHADOW = 0, CNT_PRIMARY_HITS = 1, CNT_BOUNCE = 2, CNT_BLOCKED = 3,  // offsets inside a per-call set
                   CNT_SET_A = 0, CNT_TILE_QUEUE = 4, CNT_WARPS_DONE = 5, CNT_SHADOW_TOTAL = 6, CNT_BOUNCE_TOTAL = 7, CNT_SET_B = 8,
                   CNT_SLOTS = 12 };

// one level of the bounce wavefront: dense list of shaded hits (node record: hit point | pixel, normal | path,
// shaded radiance | parent + child number; before wf_shade_kernel: ray origin | pixel, direction | path, t u v | triangle,
// parent + child number) and, per node, the radiance its n_children bounce rays bring back
struct WfLevel {
    float4* rec;
    float* child_r;  // cap * n_children * 3
    uint32_t cap, n_children;
};
constexpr int kWfLevels = 5;    // recursions <= 4
constexpr int kWfRecWords = 4;  // float4 per node record (the 4th holds parent + child number while a hit waits for shading)

struct TraceParams {
    DevCamera cam;
    // acceleration structures (either may be null when not built)
    const float4* oct_nodes;
    const float4* oct_tris;
    const float4* bvh_nodes;
    uint32_t bvh_top_count;  // nodes [0, bvh_top_count) of bvh_nodes are in breadth-first order (host SAH tree; 0 for the GPU-built tree)
    const float4* bvh_tris;
    // perspective grid of the camera rays (pgrid_build.cu; instantiation ACCEL = 4 of the trace kernels): the (u, v) sample plane of
    // Camera::get_ray cut into square cells of 2^pg_shift pixels, per cell the slots of bvh_tris whose projection can reach it
    const uint32_t* pg_start;  // cell -> first entry of pg_tris; n_cells + 1 offsets
    const uint2* pg_tris;      // (slot of bvh_tris, smallest Z of the triangle as float bits), every list in ascending Z
    uint32_t pg_nx, pg_shift;
    // cube of perspective grids around every point light, for shadow rays (null = shadow rays walk the BVH): light li, face f (2 * axis +
    // (negative side)), cell (cy, cx) -> lg_start[((li * 6 + f) * lg_n + cy) * lg_n + cx]
    const uint32_t* lg_start;
    const uint2* lg_tris;      // (slot, lower bound of the triangle's distance from the light), every list in ascending distance
    uint32_t lg_n, lg_shift;   // cells per face edge; cell edge = 2^lg_shift grid units
    float lg_half;             // half a face edge in grid units
    uint32_t grid_lines[4];    // 128-byte lines of pg_start, pg_tris, lg_start, lg_tris (start-of-launch L2 prefetch)
    float lg_far2[kGridLights];  // |L|^2 from which a shadow ray of light li may reach surfaces lying beyond the light (it then walks the BVH)
    const float4* bvh4_nodes;  // 4-wide BVH, 8 float4 per node (bvh4_build.cpp)
    const float4* bvh4_tris;
    const uint4* cw_nodes;   // compressed 8-wide BVH, 5 words per node
    const float4* cw_tris;
    // shading data
    const float4* tri_shade;  // per global triangle
    const float4* materials;  // per geometry: rgb | texture id or -1
    const float4* lights;     // 2 per light: pos | color
    const DevTexture* textures;
    uint32_t num_lights;
    // film
    float4* film_sum;  // sum rgb | num_samples (uint bits)
    float4* film_sq;   // sum of squares rgb
    uint32_t* ldr;     // packed 0xAARRGGBB, kept current by the epilogue
    uint32_t* ldr_remote;  // optional second target (peer-mapped framebuffer of rank 0), may be null
    uint32_t* primary_ids;
    unsigned long long* counters;
    uint32_t counter_set;        // CNT_SET_A or CNT_SET_B: where this call's per-call ray counters live
    // multi-GPU fused gather: when non-null, the last warp out of the launch stores done_value here at system scope after
    // every warp's peer stores ("my stores of this frame are done", the signal half of the frame fence)
    uint32_t* done_flag;
    uint32_t done_value;
    const uint32_t* queue_items;  // number of entries of tile_order (written by tile_sort_kernel); unused when tile_order is null
    // cost-feedback tile schedule of the persistent kernel (either may be null): cycles spent per 8x4 tile in this
    // launch (written), queue slot -> tile id (read)
    uint32_t* tile_cost;
    const uint32_t* tile_order;
    // work: compact rows [0, n_rows) -> image row (first_row + c) % height, or row_list[c] when non-null
    const uint32_t* row_list;
    uint32_t first_row, n_rows;
    // sample planes: with planes != null the launch covers n_planes samples of plane_rows compact rows each
    // (n_rows = n_planes * plane_rows_padded); radiance and primitive id go to planes[(plane * plane_rows_padded + row) * width + col]
    // instead of the film, and film_accumulate_kernel adds them in sample order afterwards
    float4* planes;
    uint32_t n_planes, plane_rows, plane_rows_padded, magic_plane_rows;  // padded = plane_rows rounded up to 4
    // lap traced ahead of the band loop (raytracer.cu): per pixel, shadow rays (low byte) and bounce rays (high byte) of the sample
    // a lap build traced; the build launch writes it instead of the ray counters, the commit launch books it
    uint16_t* lap_rays;
    uint32_t film_prefetch;      // RT_TUNE_FILM_PREFETCH: 1 = ask L2 for a pixel's film record before its rays are traced, 2 = also for the sums of squares of a hit, 3 = also the whole tree at the start of the launch
    uint32_t film_prefetch_rows;  // 1 = the launch also requests the film records of all its rows up front (set by the host when they fit a share of L2)
    uint32_t bvh_node_lines, bvh_tri_lines, tri_shade_lines;  // 128-byte lines of bvh_nodes / bvh_tris / tri_shade (film_prefetch 3: the launch asks L2 for the whole tree first)
    uint32_t jitter_mode, seed;
    int32_t recursions;          // RECURSIONS (mod.rs:81); 0 = primary + shadow only
    uint32_t sub_spread;         // SUB_SPREAD (mod.rs:82)
    const float* sample_table;   // 65 536 unit vectors (sample_generator.rs), 3 floats each
    // bounce wavefront (null wf_counts = bounce rays are walked depth first inside the trace kernel)
    unsigned int* wf_counts;     // [0, L): nodes per level; [L, 2L): ray-queue head per level; [2L, 3L): shade-queue head (L = kWfLevels)
    uint32_t wf_level;           // level processed by wf_bounce_kernel / wf_combine_kernel
    WfLevel wf[kWfLevels];
    uint32_t wf_chain;           // ray-stream kernel: a lane continues in place with the single bounce ray of a node it completed
    uint32_t queue_batch, queue_batch_from_pct;  // persistent kernel: slots claimed at a time in the cheap tail of the sorted queue
    uint32_t pool_refill;        // ray-pool kernel: idle lanes of a warp that trigger a refill
    uint32_t pool_min_inner;     // ray-pool kernel: the inner-node loop yields when fewer lanes than this still descend
    uint32_t magic_w, magic_h, magic_tiles_x;  // floor(2^32 / d) for d = width, height, items per row (udiv_magic)
    // warp items of the persistent kernel: 2^item_cols_log2 columns x item_rows rows x 2^lane_samples_log2 samples = 32 lanes,
    // items_x * items_y items per launch (8 x 4 x 1 unless the lanes of an item share pixels, see finish_sample_lanes)
    uint32_t lane_samples_log2, item_cols_log2, item_rows, items_x, items_y;
    float root_lo[3], root_hi[3];  // scene AABB = octree root cube (acceptance rule of the BVH path)
};

}  // namespace rtb
