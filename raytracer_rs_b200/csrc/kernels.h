// kernels.h — launchers of the sm_100a kernels (defined in kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "device_types.h"

namespace rtb {

// fused primary ray -> closest hit -> shadow rays -> Phong/texture -> film -> tonemap/pack
// variant 0: one thread per pixel; variant 1: persistent warps pulling 8x4 tiles from an atomic queue;
// variant 2: ray pool, lanes decoupled from pixels through per-warp shared-memory ray rings
cudaError_t launch_trace(const TraceParams& p, int accel, int variant, int persistent_blocks, cudaStream_t stream);
int persistent_blocks_per_sm(int accel, int bounce);
// bounce wavefront: level p.wf_level -> level + 1 (one thread per (node, bounce ray)); bottom-up radiance combine
cudaError_t launch_wf_bounce(const TraceParams& p, int accel, int blocks, cudaStream_t stream);
cudaError_t launch_wf_shade(const TraceParams& p, int accel, int blocks, cudaStream_t stream);
// binary BVH: level p.wf_level -> level + 1 as a ray stream (bounce rays and the shadow rays of their hits share the lanes of a
// warp, finished lanes are refilled p.pool_refill at a time); replaces launch_wf_bounce + launch_wf_shade for that level
cudaError_t launch_wf_stream(const TraceParams& p, int blocks_per_sm /* 3, 4 or 5 */, int num_sms, cudaStream_t stream);
int wf_blocks_per_sm(int kind /* 0 wf_bounce_kernel, 1 wf_shade_kernel */, int accel);
cudaError_t launch_wf_combine(const TraceParams& p, int blocks, cudaStream_t stream);
// variant 2 (ray pool: binary BVH, recursions 0, one light; other configurations run variant 1)
int pool_blocks_per_sm();
// order[] = queue items by descending cost (longest-processing-time-first schedule for the persistent kernel). A tile that costs
// more than T = (balanced launch time) * split_quarters / 4 (and at least min_split_cycles) becomes 4, 8 or 16 items, the smallest
// count p with cost / p <= T, at most 2^(max_level + 1) (max_level 0 = never split): order must hold n * 2^(max_level + 1) entries
// (n when max_level = 0); *queue_items receives the item count. split_quarters 0 = never split (the octree's long coherent leaf
// loops do not profit).
cudaError_t launch_tile_sort(const uint32_t* cost, uint32_t* order, uint32_t n, uint32_t n_warps, uint32_t split_quarters, uint32_t max_level,
                             uint32_t min_split_cycles, uint32_t* queue_items, cudaStream_t stream);
cudaError_t launch_film_clear(float4* sum, float4* sq, uint32_t* ldr, uint32_t* ids, uint32_t n, cudaStream_t stream);
// Film::get_estimated_variances (film.rs:50-67): out = n * 3 floats
cudaError_t launch_film_variance(const float4* sum, const float4* sq, float* out, uint32_t n, cudaStream_t stream);
cudaError_t launch_tonemap(const float4* sum, uint32_t* ldr, uint32_t n, cudaStream_t stream);
// adds the sample planes written by a multi-sample trace launch to the film, in sample order, and packs the LDR pixels
cudaError_t launch_film_accumulate(const TraceParams& p, cudaStream_t stream);
cudaError_t launch_gather_rows(const uint32_t* ldr, const uint32_t* row_list, uint32_t n_rows, uint32_t width, uint32_t* out,
                               cudaStream_t stream);

// cross-GPU frame fence: *flag = value after everything earlier on the stream / wait until flags[0..n) >= target
// (bounded: a timeout increments *timeouts instead of hanging)
cudaError_t launch_flag_signal(uint32_t* flag, uint32_t value, cudaStream_t stream);
// *signal = value, then wait until *wait_flag >= target, in one launch
cudaError_t launch_flag_signal_wait(uint32_t* signal, uint32_t value, uint32_t* wait_flag, uint32_t target, uint32_t* timeouts, cudaStream_t stream);
// signal_slot / release_slot (-1 = none): flags[signal_slot] = target before the wait, flags[release_slot] = target after it
cudaError_t launch_flag_wait(uint32_t* flags, uint32_t n, uint32_t target, int signal_slot, int release_slot, uint32_t* timeouts,
                             cudaStream_t stream);
// GPU build of the binary BVH (lbvh_build.cu): Morton codes, radix sort, Karras hierarchy, bottom-up refit
size_t lbvh_scratch_bytes(uint32_t n);
cudaError_t build_lbvh_device(const float* d_verts, uint32_t n, const float root_lo[3], const float root_hi[3], void* scratch, float4* d_nodes,
                              float4* d_tris, uint32_t* d_tri_order, uint32_t* d_depth, cudaStream_t stream);

}  // namespace rtb
