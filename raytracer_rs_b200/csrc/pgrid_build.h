// pgrid_build.h — device build of the perspective grids (pgrid_build.cu): one frustum for the camera rays, six (a cube around the
// light) for the shadow rays of a point light.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace rtb {

struct PGridParams {
    const float4* tris;  // the binary BVH's triangle array, 3 float4 per slot (pack_triangle)
    uint32_t n_slots;
    uint32_t n_frusta;   // 1 (camera) or 6 (cube faces around a point)
    double A[6][9];      // per frustum: (X, Y, Z) = A (p - origin), row major; plane coordinates (U, V) = (X / Z, Y / Z) in grid units
    double origin[3];
    double z_eps;        // what lies nearer to the plane Z = 0 than this is clipped away (a vanishing fraction of the scene extent)
    uint32_t nx, ny;     // cells per frustum
    double cell;         // cell edge in grid units (a power of two); the margin around a projected triangle is one unit
    uint32_t cell_base;  // first cell of frustum 0 in count[] / cursor[] (frustum f starts at cell_base + f * nx * ny)
    uint32_t* count;
    uint32_t* cursor;
    uint2* entries;      // (triangle slot, key as float bits); unused by the counting pass
    uint32_t key_mode;   // key of an entry: 0 = smallest Z of the triangle's vertices (for a camera ray Z of a hit point = its t),
                         // 1 = distance from origin to the triangle's bounding box; both rounded down (lower bounds)
    uint32_t* big_queue; // [n_slots * n_frusta] footprints of more than 256 cells, queued by the counting pass (slot * 8 + frustum) ...
    uint32_t* big_count; // ... and their number; the fill pass of the same g reads both back
    float* dmin2;        // optional device word (start it at a huge value): min over triangles of the squared distance from origin to the
                         // triangle's bounding box, a lower bound of the distance to the nearest surface
};

// pass 1: adds the triangles of g to the per-cell counts (count[] zeroed by the caller); leaves g's queue of big footprints behind
cudaError_t pgrid_bin_count(const PGridParams& g, int num_sms, cudaStream_t stream);
// exclusive scan of count[0..n) -> start[0..n] (start[n] = *total) and cursor[0..n); n < 2^20; block_sums: 1024 words of scratch
cudaError_t pgrid_scan(const uint32_t* count, uint32_t* start, uint32_t* cursor, uint32_t n, uint32_t* total, uint32_t* block_sums, cudaStream_t stream);
// pass 2 (after the caller made entries[] large enough for *total; same g, its queue untouched since pass 1): the lists
cudaError_t pgrid_bin_fill(const PGridParams& g, int num_sms, cudaStream_t stream);
// pass 3: every cell's list in ascending key order (a walk can stop at the first key beyond its reach)
cudaError_t pgrid_sort_lists(const uint32_t* start, uint2* entries, uint32_t n_cells, cudaStream_t stream);

}  // namespace rtb
