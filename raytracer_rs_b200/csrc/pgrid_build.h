// pgrid_build.h — device build of the perspective grid of the camera rays (pgrid_build.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace rtb {

struct PGridParams {
    const float4* tris;  // the binary BVH's triangle array, 3 float4 per slot (pack_triangle)
    uint32_t n_slots;
    double A[9];         // (X, Y, Z) = A (p - origin), row major; sample-plane coordinates (U, V) = (X / Z, Y / Z) in pixels
    double origin[3];
    double z_eps;        // what lies nearer to the eye plane than this is clipped away (a vanishing fraction of the scene extent)
    uint32_t nx, ny;     // cells
    double cell;         // cell edge in pixels (a power of two)
    uint32_t* count;     // [nx * ny]
    uint32_t* start;     // [nx * ny + 1]
    uint32_t* cursor;    // [nx * ny]
    uint32_t* entries;   // [capacity]; unused by pgrid_count
    uint32_t* total;     // device word: number of entries
};

// pass 1: per-cell counts and their exclusive scan (start[], cursor[], *total)
cudaError_t pgrid_count(const PGridParams& g, uint32_t n_cells, int num_sms, cudaStream_t stream);
// pass 2 (after the caller made entries[] large enough for *total): the lists
cudaError_t pgrid_fill(const PGridParams& g, int num_sms, cudaStream_t stream);

}  // namespace rtb
