// pgrid_build.cu — device build of the perspective grid of the camera rays (trace kernels, ACCEL = 4; kernels.cu, pgrid_closest_hit).
//
// Camera::get_ray (camera.rs:80-90) turns the pixel (u, v) and the sub-pixel offset (xi1, xi2) into the direction
//     dir = a * r0 - dir_y * r1 + r2 + r3,   a = -max_x + 2 max_x (u + xi1) / W,   dir_y = -max_y + 2 max_y (v + xi2) / H
// (r0..r3 the rows of the rotation matrix). Inverting that map, a world point p seen from the ray origin lands on the sample
// plane at (U, V) = (X / Z, Y / Z) with (X, Y, Z) = A (p - origin); the host folds the 3x3 inverse, max_x, max_y, W and H into A
// (raytracer.cu, ensure_pgrid). Pixel (u, v) owns [u, u + 1) x [v, v + 1) of that plane whatever its offsets are.
//
// Per triangle (one warp each): the three vertices go through A in binary64, the part in front of the eye (Z >= z_eps) is
// projected, and every cell the bounding box of the projection touches — widened by a margin of one pixel plus 1e-5 of the
// coordinate, orders of magnitude more than the rounding of the f32 Moller-Trumbore test can move a hit — lists the triangle.
// count -> exclusive scan -> fill; the lists are unordered (the closest-hit rule does not depend on the order of the tests).
#include <cuda_runtime.h>

#include <cstdint>

#include "pgrid_build.h"

namespace rtb {
namespace {

struct CellBox {
    int x0, x1, y0, y1;  // inclusive cell range; x0 > x1 = nothing
};

__device__ __forceinline__ CellBox triangle_cells(const PGridParams& g, const float4* tri) {
    const float4 t0 = tri[0], t1 = tri[1], t2 = tri[2];  // v0, e1 = v1 - v0, e2 = v2 - v0 (pack_triangle)
    const double vx[3] = {(double)t0.x, (double)t0.x + (double)t0.w, (double)t0.x + (double)t1.z};
    const double vy[3] = {(double)t0.y, (double)t0.y + (double)t1.x, (double)t0.y + (double)t1.w};
    const double vz[3] = {(double)t0.z, (double)t0.z + (double)t1.y, (double)t0.z + (double)t2.x};
    double X[3], Y[3], Z[3];
    for (int k = 0; k < 3; ++k) {
        const double wx = vx[k] - g.origin[0], wy = vy[k] - g.origin[1], wz = vz[k] - g.origin[2];
        X[k] = g.A[0] * wx + g.A[1] * wy + g.A[2] * wz;
        Y[k] = g.A[3] * wx + g.A[4] * wy + g.A[5] * wz;
        Z[k] = g.A[6] * wx + g.A[7] * wy + g.A[8] * wz;
    }
    double lo_u = 1e300, hi_u = -1e300, lo_v = 1e300, hi_v = -1e300;
    bool any = false;
    for (int k = 0; k < 3; ++k) {
        const int n = k == 2 ? 0 : k + 1;
        if (Z[k] >= g.z_eps) {
            const double u = X[k] / Z[k], v = Y[k] / Z[k];
            lo_u = fmin(lo_u, u), hi_u = fmax(hi_u, u), lo_v = fmin(lo_v, v), hi_v = fmax(hi_v, v);
            any = true;
        }
        if ((Z[k] >= g.z_eps) != (Z[n] >= g.z_eps)) {  // the edge crosses the plane Z = z_eps: its crossing point bounds the visible part
            const double s = (g.z_eps - Z[k]) / (Z[n] - Z[k]);
            const double u = (X[k] + s * (X[n] - X[k])) / g.z_eps, v = (Y[k] + s * (Y[n] - Y[k])) / g.z_eps;
            lo_u = fmin(lo_u, u), hi_u = fmax(hi_u, u), lo_v = fmin(lo_v, v), hi_v = fmax(hi_v, v);
            any = true;
        }
    }
    CellBox c = {1, 0, 1, 0};
    if (!any) return c;  // entirely behind the eye: no camera ray (t >= 0) reaches it
    const double mu = 1.0 + 1e-5 * fmax(fabs(lo_u), fabs(hi_u)), mv = 1.0 + 1e-5 * fmax(fabs(lo_v), fabs(hi_v));
    lo_u -= mu, hi_u += mu, lo_v -= mv, hi_v += mv;
    const double u_end = (double)g.nx * g.cell, v_end = (double)g.ny * g.cell;
    if (!(hi_u >= 0.0 && lo_u < u_end && hi_v >= 0.0 && lo_v < v_end)) return c;  // off the sample plane
    c.x0 = (int)(fmax(lo_u, 0.0) / g.cell);
    c.x1 = (int)(fmin(hi_u, u_end - 0.5) / g.cell);
    c.y0 = (int)(fmax(lo_v, 0.0) / g.cell);
    c.y1 = (int)(fmin(hi_v, v_end - 0.5) / g.cell);
    return c;
}

// one warp per triangle; FILL = false counts, FILL = true writes the entries (cursor[] starts as a copy of start[])
template <bool FILL>
__global__ void __launch_bounds__(256) pgrid_bin_kernel(const PGridParams g) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slot < g.n_slots; slot += warps) {
        const CellBox c = triangle_cells(g, g.tris + 3 * (size_t)slot);
        if (c.x0 > c.x1 || c.y0 > c.y1) continue;
        const uint32_t w = (uint32_t)(c.x1 - c.x0 + 1), n = w * (uint32_t)(c.y1 - c.y0 + 1);
        for (uint32_t k = lane; k < n; k += 32u) {
            const uint32_t cy = (uint32_t)c.y0 + k / w, cx = (uint32_t)c.x0 + k % w;
            const uint32_t cell = cy * g.nx + cx;
            if (FILL) g.entries[atomicAdd(&g.cursor[cell], 1u)] = slot;
            else atomicAdd(&g.count[cell], 1u);
        }
    }
}

// exclusive scan of count[0..n) into start[0..n], start[n] = total; one block (n is a few hundred thousand at most)
__global__ void __launch_bounds__(1024) pgrid_scan_kernel(const uint32_t* __restrict__ count, uint32_t* __restrict__ start,
                                                          uint32_t* __restrict__ cursor, uint32_t n, uint32_t* __restrict__ total) {
    __shared__ uint32_t part[1024];
    const uint32_t t = threadIdx.x, per = (n + 1023u) / 1024u;
    const uint32_t b = min(t * per, n), e = min(b + per, n);
    uint32_t s = 0;
    for (uint32_t i = b; i < e; ++i) s += count[i];
    part[t] = s;
    __syncthreads();
    for (uint32_t d = 1; d < 1024u; d <<= 1) {  // Hillis-Steele inclusive scan of the 1024 partial sums
        const uint32_t v = t >= d ? part[t - d] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - s;
    for (uint32_t i = b; i < e; ++i) {
        start[i] = run;
        cursor[i] = run;
        run += count[i];
    }
    if (t == 1023u) {
        start[n] = part[1023];
        *total = part[1023];
    }
}

}  // namespace

cudaError_t pgrid_count(const PGridParams& g, uint32_t n_cells, int num_sms, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(g.count, 0, (size_t)n_cells * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    const uint32_t blocks = (uint32_t)max(1, min(num_sms * 8, (int)((g.n_slots + 7u) / 8u)));
    pgrid_bin_kernel<false><<<blocks, 256, 0, stream>>>(g);
    pgrid_scan_kernel<<<1, 1024, 0, stream>>>(g.count, g.start, g.cursor, n_cells, g.total);
    return cudaGetLastError();
}
cudaError_t pgrid_fill(const PGridParams& g, int num_sms, cudaStream_t stream) {
    const uint32_t blocks = (uint32_t)max(1, min(num_sms * 8, (int)((g.n_slots + 7u) / 8u)));
    pgrid_bin_kernel<true><<<blocks, 256, 0, stream>>>(g);
    return cudaGetLastError();
}

}  // namespace rtb
