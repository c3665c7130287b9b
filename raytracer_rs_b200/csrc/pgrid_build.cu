// pgrid_build.cu — device build of the perspective grids (trace kernels, ACCEL = 4): the grid of the camera rays (kernels.cu,
// pgrid_closest_hit) and the cube of grids around a point light for the shadow rays (lgrid_shadow_blocked; the same binning with six
// frusta, A = the face's axes).
//
// Camera::get_ray (camera.rs:80-90) turns the pixel (u, v) and the sub-pixel offset (xi1, xi2) into the direction
//     dir = a * r0 - dir_y * r1 + r2 + r3,   a = -max_x + 2 max_x (u + xi1) / W,   dir_y = -max_y + 2 max_y (v + xi2) / H
// (r0..r3 the rows of the rotation matrix). Inverting that map, a world point p seen from the ray origin lands on the sample
// plane at (U, V) = (X / Z, Y / Z) with (X, Y, Z) = A (p - origin); the host folds the 3x3 inverse, max_x, max_y, W and H into A
// (raytracer.cu, ensure_pgrid). Pixel (u, v) owns [u, u + 1) x [v, v + 1) of that plane whatever its offsets are.
//
// Per triangle (one warp each): the three vertices go through A in binary64, the triangle is clipped against the frustum (almost always
// a trivial accept or reject), the rest is projected, and every cell the bounding box of the projection touches — widened by a margin
// of one grid unit, orders of magnitude more than the rounding of the f32 Moller-Trumbore test can move a hit — lists the triangle,
// unless the cell lies entirely beyond one edge of the projected triangle (conservative rasterisation, same margin). A footprint of
// more than 256 cells is queued and shared by a team of blocks in a second launch.
// count -> exclusive scan -> fill -> every list sorted by a key (the triangle's smallest Z for the camera, a lower bound of its distance
// for a light): the closest-hit rule does not depend on the order of the tests, and a walk can stop at the first key beyond its reach.
#include <cuda_runtime.h>

#include <cstdint>

#include "pgrid_build.h"

namespace rtb {
namespace {

struct CellBox {
    int x0, x1, y0, y1;  // inclusive cell range; x0 > x1 = nothing
};

struct HPoint {
    double x, y, z;  // plane coordinates (U, V) = (x / z, y / z)
};
// Sutherland-Hodgman step: keeps the part of the polygon with a x + b y + c z + d >= 0
__device__ __forceinline__ int clip_polygon(const HPoint* in, int n, HPoint* out, double a, double b, double c, double d) {
    int m = 0;
    for (int k = 0; k < n; ++k) {
        const HPoint p = in[k], q = in[k + 1 == n ? 0 : k + 1];
        const double fp = a * p.x + b * p.y + c * p.z + d, fq = a * q.x + b * q.y + c * q.z + d;
        if (fp >= 0.0) out[m++] = p;
        if ((fp >= 0.0) != (fq >= 0.0)) {
            const double s = fp / (fp - fq);
            out[m++] = HPoint{p.x + s * (q.x - p.x), p.y + s * (q.y - p.y), p.z + s * (q.z - p.z)};
        }
    }
    return m;
}

// Cells the triangle can reach inside one frustum: the triangle is clipped against the frustum (the plane Z = z_eps and the four sides,
// moved out by two grid units), what is left is projected, and the bounding box of the projection grows by the margin of one unit.
__device__ __forceinline__ CellBox triangle_cells(const PGridParams& g, const double* A, const float4* tri, double* z_min, float* uv /* [6], valid when the return's tight flag is set */,
                                              int* tight) {
    const float4 t0 = tri[0], t1 = tri[1], t2 = tri[2];  // v0, e1 = v1 - v0, e2 = v2 - v0 (pack_triangle)
    const double vx[3] = {(double)t0.x, (double)t0.x + (double)t0.w, (double)t0.x + (double)t1.z};
    const double vy[3] = {(double)t0.y, (double)t0.y + (double)t1.x, (double)t0.y + (double)t1.w};
    const double vz[3] = {(double)t0.z, (double)t0.z + (double)t1.y, (double)t0.z + (double)t2.x};
    HPoint pa[10], pb[10];
    for (int k = 0; k < 3; ++k) {
        const double wx = vx[k] - g.origin[0], wy = vy[k] - g.origin[1], wz = vz[k] - g.origin[2];
        pa[k].x = A[0] * wx + A[1] * wy + A[2] * wz;
        pa[k].y = A[3] * wx + A[4] * wy + A[5] * wz;
        pa[k].z = A[6] * wx + A[7] * wy + A[8] * wz;
    }
    *z_min = fmin(pa[0].z, fmin(pa[1].z, pa[2].z));
    const double u_end = (double)g.nx * g.cell, v_end = (double)g.ny * g.cell, m = 2.0;
    CellBox c = {1, 0, 1, 0};
    // the five planes of the frustum (Z >= z_eps: behind the eye no ray with t >= 0 arrives; the four sides moved out by m units). Almost every
    // triangle lies entirely on one side of every plane: all vertices outside one plane -> nothing to list; all inside all planes -> nothing to clip
    int outside_all = 0, outside_any = 0;
    for (int pl = 0; pl < 5; ++pl) {
        int out = 0;
        for (int k = 0; k < 3; ++k) {
            const double f = pl == 0   ? pa[k].z - g.z_eps
                             : pl == 1 ? pa[k].x + m * pa[k].z
                             : pl == 2 ? (u_end + m) * pa[k].z - pa[k].x
                             : pl == 3 ? pa[k].y + m * pa[k].z
                                       : (v_end + m) * pa[k].z - pa[k].y;
            out += f < 0.0;
        }
        outside_all |= out == 3;
        outside_any |= out != 0;
    }
    if (outside_all) return c;
    double lo_u = 1e300, hi_u = -1e300, lo_v = 1e300, hi_v = -1e300;
    *tight = 0;
    if (!outside_any) {
        for (int k = 0; k < 3; ++k) {
            const double u = pa[k].x / pa[k].z, v = pa[k].y / pa[k].z;
            lo_u = fmin(lo_u, u), hi_u = fmax(hi_u, u), lo_v = fmin(lo_v, v), hi_v = fmax(hi_v, v);
            uv[2 * k] = (float)u, uv[2 * k + 1] = (float)v;
        }
        *tight = 1;  // the whole triangle projects: its three edges can reject cells of the bounding box
    } else {
        int n = clip_polygon(pa, 3, pb, 0.0, 0.0, 1.0, -g.z_eps);
        if (n) n = clip_polygon(pb, n, pa, 1.0, 0.0, m, 0.0);           // U >= -m
        if (n) n = clip_polygon(pa, n, pb, -1.0, 0.0, u_end + m, 0.0);  // U <= u_end + m
        if (n) n = clip_polygon(pb, n, pa, 0.0, 1.0, m, 0.0);           // V >= -m
        if (n) n = clip_polygon(pa, n, pb, 0.0, -1.0, v_end + m, 0.0);  // V <= v_end + m
        if (n == 0) return c;
        for (int k = 0; k < n; ++k) {
            const double u = pb[k].x / pb[k].z, v = pb[k].y / pb[k].z;
            lo_u = fmin(lo_u, u), hi_u = fmax(hi_u, u), lo_v = fmin(lo_v, v), hi_v = fmax(hi_v, v);
        }
    }
    lo_u -= 1.0, hi_u += 1.0, lo_v -= 1.0, hi_v += 1.0;
    if (!(hi_u >= 0.0 && lo_u < u_end && hi_v >= 0.0 && lo_v < v_end)) return c;  // off the plane
    c.x0 = (int)(fmax(lo_u, 0.0) / g.cell);
    c.x1 = (int)(fmin(hi_u, u_end - 0.5) / g.cell);
    c.y0 = (int)(fmax(lo_v, 0.0) / g.cell);
    c.y1 = (int)(fmin(hi_v, v_end - 0.5) / g.cell);
    return c;
}

// What one (triangle, frustum) pair contributes: its cell box, the projected vertices when nothing was clipped, its key.
struct Footprint {
    int x0, x1, y0, y1;
    int tight;
    float key;
    float pu[3], pv[3];
};
// cells [first, first + step, ...) of the footprint's box (index k = row-major inside the box), for the lanes / threads of the caller
template <bool FILL>
__device__ __forceinline__ void bin_cells(const PGridParams& g, const Footprint& fp, uint32_t base, uint32_t slot, uint32_t first, uint32_t step) {
    const uint32_t w = (uint32_t)(fp.x1 - fp.x0 + 1), n = w * (uint32_t)(fp.y1 - fp.y0 + 1);
    // conservative rasterisation: a cell (grown by the margin of one unit and by the rounding of the f32 coordinates) that lies
    // entirely beyond one edge of the projected triangle cannot be reached by it
    const float area2 = (fp.pu[1] - fp.pu[0]) * (fp.pv[2] - fp.pv[0]) - (fp.pu[2] - fp.pu[0]) * (fp.pv[1] - fp.pv[0]);
    const bool edges = fp.tight && n > 1u && fabsf(area2) > 1e-3f;
    const float sgn = area2 < 0.f ? -1.f : 1.f, cs = (float)g.cell;
    for (uint32_t k = first; k < n; k += step) {
        const uint32_t cy = (uint32_t)fp.y0 + k / w, cx = (uint32_t)fp.x0 + k % w;
        if (edges) {
            const float grow = 1.0f + 2e-3f * cs + 1e-6f * ((float)(cx + cy + 2u) * cs);
            const float xl = (float)cx * cs - grow, xh = (float)(cx + 1u) * cs + grow, yl = (float)cy * cs - grow, yh = (float)(cy + 1u) * cs + grow;
            bool reach = true;
            for (int i = 0; i < 3; ++i) {
                const int j = i == 2 ? 0 : i + 1;
                const float a = sgn * (fp.pu[j] - fp.pu[i]), b = sgn * (fp.pv[j] - fp.pv[i]);  // inside: a (y - v_i) - b (x - u_i) >= 0
                const float best = a * ((a > 0.f ? yh : yl) - fp.pv[i]) - b * ((b > 0.f ? xl : xh) - fp.pu[i]);
                // (the products round: keep a cell unless it is beyond the edge by more than that)
                reach = reach && best >= -1e-4f * (fabsf(a) + fabsf(b)) * (fabsf(xh) + fabsf(yh) + fabsf(fp.pu[i]) + fabsf(fp.pv[i]));
            }
            if (!reach) continue;
        }
        const uint32_t cell = base + cy * g.nx + cx;
        if (FILL) g.entries[atomicAdd(&g.cursor[cell], 1u)] = make_uint2(slot, __float_as_uint(fp.key));
        else atomicAdd(&g.count[cell], 1u);
    }
}
// distance from the origin to the triangle's bounding box, squared: a lower bound of the distance to the triangle
__device__ __forceinline__ double box_distance2(const PGridParams& g, const float4* tri) {
    const float4 t0 = tri[0], t1 = tri[1], t2 = tri[2];
    const double v0[3] = {t0.x, t0.y, t0.z}, e1[3] = {t0.w, t1.x, t1.y}, e2[3] = {t1.z, t1.w, t2.x};
    double d2 = 0.0;
    for (int a = 0; a < 3; ++a) {
        const double lo = v0[a] + fmin(0.0, fmin(e1[a], e2[a])), hi = v0[a] + fmax(0.0, fmax(e1[a], e2[a]));
        const double d = fmax(0.0, fmax(lo - g.origin[a], g.origin[a] - hi));
        d2 += d * d;
    }
    return d2;
}
__device__ __forceinline__ Footprint footprint_of(const PGridParams& g, uint32_t f, const float4* tri) {
    Footprint fp;
    double z_min;
    float uv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const CellBox c = triangle_cells(g, g.A[f], tri, &z_min, uv, &fp.tight);
    fp.x0 = c.x0, fp.x1 = c.x1, fp.y0 = c.y0, fp.y1 = c.y1;
    fp.key = g.key_mode == 1u ? __double2float_rd(sqrt(box_distance2(g, tri)) * (1.0 - 1e-12)) : __double2float_rd(z_min);
    for (int k = 0; k < 3; ++k) fp.pu[k] = uv[2 * k], fp.pv[k] = uv[2 * k + 1];
    return fp;
}

constexpr uint32_t kBigCells = 256;  // a footprint of more cells than this is left to pgrid_big_kernel (all blocks share its cells)

// One warp per triangle. Lane f works out the footprint in frustum f (binary64 clipping and projection: once per frustum, not once per
// lane), the warp then walks each frustum's cells together. FILL = false counts and queues the big footprints, FILL = true writes the
// entries (cursor[] starts as a copy of start[]).
template <bool FILL>
__global__ void __launch_bounds__(256) pgrid_bin_kernel(const PGridParams g) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slot < g.n_slots; slot += warps) {
        const float4* tri = g.tris + 3 * (size_t)slot;
        Footprint mine;
        mine.x0 = 1, mine.x1 = 0, mine.y0 = 1, mine.y1 = 0, mine.tight = 0, mine.key = 0.f;
        for (int k = 0; k < 3; ++k) mine.pu[k] = mine.pv[k] = 0.f;
        if (lane < g.n_frusta) mine = footprint_of(g, lane, tri);
        if (!FILL && g.dmin2 && lane == 0u)  // (rounded down; non-negative floats order like their bit patterns)
            atomicMin(reinterpret_cast<unsigned int*>(g.dmin2), __float_as_uint(__double2float_rd(box_distance2(g, tri) * (1.0 - 1e-12))));
        for (uint32_t f = 0; f < g.n_frusta; ++f) {
            Footprint fp;
            fp.x0 = __shfl_sync(0xffffffffu, mine.x0, f), fp.x1 = __shfl_sync(0xffffffffu, mine.x1, f);
            fp.y0 = __shfl_sync(0xffffffffu, mine.y0, f), fp.y1 = __shfl_sync(0xffffffffu, mine.y1, f);
            fp.key = __shfl_sync(0xffffffffu, mine.key, f);
            fp.tight = __shfl_sync(0xffffffffu, mine.tight, f);
            for (int k = 0; k < 3; ++k) fp.pu[k] = __shfl_sync(0xffffffffu, mine.pu[k], f), fp.pv[k] = __shfl_sync(0xffffffffu, mine.pv[k], f);
            if (fp.x0 > fp.x1 || fp.y0 > fp.y1) continue;
            const uint32_t n = (uint32_t)(fp.x1 - fp.x0 + 1) * (uint32_t)(fp.y1 - fp.y0 + 1);
            if (n > kBigCells) {  // one warp would walk thousands of cells while the rest of the GPU has finished
                if (!FILL && lane == 0u) g.big_queue[atomicAdd(g.big_count, 1u)] = slot * 8u + f;
                continue;
            }
            bin_cells<FILL>(g, fp, g.cell_base + f * g.nx * g.ny, slot, lane, 32u);
        }
    }
}
// the big footprints queued by the counting pass: the blocks split into one team per footprint (when there are more footprints than
// blocks a block takes several in turn), and a team shares its footprint's cells — no block waits for a footprint it does not work on
template <bool FILL>
__global__ void __launch_bounds__(256) pgrid_big_kernel(const PGridParams g) {
    __shared__ Footprint fp;
    const uint32_t nq = *g.big_count;
    if (nq == 0u) return;
    const uint32_t team_size = max(1u, gridDim.x / nq);      // blocks per footprint
    const uint32_t teams = min(nq, gridDim.x);                // footprints in flight at once
    const uint32_t team = blockIdx.x % teams, member = blockIdx.x / teams;
    if (member >= team_size) return;
    for (uint32_t q = team; q < nq; q += teams) {
        const uint32_t item = g.big_queue[q], slot = item >> 3, f = item & 7u;
        if (threadIdx.x == 0u) fp = footprint_of(g, f, g.tris + 3 * (size_t)slot);
        __syncthreads();
        const Footprint local = fp;
        bin_cells<FILL>(g, local, g.cell_base + f * g.nx * g.ny, slot, member * blockDim.x + threadIdx.x, team_size * blockDim.x);
        __syncthreads();
    }
}

// One warp per cell: the list goes through shared memory (up to 128 entries, bitonic network on (key, slot) — the slot as second key
// makes the order independent of the timing of the fill's atomics). A longer list stays as it is and gets keys that never stop a walk.
constexpr uint32_t kSortMax = 128;
__global__ void __launch_bounds__(256) pgrid_sort_kernel(const uint32_t* __restrict__ start, uint2* __restrict__ entries, uint32_t n_cells) {
    __shared__ unsigned long long buf[8][kSortMax];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; cell < n_cells; cell += warps) {
        const uint32_t b = start[cell], n = start[cell + 1u] - b;
        if (n < 2u) continue;
        if (n > kSortMax) {
            for (uint32_t i = lane; i < n; i += 32u) entries[b + i].y = 0xff7fffffu;  // -FLT_MAX
            continue;
        }
        uint32_t m = 2u;
        while (m < n) m <<= 1;
        for (uint32_t i = lane; i < m; i += 32u) {
            unsigned long long v = ~0ull;  // padding sorts to the end
            if (i < n) {
                const uint2 e = entries[b + i];
                // order-preserving map of the float key to an unsigned integer (keys can be negative: a triangle partly behind the eye)
                const uint32_t k = (e.y & 0x80000000u) ? ~e.y : (e.y | 0x80000000u);
                v = ((unsigned long long)k << 32) | e.x;
            }
            buf[w][i] = v;
        }
        __syncwarp();
        for (uint32_t k = 2u; k <= m; k <<= 1)
            for (uint32_t j = k >> 1; j > 0u; j >>= 1) {
                for (uint32_t i = lane; i < m; i += 32u) {
                    const uint32_t p = i ^ j;
                    if (p > i) {
                        const unsigned long long x = buf[w][i], y = buf[w][p];
                        const bool up = (i & k) == 0u;
                        if ((x > y) == up) {
                            buf[w][i] = y;
                            buf[w][p] = x;
                        }
                    }
                }
                __syncwarp();
            }
        for (uint32_t i = lane; i < n; i += 32u) {
            const unsigned long long v = buf[w][i];
            const uint32_t k = (uint32_t)(v >> 32);
            entries[b + i] = make_uint2((uint32_t)v, (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
        }
        __syncwarp();
    }
}

// exclusive scan of count[0..n) into start[0..n] (start[n] = total) and cursor[0..n), three launches: sums of 1024-cell blocks, a scan of
// those sums in one block (n <= 2^20 cells), the scan inside every block plus its offset
__device__ __forceinline__ uint32_t block_scan_1024(uint32_t v, uint32_t* warp_sums /* [32] shared */, uint32_t* block_total) {
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    uint32_t x = v;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    if (lane == 31u) warp_sums[w] = x;
    __syncthreads();
    if (w == 0u) {
        uint32_t s = warp_sums[lane];
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= (uint32_t)d) s += y;
        }
        warp_sums[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    *block_total = warp_sums[31];
    return x - v + (w ? warp_sums[w - 1u] : 0u);  // exclusive prefix of this thread
}
__global__ void __launch_bounds__(1024) pgrid_block_sums_kernel(const uint32_t* __restrict__ count, uint32_t* __restrict__ sums, uint32_t n) {
    __shared__ uint32_t ws[32];
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
    uint32_t total;
    block_scan_1024(i < n ? count[i] : 0u, ws, &total);
    if (threadIdx.x == 0u) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) pgrid_scan_sums_kernel(uint32_t* __restrict__ sums, uint32_t nb, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t ws[32];
    uint32_t total;
    const uint32_t v = threadIdx.x < nb ? sums[threadIdx.x] : 0u;
    const uint32_t ex = block_scan_1024(v, ws, &total);
    if (threadIdx.x < nb) sums[threadIdx.x] = ex;
    if (threadIdx.x == 0u) *total_out = total;
}
__global__ void __launch_bounds__(1024) pgrid_scan_apply_kernel(const uint32_t* __restrict__ count, const uint32_t* __restrict__ sums,
                                                                const uint32_t* __restrict__ total, uint32_t* __restrict__ start,
                                                                uint32_t* __restrict__ cursor, uint32_t n) {
    __shared__ uint32_t ws[32];
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
    uint32_t block_total;
    const uint32_t ex = block_scan_1024(i < n ? count[i] : 0u, ws, &block_total) + sums[blockIdx.x];
    if (i < n) {
        start[i] = ex;
        cursor[i] = ex;
    }
    if (i == n) start[n] = *total;
}

}  // namespace

cudaError_t pgrid_bin_count(const PGridParams& g, int num_sms, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(g.big_count, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    const uint32_t blocks = (uint32_t)max(1, min(num_sms * 8, (int)((g.n_slots + 7u) / 8u)));
    pgrid_bin_kernel<false><<<blocks, 256, 0, stream>>>(g);
    pgrid_big_kernel<false><<<num_sms * 2, 256, 0, stream>>>(g);
    return cudaGetLastError();
}
cudaError_t pgrid_scan(const uint32_t* count, uint32_t* start, uint32_t* cursor, uint32_t n, uint32_t* total, uint32_t* block_sums, cudaStream_t stream) {
    const uint32_t nb = n / 1024u + 1u;  // one more than needed when 1024 divides n: some block has to hold i == n
    if (nb > 1024u) return cudaErrorInvalidValue;
    pgrid_block_sums_kernel<<<nb, 1024, 0, stream>>>(count, block_sums, n);
    pgrid_scan_sums_kernel<<<1, 1024, 0, stream>>>(block_sums, nb, total);
    pgrid_scan_apply_kernel<<<nb, 1024, 0, stream>>>(count, block_sums, total, start, cursor, n);
    return cudaGetLastError();
}
cudaError_t pgrid_bin_fill(const PGridParams& g, int num_sms, cudaStream_t stream) {
    const uint32_t blocks = (uint32_t)max(1, min(num_sms * 8, (int)((g.n_slots + 7u) / 8u)));
    pgrid_bin_kernel<true><<<blocks, 256, 0, stream>>>(g);
    pgrid_big_kernel<true><<<num_sms * 2, 256, 0, stream>>>(g);  // the queue the counting pass of the same g left behind
    return cudaGetLastError();
}

cudaError_t pgrid_sort_lists(const uint32_t* start, uint2* entries, uint32_t n_cells, cudaStream_t stream) {
    pgrid_sort_kernel<<<min((n_cells + 7u) / 8u, 148u * 8u), 256, 0, stream>>>(start, entries, n_cells);
    return cudaGetLastError();
}

}  // namespace rtb
