// kernels.cu — hand-written sm_100a kernels for the per-pixel render loop.
//
// One fused kernel per launch: camera ray (camera.rs:80-90) -> closest hit (octree: oct_tree_intersector.rs:148-272,
// or BVH) -> normal -> per light: facing test, shadow closest hit, Phong + texture (mod.rs:207-261) -> film add
// (film.rs:20-24) -> mean, tonemap, pack (film.rs:43-48, tonemap.rs:4-10, color.rs:89-95) stored to the LDR frame.
//
// Numerics: every operation whose result can change a pixel uses the explicit round-to-nearest intrinsics
// (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/__fsqrt_rn). They are never contracted into FMAs, so the results are
// bit-identical to the reference's Rust f32 arithmetic (and to oracle/rt_oracle.cpp) regardless of compiler flags.
// BVH box tests are allowed to use FMAs because the boxes are padded and only decide which triangles get tested.
//
// Kernels in this file:
//   trace_shade_persistent_kernel<ACCEL, BOUNCE, SAMPLE_LANES, LAP>
//                                                 the default: persistent warps, cost-sorted 8x4 tile queue (variant 1), one launch per trace
//                                                 call and nothing else (the last warp out resets queue and counters, warp_checkout);
//                                                 SAMPLE_LANES: multi-sample launches, the lanes of an item hold the samples of a pixel;
//                                                 LAP: traces a lap ahead of the band loop into a frame-aligned sample plane
//   trace_shade_kernel<ACCEL, BOUNCE>             one thread per pixel (variant 0, kept for A/B runs)
//   trace_shade_pool_kernel                       ray pool with shared-memory ray rings (variant 2)
//   wf_stream_kernel<MIN_BLOCKS>                  bounce wavefront (RECURSIONS > 0) on the binary BVH as a ray stream: lanes decoupled from
//                                                 rays, hits shaded in place, bounce levels chained in one launch
//   wf_bounce_kernel<ACCEL>, wf_shade_kernel<ACCEL>   the lockstep form of a bounce level (other structures)
//   wf_combine_kernel                             bottom-up radiance combine of the bounce tree; level 0 adds to the film
//   film_accumulate_kernel                        ordered accumulation of sample planes (odd spp) / commit of a band of the lap plane
//   film_variance_kernel                          Film::get_estimated_variances (film.rs:50-67)
//   tile_sort_kernel                              heaviest-first order of the tile queue from last launch's tile costs; heavy tiles
//                                                 become 4, 8 or 16 items
//   flag_signal_kernel, flag_signal_wait_kernel, flag_wait_kernel    cross-GPU frame fence of the fused peer-store gather
//   film_clear_kernel, tonemap_pack_kernel, gather_rows_kernel
// ACCEL: 0 exact octree, 1 binary BVH (host SAH or GPU LBVH), 2 compressed 8-wide BVH, 3 4-wide BVH.
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "device_types.h"
#include "kernels.h"

namespace rtb {

// ------------------------------------------------------------------------------------------------------
// exact f32 helpers
// ------------------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return {fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)}; }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return {fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)}; }
__device__ __forceinline__ V3 vscale(V3 a, float s) { return {fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)}; }
__device__ __forceinline__ float vdot(V3 a, V3 b) {  // vecmath.rs:74-76, left to right
    return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z));
}
__device__ __forceinline__ V3 vcross(V3 a, V3 b) {  // vecmath.rs:79-85
    return {fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)), fsub(fmul(a.x, b.y), fmul(a.y, b.x))};
}
__device__ __forceinline__ V3 vunit(V3 v) {  // Vec3::normalized, vecmath.rs:23-26
    const float len = __fsqrt_rn(fadd(fadd(fmul(v.x, v.x), fmul(v.y, v.y)), fmul(v.z, v.z)));
    return {fdiv(v.x, len), fdiv(v.y, len), fdiv(v.z, len)};
}

// MUFU.RCP (~1 ulp). Only for values that feed the conservative, padded box tests — never for anything that reaches a pixel.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct HitRec {
    float t, u, v;
    uint32_t tri;  // global triangle index
};

// counter-based generator shared with the oracle (rt_oracle.cpp: mix32/hash4/u01)
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x7feb352dU;
    h ^= h >> 15;
    h *= 0x846ca68bU;
    h ^= h >> 16;
    return h;
}
__device__ __forceinline__ uint32_t hash4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t h = mix32(a + 0x9e3779b9U);
    h = mix32(h ^ (b + 0x85ebca6bU));
    h = mix32(h ^ (c + 0xc2b2ae35U));
    h = mix32(h ^ (d + 0x27d4eb2fU));
    return h;
}
__device__ __forceinline__ float u01(uint32_t h) { return fmul((float)(h >> 8), 1.0f / 16777216.0f); }

// ------------------------------------------------------------------------------------------------------
// Moller-Trumbore, "late out" variant (intersect.rs:62-98) on a packed triangle record
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool moller_trumbore(const V3& o, const V3& d, const float4 t0, const float4 t1, const float4 t2, float* t,
                                                float* u, float* v) {
    const V3 v0 = {t0.x, t0.y, t0.z};
    const V3 e1 = {t0.w, t1.x, t1.y};
    const V3 e2 = {t1.z, t1.w, t2.x};
    const V3 pvec = vcross(d, e2);
    const float det = vdot(e1, pvec);
    if (fabsf(det) < FLT_EPSILON) return false;
    const float inv_det = fdiv(1.0f, det);
    const V3 tvec = vsub(o, v0);
    const float uu = fmul(vdot(tvec, pvec), inv_det);
    const V3 qvec = vcross(tvec, e1);
    const float vv = fmul(vdot(d, qvec), inv_det);
    const float tt = fmul(vdot(e2, qvec), inv_det);
    if (uu < 0.0f || uu > 1.0f) return false;
    if (vv < 0.0f || fadd(uu, vv) > 1.0f) return false;
    if (tt < 0.0f) return false;
    *t = tt;
    *u = uu;
    *v = vv;
    return true;
}

// ------------------------------------------------------------------------------------------------------
// exact octree traversal (oct_tree_intersector.rs:148-196, 240-272, 348-372)
//
// The recursion becomes an explicit stack: an inner node slab-tests its 8 children, and pushes the hit ones in
// DESCENDING (t, child) order so they pop in the stable ascending order of the reference's sort_by. The 8 child
// cubes are not loaded: they are (min | mid | max) selections of the parent's cube, with
// mid = 0.5*(max+min) evaluated exactly as generate_child_cubes (:275) did when the stored cubes were built, so
// the 48 subtractions/multiplications of the reference collapse to 9 with identical values.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool octree_closest_hit(const TraceParams& P, const V3& o, const V3& d, HitRec* out) {
    const V3 inv = {fdiv(1.0f, d.x), fdiv(1.0f, d.y), fdiv(1.0f, d.z)};  // :241-244
    int stack[kOctStack];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const int node = stack[--sp];
        const float4 A = __ldg(&P.oct_nodes[2 * node]);
        const float4 B = __ldg(&P.oct_nodes[2 * node + 1]);
        const uint32_t meta = __float_as_uint(B.w);
        if (meta & kOctLeafFlag) {
            // ---- leaf: closest triangle of the list (strict <, first wins ties), then the in-cube test ----
            const uint32_t count = meta & ~kOctLeafFlag;
            const float4* tri = P.oct_tris + 3 * (size_t)__float_as_uint(A.w);
            bool have = false;
            HitRec best;
            best.t = 0.f;
            best.u = 0.f;
            best.v = 0.f;
            best.tri = kNoHit;
            for (uint32_t i = 0; i < count; ++i) {
                const float4 t0 = __ldg(tri + 3 * i), t1 = __ldg(tri + 3 * i + 1), t2 = __ldg(tri + 3 * i + 2);
                float t, u, v;
                if (!moller_trumbore(o, d, t0, t1, t2, &t, &u, &v)) continue;
                if (!have || t < best.t) {
                    have = true;
                    best.t = t;
                    best.u = u;
                    best.v = v;
                    best.tri = __float_as_uint(t2.y);
                }
            }
            if (have) {
                const V3 hp = vadd(o, vscale(d, best.t));  // ray.pos + ray.dir * t  (:164)
                const bool outside = hp.x < A.x || hp.x > B.x || hp.y < A.y || hp.y > B.y || hp.z < A.z || hp.z > B.z;
                if (!outside) {
                    *out = best;
                    return true;
                }
            }
            continue;
        }
        // ---- inner node ----
        const int first_child = (int)__float_as_uint(A.w);
        uint32_t live = meta & 0xffu;  // children that hold at least one triangle (empty leaves return None at once)
        const float midx = fmul(0.5f, fadd(B.x, A.x)), midy = fmul(0.5f, fadd(B.y, A.y)), midz = fmul(0.5f, fadd(B.z, A.z));
        const float xl = fmul(fsub(A.x, o.x), inv.x), xm = fmul(fsub(midx, o.x), inv.x), xh = fmul(fsub(B.x, o.x), inv.x);
        const float yl = fmul(fsub(A.y, o.y), inv.y), ym = fmul(fsub(midy, o.y), inv.y), yh = fmul(fsub(B.y, o.y), inv.y);
        const float zl = fmul(fsub(A.z, o.z), inv.z), zm = fmul(fsub(midz, o.z), inv.z), zh = fmul(fsub(B.z, o.z), inv.z);
        const float nx[2] = {fminf(xl, xm), fminf(xm, xh)}, fx[2] = {fmaxf(xl, xm), fmaxf(xm, xh)};
        const float ny[2] = {fminf(yl, ym), fminf(ym, yh)}, fy[2] = {fmaxf(yl, ym), fmaxf(ym, yh)};
        const float nz[2] = {fminf(zl, zm), fminf(zm, zh)}, fz[2] = {fmaxf(zl, zm), fmaxf(zm, zh)};
        float tc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float tmin = fmaxf(fmaxf(nx[c & 1], ny[(c >> 1) & 1]), nz[c >> 2]);
            const float tmax = fminf(fminf(fx[c & 1], fy[(c >> 1) & 1]), fz[c >> 2]);
            tc[c] = tmin;
            if (!(tmax >= tmin && tmax > 0.0f)) live &= ~(1u << c);
        }
        while (live) {
            int pick = -1;
            float pt = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if ((live >> c) & 1u) {
                    if (pick < 0 || tc[c] >= pt) {
                        pick = c;
                        pt = tc[c];
                    }
                }
            }
            live &= ~(1u << pick);
            stack[sp++] = first_child + pick;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------------
// BVH traversal: true closest hit with the reference's tie rule (lowest global triangle index wins equal t,
// oct_tree_intersector.rs:258-268 + :332-342) followed by the root-cube acceptance rule (:164-169 applied to
// the scene AABB). `t_limit` (exclusive) clips the search; `early_t`: any hit with t <= early_t ends
// the search at once (shadow rays: such a hit makes the point lit whatever lies beyond, mod.rs:226-230).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool bvh_closest_hit(const TraceParams& P, const V3& o, const V3& d, float t_limit, float early_t, HitRec* out) {
    const float ix = fdiv(1.0f, d.x), iy = fdiv(1.0f, d.y), iz = fdiv(1.0f, d.z);
    const float ox = -o.x * ix, oy = -o.y * iy, oz = -o.z * iz;
    int stack_node[kBvhStack];
    float stack_t[kBvhStack];
    int sp = 0;
    HitRec best;
    best.t = t_limit;
    best.u = 0.f;
    best.v = 0.f;
    best.tri = kNoHit;
    int cur = 0;
    for (;;) {
        if (cur >= 0) {
            const float4* n = P.bvh_nodes + 4 * (size_t)cur;
            const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
            // child 0: lo = (q0.x q0.y q0.z) hi = (q0.w q1.x q1.y); child 1: lo = (q1.z q1.w q2.x) hi = (q2.y q2.z q2.w)
            const float a0x = fmaf(q0.x, ix, ox), b0x = fmaf(q0.w, ix, ox);
            const float a0y = fmaf(q0.y, iy, oy), b0y = fmaf(q1.x, iy, oy);
            const float a0z = fmaf(q0.z, iz, oz), b0z = fmaf(q1.y, iz, oz);
            const float a1x = fmaf(q1.z, ix, ox), b1x = fmaf(q2.y, ix, ox);
            const float a1y = fmaf(q1.w, iy, oy), b1y = fmaf(q2.z, iy, oy);
            const float a1z = fmaf(q2.x, iz, oz), b1z = fmaf(q2.w, iz, oz);
            const float n0 = fmaxf(fmaxf(fminf(a0x, b0x), fminf(a0y, b0y)), fmaxf(fminf(a0z, b0z), 0.0f));
            const float f0 = fminf(fminf(fmaxf(a0x, b0x), fmaxf(a0y, b0y)), fminf(fmaxf(a0z, b0z), best.t));
            const float n1 = fmaxf(fmaxf(fminf(a1x, b1x), fminf(a1y, b1y)), fmaxf(fminf(a1z, b1z), 0.0f));
            const float f1 = fminf(fminf(fmaxf(a1x, b1x), fmaxf(a1y, b1y)), fminf(fmaxf(a1z, b1z), best.t));
            const bool h0 = n0 <= f0, h1 = n1 <= f1;
            const int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
            if (h0 && h1) {
                const bool first0 = n0 <= n1;
                stack_node[sp] = first0 ? c1 : c0;
                stack_t[sp] = first0 ? n1 : n0;
                ++sp;
                cur = first0 ? c0 : c1;
                continue;
            }
            if (h0) {
                cur = c0;
                continue;
            }
            if (h1) {
                cur = c1;
                continue;
            }
        } else {
            const uint32_t ref = (uint32_t)~cur;
            const uint32_t count = ref & 15u;
            const float4* tri = P.bvh_tris + 3 * (size_t)(ref >> 4);
            for (uint32_t i = 0; i < count; ++i) {
                const float4 t0 = __ldg(tri + 3 * i), t1 = __ldg(tri + 3 * i + 1), t2 = __ldg(tri + 3 * i + 2);
                float t, u, v;
                if (!moller_trumbore(o, d, t0, t1, t2, &t, &u, &v)) continue;
                const uint32_t id = __float_as_uint(t2.y);
                if (t < best.t || (t == best.t && id < best.tri)) {
                    best.t = t;
                    best.u = u;
                    best.v = v;
                    best.tri = id;
                    if (t <= early_t) {
                        *out = best;
                        return true;
                    }
                }
            }
        }
        // pop, skipping subtrees that start beyond the current closest hit
        for (;;) {
            if (sp == 0) goto done;
            --sp;
            if (stack_t[sp] <= best.t) {
                cur = stack_node[sp];
                break;
            }
        }
    }
done:
    if (best.tri == kNoHit) return false;
    const V3 hp = vadd(o, vscale(d, best.t));
    const bool outside = hp.x < P.root_lo[0] || hp.x > P.root_hi[0] || hp.y < P.root_lo[1] || hp.y > P.root_hi[1] || hp.z < P.root_lo[2] ||
                         hp.z > P.root_hi[2];
    if (outside) return false;
    *out = best;
    return true;
}


// ------------------------------------------------------------------------------------------------------
// "while-while" forms of the two traversals (same results, different control flow): the inner loop keeps
// descending inner nodes until EVERY lane of the warp holds a leaf (or is finished), then all lanes test triangles
// together. Lanes no longer alternate between node code and triangle code inside one loop body, which is what
// made only ~11 of 32 lanes useful per instruction in the first version (profiles/r1_v1_*.csv).
// ------------------------------------------------------------------------------------------------------
constexpr int kSentinel = 0x7fffffff;
// queue item = tile id (bits 0..23) | part (bits 24..27) | split level (bits 28..29): level 0 = the whole 32-lane tile,
// level L > 0 = one of 2^(L+1) parts (4, 8 or 16) of 2^(4-L) consecutive lanes (8, 4 or 2 pixels of a single-sample tile)
constexpr uint32_t kItemTileMask = 0x00ffffffu;
constexpr int kItemPartShift = 24, kItemLevelShift = 28;
#ifdef RT_DEBUG_STEP_COUNTS
#define g_dbg_nodes (*dbg_nodes_ptr())
#define g_dbg_tris (*dbg_tris_ptr())
__device__ __forceinline__ uint32_t* dbg_nodes_ptr() { extern __shared__ uint32_t dbg_sm[]; return &dbg_sm[threadIdx.x]; }
__device__ __forceinline__ uint32_t* dbg_tris_ptr() { extern __shared__ uint32_t dbg_sm[]; return &dbg_sm[256 + threadIdx.x]; }
#endif

__device__ __forceinline__ bool octree_closest_hit_ww(const TraceParams& P, const V3& o, const V3& d, HitRec* out) {
    const V3 inv = {fdiv(1.0f, d.x), fdiv(1.0f, d.y), fdiv(1.0f, d.z)};
    int stack[kOctStack + 1];
    int sp = 0;
    stack[sp++] = kSentinel;
    int cur = 0;
    float4 A, B;
    for (;;) {
        // ---- descend until a (non-empty) leaf is on top ----
        for (;;) {
            if (cur == kSentinel) return false;
            A = __ldg(&P.oct_nodes[2 * cur]);
            B = __ldg(&P.oct_nodes[2 * cur + 1]);
            const uint32_t meta = __float_as_uint(B.w);
            if (meta & kOctLeafFlag) break;
            const int first_child = (int)__float_as_uint(A.w);
            uint32_t live = meta & 0xffu;
            const float midx = fmul(0.5f, fadd(B.x, A.x)), midy = fmul(0.5f, fadd(B.y, A.y)), midz = fmul(0.5f, fadd(B.z, A.z));
            const float xl = fmul(fsub(A.x, o.x), inv.x), xm = fmul(fsub(midx, o.x), inv.x), xh = fmul(fsub(B.x, o.x), inv.x);
            const float yl = fmul(fsub(A.y, o.y), inv.y), ym = fmul(fsub(midy, o.y), inv.y), yh = fmul(fsub(B.y, o.y), inv.y);
            const float zl = fmul(fsub(A.z, o.z), inv.z), zm = fmul(fsub(midz, o.z), inv.z), zh = fmul(fsub(B.z, o.z), inv.z);
            const float nx[2] = {fminf(xl, xm), fminf(xm, xh)}, fx[2] = {fmaxf(xl, xm), fmaxf(xm, xh)};
            const float ny[2] = {fminf(yl, ym), fminf(ym, yh)}, fy[2] = {fmaxf(yl, ym), fmaxf(ym, yh)};
            const float nz[2] = {fminf(zl, zm), fminf(zm, zh)}, fz[2] = {fmaxf(zl, zm), fmaxf(zm, zh)};
            float tc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float tmin = fmaxf(fmaxf(nx[c & 1], ny[(c >> 1) & 1]), nz[c >> 2]);
                const float tmax = fminf(fminf(fx[c & 1], fy[(c >> 1) & 1]), fz[c >> 2]);
                tc[c] = tmin;
                if (!(tmax >= tmin && tmax > 0.0f)) live &= ~(1u << c);
            }
            while (live) {  // push in descending (t, child) order = pop in the reference's stable ascending order
                int pick = -1;
                float pt = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if ((live >> c) & 1u) {
                        if (pick < 0 || tc[c] >= pt) {
                            pick = c;
                            pt = tc[c];
                        }
                    }
                }
                live &= ~(1u << pick);
                stack[sp++] = first_child + pick;
            }
            cur = stack[--sp];
        }
        // ---- leaf: closest triangle of the list (strict <), then the in-cube acceptance test ----
        const uint32_t count = __float_as_uint(B.w) & ~kOctLeafFlag;
        const float4* tri = P.oct_tris + 3 * (size_t)__float_as_uint(A.w);
        bool have = false;
        HitRec best;
        best.t = 0.f;
        best.u = 0.f;
        best.v = 0.f;
        best.tri = kNoHit;
        for (uint32_t i = 0; i < count; ++i) {
            const float4 t0 = __ldg(tri + 3 * i), t1 = __ldg(tri + 3 * i + 1), t2 = __ldg(tri + 3 * i + 2);
            float t, u, v;
            if (!moller_trumbore(o, d, t0, t1, t2, &t, &u, &v)) continue;
            if (!have || t < best.t) {
                have = true;
                best.t = t;
                best.u = u;
                best.v = v;
                best.tri = __float_as_uint(t2.y);
            }
        }
        if (have) {
            const V3 hp = vadd(o, vscale(d, best.t));
            const bool outside = hp.x < A.x || hp.x > B.x || hp.y < A.y || hp.y > B.y || hp.z < A.z || hp.z > B.z;
            if (!outside) {
                *out = best;
                return true;
            }
        }
        cur = stack[--sp];
    }
}

// Build-time experiments on the traversal stack and the top of the tree (A/B runs documented in profiles/r2_default_kernel_ab.md;
// none is part of the default build):
//   RT_STACK_SMEM=K    the first K stack entries of every thread live in shared memory ([K][256] int2), deeper ones in local memory
//   RT_STACK_REGTOP=1  the top of the stack is cached in two registers: a push spills the previous top to local memory only when
//                      there is one, a pop reads local memory only when the cached entry is gone
//   RT_TOP_SMEM=N      every block stages the first N nodes of the (breadth-first ordered) node array in shared memory and reads
//                      node references < N with LDS instead of LDG
#ifndef RT_STACK_SMEM
#define RT_STACK_SMEM 0
#endif
#ifndef RT_STACK_REGTOP
#define RT_STACK_REGTOP 0
#endif
#ifndef RT_TOP_SMEM
#define RT_TOP_SMEM 0
#endif
#if RT_TOP_SMEM > 0
__device__ __forceinline__ float4* top_nodes_smem() {
    __shared__ float4 top_nodes[4 * RT_TOP_SMEM];
    return top_nodes;
}
// called by all threads of a block before its first traversal
__device__ __forceinline__ void stage_top_nodes(const TraceParams& P) {
    float4* top = top_nodes_smem();
    const uint32_t n = min((uint32_t)RT_TOP_SMEM, P.bvh_top_count);
    for (uint32_t i = threadIdx.x; i < 4u * n; i += blockDim.x) top[i] = __ldg(P.bvh_nodes + i);
    __syncthreads();
}
#endif

__device__ __forceinline__ bool bvh_closest_hit_ww(const TraceParams& P, const V3& o, const V3& d, float t_limit, float early_t, HitRec* out) {
    // The lanes that enter together stay in one loop and meet at its head after every round (one vote per round).
    // Without this the hardware is free to let sub-groups of the warp that left a leaf at different times run the
    // inner-node loop separately for the rest of the traversal: measured 12.5 instead of 18.7 active lanes in that
    // loop and 174 M instead of 124 M warp instructions per frame, depending on where the compiler happened to place
    // its reconvergence points.
    const uint32_t mask = __activemask();
    const float ix = rcp_approx(d.x), iy = rcp_approx(d.y), iz = rcp_approx(d.z);  // box tests only (padded boxes)
    const float ox = -o.x * ix, oy = -o.y * iy, oz = -o.z * iz;
    // stack entries: (node reference, entry distance as float bits): one 8-byte store / load per entry
#if RT_STACK_SMEM > 0
    __shared__ int2 s_stack[RT_STACK_SMEM][256];
    int2 l_stack[kBvhStack > RT_STACK_SMEM ? kBvhStack - RT_STACK_SMEM : 1];
#define RT_STK_ST(i, v)                                        \
    {                                                          \
        if ((i) < RT_STACK_SMEM) s_stack[(i)][threadIdx.x] = (v); \
        else l_stack[(i) - RT_STACK_SMEM] = (v);               \
    }
#define RT_STK_LD(i) ((i) < RT_STACK_SMEM ? s_stack[(i)][threadIdx.x] : l_stack[(i) - RT_STACK_SMEM])
#else
    int2 stack[kBvhStack];
#define RT_STK_ST(i, v) stack[(i)] = (v)
#define RT_STK_LD(i) stack[(i)]
#endif
    RT_STK_ST(0, make_int2(kSentinel, __float_as_int(-FLT_MAX)));
    int sp = 1;
#if RT_STACK_REGTOP
    int2 top = make_int2(0, 0);
    bool top_valid = false;
#define RT_PUSH(v)                     \
    {                                  \
        if (top_valid) {               \
            RT_STK_ST(sp, top);        \
            ++sp;                      \
        }                              \
        top = (v);                     \
        top_valid = true;              \
    }
#define RT_POP(e)                      \
    {                                  \
        if (top_valid) {               \
            e = top;                   \
            top_valid = false;         \
        } else {                       \
            --sp;                      \
            e = RT_STK_LD(sp);         \
        }                              \
    }
#else
#define RT_PUSH(v)          \
    {                       \
        RT_STK_ST(sp, (v)); \
        ++sp;               \
    }
#define RT_POP(e)           \
    {                       \
        --sp;               \
        e = RT_STK_LD(sp);  \
    }
#endif
    HitRec best;
    best.t = t_limit;
    best.u = 0.f;
    best.v = 0.f;
    best.tri = kNoHit;
    int cur = 0;
    bool early = false;  // a shadow ray met a hit with t <= early_t: that hit decides, nothing else is tested
    while (__any_sync(mask, cur != kSentinel)) {
        while ((unsigned)cur < (unsigned)kSentinel) {  // inner nodes
#ifdef RT_DEBUG_STEP_COUNTS
            ++g_dbg_nodes;
#endif
            float4 q0, q1, q2, q3;
#if RT_TOP_SMEM > 0
            if ((uint32_t)cur < min((uint32_t)RT_TOP_SMEM, P.bvh_top_count)) {
                const float4* n = top_nodes_smem() + 4 * cur;
                q0 = n[0], q1 = n[1], q2 = n[2], q3 = n[3];
            } else
#endif
            {
                const float4* n = P.bvh_nodes + 4 * (size_t)cur;
                q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
            }
            const float a0x = fmaf(q0.x, ix, ox), b0x = fmaf(q0.w, ix, ox);
            const float a0y = fmaf(q0.y, iy, oy), b0y = fmaf(q1.x, iy, oy);
            const float a0z = fmaf(q0.z, iz, oz), b0z = fmaf(q1.y, iz, oz);
            const float a1x = fmaf(q1.z, ix, ox), b1x = fmaf(q2.y, ix, ox);
            const float a1y = fmaf(q1.w, iy, oy), b1y = fmaf(q2.z, iy, oy);
            const float a1z = fmaf(q2.x, iz, oz), b1z = fmaf(q2.w, iz, oz);
            const float n0 = fmaxf(fmaxf(fminf(a0x, b0x), fminf(a0y, b0y)), fmaxf(fminf(a0z, b0z), 0.0f));
            const float f0 = fminf(fminf(fmaxf(a0x, b0x), fmaxf(a0y, b0y)), fminf(fmaxf(a0z, b0z), best.t));
            const float n1 = fmaxf(fmaxf(fminf(a1x, b1x), fminf(a1y, b1y)), fmaxf(fminf(a1z, b1z), 0.0f));
            const float f1 = fminf(fminf(fmaxf(a1x, b1x), fmaxf(a1y, b1y)), fminf(fmaxf(a1z, b1z), best.t));
            const bool h0 = n0 <= f0, h1 = n1 <= f1;
            const int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
            const bool go1 = h1 && (!h0 || n1 < n0);  // child 1 first (child 0 wins ties, as before)
            if (h0 && h1) {  // the far child waits on the stack (predicated stores: no local-memory traffic otherwise)
                RT_PUSH(make_int2(go1 ? c0 : c1, __float_as_int(go1 ? n0 : n1)));
            }
            if (h0 || h1) {
                cur = go1 ? c1 : c0;
            } else {
                int2 e;
                do {
                    RT_POP(e);
                } while (__int_as_float(e.y) > best.t);
                cur = e.x;
            }
        }
        if (cur != kSentinel) {  // leaf
            const uint32_t ref = (uint32_t)~cur;
            const uint32_t count = ref & 15u;
            const float4* tri = P.bvh_tris + 3 * (size_t)(ref >> 4);
            for (uint32_t i = 0; i < count; ++i) {
#ifdef RT_DEBUG_STEP_COUNTS
                ++g_dbg_tris;
#endif
                const float4 t0 = __ldg(tri + 3 * i), t1 = __ldg(tri + 3 * i + 1), t2 = __ldg(tri + 3 * i + 2);
                float t, u, v;
                if (!moller_trumbore(o, d, t0, t1, t2, &t, &u, &v)) continue;
                const uint32_t id = __float_as_uint(t2.y);
                if (t < best.t || (t == best.t && id < best.tri)) {
                    best.t = t;
                    best.u = u;
                    best.v = v;
                    best.tri = id;
                    if (t <= early_t) {
                        early = true;
                        break;
                    }
                }
            }
            if (early) {
                cur = kSentinel;
            } else {
                int2 e;
                do {
                    RT_POP(e);
                } while (__int_as_float(e.y) > best.t);
                cur = e.x;
            }
        }
    }
#undef RT_PUSH
#undef RT_POP
#undef RT_STK_ST
#undef RT_STK_LD
    if (best.tri == kNoHit) return false;
    if (!early) {
        const V3 hp = vadd(o, vscale(d, best.t));
        const bool outside = hp.x < P.root_lo[0] || hp.x > P.root_hi[0] || hp.y < P.root_lo[1] || hp.y > P.root_hi[1] || hp.z < P.root_lo[2] ||
                             hp.z > P.root_hi[2];
        if (outside) return false;
    }
    *out = best;
    return true;
}

// ------------------------------------------------------------------------------------------------------
// 4-wide BVH traversal (node layout: bvh4_build.cpp, one node = one 128-byte line). Same hit rules as
// bvh_closest_hit; half as many dependent node visits per ray. The (up to four) children a ray enters are sorted by
// entry distance with a five-exchange network; the nearest is visited next, the others are pushed far to near.
// An empty slot is the box lo = hi = +inf, which no ray can enter (both slab planes lie at the same infinity).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cswap_near(float& da, int& ca, float& db, int& cb) {
    const bool s = db < da;
    const float td = s ? db : da;
    const int tc = s ? cb : ca;
    db = s ? da : db;
    cb = s ? ca : cb;
    da = td;
    ca = tc;
}
__device__ __forceinline__ bool bvh4_closest_hit_ww(const TraceParams& P, const V3& o, const V3& d, float t_limit, float early_t, HitRec* out) {
    const uint32_t mask = __activemask();  // the lanes that enter together meet again after every round (see bvh_closest_hit_ww)
    const float ix = rcp_approx(d.x), iy = rcp_approx(d.y), iz = rcp_approx(d.z);  // box tests only (padded boxes)
    const float ox = -o.x * ix, oy = -o.y * iy, oz = -o.z * iz;
    bool early = false;
    int stack_node[kBvh4Stack];
    float stack_t[kBvh4Stack];
    stack_node[0] = kSentinel;
    stack_t[0] = -FLT_MAX;
    int sp = 1;
    HitRec best;
    best.t = t_limit;
    best.u = 0.f;
    best.v = 0.f;
    best.tri = kNoHit;
    int cur = 0;
    const float kMiss = __int_as_float(0x7f800000);  // +inf
    while (__any_sync(mask, cur != kSentinel)) {
        while ((unsigned)cur < (unsigned)kSentinel) {  // inner nodes
#ifdef RT_DEBUG_STEP_COUNTS
            ++g_dbg_nodes;
#endif
            const float4* n = P.bvh4_nodes + 8 * (size_t)cur;
            const float4 LX = __ldg(n), LY = __ldg(n + 1), LZ = __ldg(n + 2), HX = __ldg(n + 3), HY = __ldg(n + 4), HZ = __ldg(n + 5);
            const float4 CR = __ldg(n + 6);
            float d0, d1, d2, d3;
#define RT_BOX4(K, DK)                                                                                          \
    {                                                                                                           \
        const float ax = fmaf(LX.K, ix, ox), bx = fmaf(HX.K, ix, ox);                                           \
        const float ay = fmaf(LY.K, iy, oy), by = fmaf(HY.K, iy, oy);                                           \
        const float az = fmaf(LZ.K, iz, oz), bz = fmaf(HZ.K, iz, oz);                                           \
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));                \
        const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), best.t));              \
        DK = tn <= tf ? tn : kMiss;                                                                             \
    }
            RT_BOX4(x, d0)
            RT_BOX4(y, d1)
            RT_BOX4(z, d2)
            RT_BOX4(w, d3)
#undef RT_BOX4
            int c0 = __float_as_int(CR.x), c1 = __float_as_int(CR.y), c2 = __float_as_int(CR.z), c3 = __float_as_int(CR.w);
            cswap_near(d0, c0, d1, c1);
            cswap_near(d2, c2, d3, c3);
            cswap_near(d0, c0, d2, c2);
            cswap_near(d1, c1, d3, c3);
            cswap_near(d1, c1, d2, c2);
            // far to near; an entry is kept only if its child was entered (misses sort to the end as +inf)
            if (d1 < kMiss) {  // d1 <= d2 <= d3: nothing to push unless at least two children were entered
                if (d3 < kMiss) {
                    stack_node[sp] = c3;
                    stack_t[sp] = d3;
                    ++sp;
                }
                if (d2 < kMiss) {
                    stack_node[sp] = c2;
                    stack_t[sp] = d2;
                    ++sp;
                }
                stack_node[sp] = c1;
                stack_t[sp] = d1;
                ++sp;
            }
            if (d0 < kMiss) {
                cur = c0;
            } else {
                do {
                    --sp;
                    cur = stack_node[sp];
                } while (stack_t[sp] > best.t);
            }
        }
        if (cur != kSentinel) {  // leaf
            const uint32_t ref = (uint32_t)~cur;
            const uint32_t count = ref & 15u;
            const float4* tri = P.bvh4_tris + 3 * (size_t)(ref >> 4);
            for (uint32_t i = 0; i < count; ++i) {
#ifdef RT_DEBUG_STEP_COUNTS
                ++g_dbg_tris;
#endif
                const float4 t0 = __ldg(tri + 3 * i), t1 = __ldg(tri + 3 * i + 1), t2 = __ldg(tri + 3 * i + 2);
                float t, u, v;
                if (!moller_trumbore(o, d, t0, t1, t2, &t, &u, &v)) continue;
                const uint32_t id = __float_as_uint(t2.y);
                if (t < best.t || (t == best.t && id < best.tri)) {
                    best.t = t;
                    best.u = u;
                    best.v = v;
                    best.tri = id;
                    if (t <= early_t) {
                        early = true;
                        break;
                    }
                }
            }
            if (early) {
                cur = kSentinel;
            } else {
                do {
                    --sp;
                    cur = stack_node[sp];
                } while (stack_t[sp] > best.t);
            }
        }
    }
    if (best.tri == kNoHit) return false;
    if (!early) {
        const V3 hp = vadd(o, vscale(d, best.t));
        const bool outside = hp.x < P.root_lo[0] || hp.x > P.root_hi[0] || hp.y < P.root_lo[1] || hp.y > P.root_hi[1] || hp.z < P.root_lo[2] ||
                             hp.z > P.root_hi[2];
        if (outside) return false;
    }
    *out = best;
    return true;
}

// ------------------------------------------------------------------------------------------------------
// Compressed 8-wide BVH traversal (node layout: cwbvh_build.cpp; scheme after Ylitie, Karras, Laine, HPG 2017).
// Same hit rules as bvh_closest_hit. One node visit = five 16-byte loads and eight box tests, so the chain of
// dependent loads of a ray is ~3x shorter than in the binary tree. Traversal state is a "node group"
// (first child node | hit bits of the not yet visited inner children, ordered by the ray's octant | imask) and a
// "triangle group" (first triangle | hit bits); only node groups are ever pushed.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) {  // every byte: 0xff if its top bit is set, else 0x00
    uint32_t r;
    asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
    return r;
}
// byte j of w as a float without the quarter-rate I2F: place it in the mantissa of 2^23, subtract 2^23 (both exact)
template <int J>
__device__ __forceinline__ float byte_to_float(uint32_t w) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0x4B000000u), "n"(0x7650 + J));
    return __fsub_rn(__uint_as_float(r), 8388608.0f);
}

__device__ __forceinline__ bool cwbvh_closest_hit(const TraceParams& P, const V3& o, const V3& d, float t_limit, float early_t, HitRec* out) {
    const uint32_t mask = __activemask();  // the lanes that enter together meet again after every round (see bvh_closest_hit_ww)
    bool finished = false, early = false;
    // box tests only: keep the reciprocal finite (the triangle test below uses the unmodified direction)
    const float kTiny = 1e-18f;
    const float ix = 1.0f / (fabsf(d.x) > kTiny ? d.x : copysignf(kTiny, d.x));
    const float iy = 1.0f / (fabsf(d.y) > kTiny ? d.y : copysignf(kTiny, d.y));
    const float iz = 1.0f / (fabsf(d.z) > kTiny ? d.z : copysignf(kTiny, d.z));
    // slot s holds the child towards +axis where bit a of s is set; a ray travelling towards +a meets the children
    // with that bit clear first: priority of slot s = s ^ octinv, highest first
    const uint32_t octinv = (d.x < 0.0f ? 0u : 1u) | (d.y < 0.0f ? 0u : 2u) | (d.z < 0.0f ? 0u : 4u);
    const uint32_t octinv4 = octinv * 0x01010101u;
    uint2 stack[kCwStack];
    int sp = 0;
    HitRec best;
    best.t = t_limit;
    best.u = 0.f;
    best.v = 0.f;
    best.tri = kNoHit;
    uint2 G = make_uint2(0u, 0x80000000u);  // the root as a one-child group
    uint2 T = make_uint2(0u, 0u);
    while (__any_sync(mask, !finished)) {
        // ---- node phase: descend until some triangles are waiting (or the traversal is over) ----
        while (T.y == 0u && !finished) {
            if ((G.y & 0xff000000u) == 0u) {
                if (sp == 0) {
                    finished = true;
                    break;
                }
                G = stack[--sp];
            }
            const uint32_t bit = 31u - (uint32_t)__clz((int)G.y);
            G.y &= ~(1u << bit);
            if (G.y & 0xff000000u) stack[sp++] = G;
            const uint32_t slot = (bit - 24u) ^ octinv;
            const uint32_t rel = (uint32_t)__popc(G.y & ~(0xffffffffu << slot));  // inner children in lower slots
            const uint4* n = P.cw_nodes + 5 * (size_t)(G.x + rel);
#ifdef RT_DEBUG_STEP_COUNTS
            ++g_dbg_nodes;
#endif
            const uint4 n0 = __ldg(n), n1 = __ldg(n + 1), n2 = __ldg(n + 2), n3 = __ldg(n + 3), n4 = __ldg(n + 4);
            const float ax = __uint_as_float((n0.w & 0xffu) << 23) * ix;
            const float ay = __uint_as_float(((n0.w >> 8) & 0xffu) << 23) * iy;
            const float az = __uint_as_float(((n0.w >> 16) & 0xffu) << 23) * iz;
            const float bx = (__uint_as_float(n0.x) - o.x) * ix;
            const float by = (__uint_as_float(n0.y) - o.y) * iy;
            const float bz = (__uint_as_float(n0.z) - o.z) * iz;
            uint32_t hitmask = 0u;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t meta4 = half ? n1.w : n1.z;
                const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
                const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
                const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
                const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
                const uint32_t qlx = half ? n2.y : n2.x, qly = half ? n2.w : n2.z, qlz = half ? n3.y : n3.x;
                const uint32_t qhx = half ? n3.w : n3.z, qhy = half ? n4.y : n4.x, qhz = half ? n4.w : n4.z;
                // entry / exit planes by the sign of the direction
                const uint32_t nx = d.x < 0.0f ? qhx : qlx, fx = d.x < 0.0f ? qlx : qhx;
                const uint32_t ny = d.y < 0.0f ? qhy : qly, fy = d.y < 0.0f ? qly : qhy;
                const uint32_t nz = d.z < 0.0f ? qhz : qlz, fz = d.z < 0.0f ? qlz : qhz;
#define RT_CW_CHILD(J)                                                                                                   \
    {                                                                                                                    \
        const float t0x = fmaf(byte_to_float<J>(nx), ax, bx), t1x = fmaf(byte_to_float<J>(fx), ax, bx);                  \
        const float t0y = fmaf(byte_to_float<J>(ny), ay, by), t1y = fmaf(byte_to_float<J>(fy), ay, by);                  \
        const float t0z = fmaf(byte_to_float<J>(nz), az, bz), t1z = fmaf(byte_to_float<J>(fz), az, bz);                  \
        const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));                                                       \
        const float tf = fminf(fminf(t1x, t1y), fminf(t1z, best.t));                                                     \
        if (tn <= tf) hitmask |= ((child_bits4 >> (8 * J)) & 0xffu) << ((bit_index4 >> (8 * J)) & 0xffu);                \
    }
                RT_CW_CHILD(0)
                RT_CW_CHILD(1)
                RT_CW_CHILD(2)
                RT_CW_CHILD(3)
#undef RT_CW_CHILD
            }
            G = make_uint2(n1.x, (hitmask & 0xff000000u) | (n0.w >> 24));
            T = make_uint2(n1.y, hitmask & 0x00ffffffu);
        }
        // ---- triangle phase ----
        while (T.y != 0u) {
            const uint32_t bit = 31u - (uint32_t)__clz((int)T.y);
            T.y &= ~(1u << bit);
#ifdef RT_DEBUG_STEP_COUNTS
            ++g_dbg_tris;
#endif
            const float4* tri = P.cw_tris + 3 * (size_t)(T.x + bit);
            const float4 t0 = __ldg(tri), t1 = __ldg(tri + 1), t2 = __ldg(tri + 2);
            float t, u, v;
            if (!moller_trumbore(o, d, t0, t1, t2, &t, &u, &v)) continue;
            const uint32_t id = __float_as_uint(t2.y);
            if (t < best.t || (t == best.t && id < best.tri)) {
                best.t = t;
                best.u = u;
                best.v = v;
                best.tri = id;
                if (t <= early_t) {
                    early = true;
                    finished = true;
                    T.y = 0u;
                }
            }
        }
    }
    if (best.tri == kNoHit) return false;
    if (!early) {
        const V3 hp = vadd(o, vscale(d, best.t));
        const bool outside = hp.x < P.root_lo[0] || hp.x > P.root_hi[0] || hp.y < P.root_lo[1] || hp.y > P.root_hi[1] || hp.z < P.root_lo[2] ||
                             hp.z > P.root_hi[2];
        if (outside) return false;
    }
    *out = best;
    return true;
}

// ------------------------------------------------------------------------------------------------------
// Camera rays through the perspective grid (ACCEL = 4; pgrid_build.cu, after Hunt & Mark, "Ray-specialized acceleration structures
// for ray tracing", 2008): all camera rays leave one point, so the sample plane of Camera::get_ray is itself an index — the pixel's
// (u, v) names a cell, the cell lists every triangle whose projection (plus a margin of a pixel) can reach it, and the closest hit is
// the minimum over that list of the same Moller-Trumbore test with the same tie rule and the same root-cube acceptance as
// bvh_closest_hit. No tree is walked: a background pixel costs one empty list. Shadow and bounce rays of an ACCEL = 4 kernel go
// through the binary BVH (they do not start at the eye).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pgrid_closest_hit(const TraceParams& P, const V3& o, const V3& d, uint32_t pu, uint32_t pv, HitRec* out) {
    const uint32_t mask = __activemask();  // the lanes of a tile share one or two cells: walk the lists in step
    const uint32_t cell = (pv >> P.pg_shift) * P.pg_nx + (pu >> P.pg_shift);
    uint32_t i = __ldg(P.pg_start + cell);
    const uint32_t end = __ldg(P.pg_start + cell + 1u);
    HitRec best;
    best.t = FLT_MAX;
    best.u = 0.f;
    best.v = 0.f;
    best.tri = kNoHit;
    while (__any_sync(mask, i < end)) {
        if (i < end) {
            const uint2 e = __ldg(P.pg_tris + i);
            // the list is in ascending order of the triangles' smallest Z, and Z of a point on a camera ray is its t (pgrid_build.cu): nothing
            // from here on can come closer than the hit already held (1e-4: rounding of the f32 t against the binary64 Z)
            if (__uint_as_float(e.y) > best.t * 1.0001f) {
                i = end;
            } else {
                const float4* tri = P.bvh_tris + 3 * (size_t)e.x;
                const float4 t0 = __ldg(tri), t1 = __ldg(tri + 1), t2 = __ldg(tri + 2);
                float t, u, v;
                if (moller_trumbore(o, d, t0, t1, t2, &t, &u, &v)) {
                    const uint32_t id = __float_as_uint(t2.y);
                    if (t < best.t || (t == best.t && id < best.tri)) {
                        best.t = t;
                        best.u = u;
                        best.v = v;
                        best.tri = id;
                    }
                }
                ++i;
            }
        }
    }
    if (best.tri == kNoHit) return false;
    const V3 hp = vadd(o, vscale(d, best.t));
    const bool outside = hp.x < P.root_lo[0] || hp.x > P.root_hi[0] || hp.y < P.root_lo[1] || hp.y > P.root_hi[1] || hp.z < P.root_lo[2] ||
                         hp.z > P.root_hi[2];
    if (outside) return false;
    *out = best;
    return true;
}

// ACCEL: 0 octree / 1 binary BVH / 2 compressed 8-wide BVH / 3 4-wide BVH / 4 binary BVH with the camera rays through the perspective grid
// (closest_hit and shadow_blocked treat 4 as 1; trace_pixel_radiance sends the camera ray to pgrid_closest_hit); WW: 0 single-loop, 1 while-while
template <int ACCEL, int WW>
__device__ __forceinline__ bool closest_hit(const TraceParams& P, const V3& o, const V3& d, HitRec* out) {
    if (ACCEL == 0) return WW ? octree_closest_hit_ww(P, o, d, out) : octree_closest_hit(P, o, d, out);
    if (ACCEL == 2) return cwbvh_closest_hit(P, o, d, FLT_MAX, -1.0f, out);
    if (ACCEL == 3) return bvh4_closest_hit_ww(P, o, d, FLT_MAX, -1.0f, out);
    return WW ? bvh_closest_hit_ww(P, o, d, FLT_MAX, -1.0f, out) : bvh_closest_hit(P, o, d, FLT_MAX, -1.0f, out);
}
// Shadow rays through the cube of grids around their light (ACCEL = 4): every point of the segment from the shaded point to the light is
// seen from the light in the same direction, so that direction names a cell whose list holds every triangle the segment can meet. The
// decision is the one of shadow_blocked: c = the closest hit with t <= 1 — a hit with t <= 0.01 ends the search, the point is lit —,
// blocked iff 0.01 < c < 1 and the hit point passes the root-cube rule. The ray runs 0.01 |L| past the light (t up to 1 from an origin
// 0.01 L along, mod.rs:224-225); surfaces there are seen from the light in the opposite direction, so a ray long enough to reach the
// nearest surface beyond the light (|L|^2 >= lg_far2) walks the BVH instead.
template <int WW>
__device__ __forceinline__ bool lgrid_shadow_blocked(const TraceParams& P, const V3& o, const V3& d, uint32_t li, const float4 lp) {
    const uint32_t mask = __activemask();
    const bool far = !(vdot(d, d) < P.lg_far2[li]);
    uint32_t i = 0u, end = 0u;
    if (!far) {
        const float wx = o.x - lp.x, wy = o.y - lp.y, wz = o.z - lp.z;
        const float ax = fabsf(wx), ay = fabsf(wy), az = fabsf(wz);
        uint32_t face;
        float xf, yf, zf;
        if (ax >= ay && ax >= az) {
            face = wx < 0.f ? 1u : 0u, zf = ax, xf = wy, yf = wz;
        } else if (ay >= az) {
            face = wy < 0.f ? 3u : 2u, zf = ay, xf = wz, yf = wx;
        } else {
            face = wz < 0.f ? 5u : 4u, zf = az, xf = wx, yf = wy;
        }
        const float top = 2.0f * P.lg_half - 1.0f;
        const float u = fminf(fmaxf(P.lg_half * (1.0f + xf / zf), 0.0f), top), v = fminf(fmaxf(P.lg_half * (1.0f + yf / zf), 0.0f), top);
        const uint32_t cell = ((li * 6u + face) * P.lg_n + ((uint32_t)v >> P.lg_shift)) * P.lg_n + ((uint32_t)u >> P.lg_shift);
        i = __ldg(P.lg_start + cell);
        end = __ldg(P.lg_start + cell + 1u);
    }
    float best = FLT_MAX;
    // the list is in ascending order of the triangles' distance from the light (a lower bound of it): what lies farther from the light than
    // the ray's origin (0.99 |L|) is behind the shaded surface and cannot be met with t >= 0
    const float reach = 0.99f * sqrtf(vdot(d, d)) * 1.0001f;
    while (__any_sync(mask, i < end)) {
        if (i < end) {
            const uint2 e = __ldg(P.lg_tris + i);
            if (__uint_as_float(e.y) > reach) {
                i = end;
            } else {
                const float4* tri = P.bvh_tris + 3 * (size_t)e.x;
                const float4 t0 = __ldg(tri), t1 = __ldg(tri + 1), t2 = __ldg(tri + 2);
                float t, u, v;
                ++i;
                if (moller_trumbore(o, d, t0, t1, t2, &t, &u, &v) && t <= 1.0f && t < best) {
                    best = t;
                    if (t <= 0.01f) i = end;  // decides "lit" whatever else lies on the segment
                }
            }
        }
    }
    if (far) {
        HitRec h;
        if (!(WW ? bvh_closest_hit_ww(P, o, d, 1.0f, 0.01f, &h) : bvh_closest_hit(P, o, d, 1.0f, 0.01f, &h))) return false;
        return h.t > 0.01f && h.t < 1.0f;
    }
    if (!(best > 0.01f && best < 1.0f)) return false;
    const V3 hp = vadd(o, vscale(d, best));
    const bool outside = hp.x < P.root_lo[0] || hp.x > P.root_hi[0] || hp.y < P.root_lo[1] || hp.y > P.root_hi[1] || hp.z < P.root_lo[2] ||
                         hp.z > P.root_hi[2];
    return !outside;
}

// blocked <=> the closest hit has 0.01 < t < 1.0 (mod.rs:226-230)
template <int ACCEL, int WW>
__device__ __forceinline__ bool shadow_blocked(const TraceParams& P, const V3& o, const V3& d) {
    HitRec h;
    if (ACCEL == 0) {
        if (!(WW ? octree_closest_hit_ww(P, o, d, &h) : octree_closest_hit(P, o, d, &h))) return false;
        return h.t > 0.01f && h.t < 1.0f;
    }
    // BVH: hits with t >= 1 can never block, a hit with t <= 0.01 decides "lit" immediately
    if (ACCEL == 2) {
        if (!cwbvh_closest_hit(P, o, d, 1.0f, 0.01f, &h)) return false;
        return h.t > 0.01f && h.t < 1.0f;
    }
    if (ACCEL == 3) {
        if (!bvh4_closest_hit_ww(P, o, d, 1.0f, 0.01f, &h)) return false;
        return h.t > 0.01f && h.t < 1.0f;
    }
    if (!(WW ? bvh_closest_hit_ww(P, o, d, 1.0f, 0.01f, &h) : bvh_closest_hit(P, o, d, 1.0f, 0.01f, &h))) return false;
    return h.t > 0.01f && h.t < 1.0f;
}

// x^32 by five squarings in binary64, rounded once to binary32: the correctly rounded value of powf(x, 32)
// (glibc's powf, which the reference reaches through f32::powf, is within 1 ulp of it; mod.rs:255)
__device__ __forceinline__ float pow32(float x) {
    double p = (double)x;
    p *= p;
    p *= p;
    p *= p;
    p *= p;
    p *= p;
    return (float)p;
}

__device__ __forceinline__ uint32_t to_u8(float x) {  // color.rs:89-93: (x.min(1).max(0) * 255) as u8
    // Rust's f32::min drops a NaN operand, so NaN -> 1.0 -> 255 (never-sampled pixels are white). The NaN case is
    // spelled out: ptxas may fold fmaxf(fminf(x, 1), 0) into a saturating move, and .sat turns NaN into +0 (seen on the
    // full-frame tonemap kernel: inf / (1 + inf) came out black).
    const float c = (x != x) ? 1.0f : x;
    return __float2uint_rz(fmul(fmaxf(fminf(c, 1.0f), 0.0f), 255.0f));
}
__device__ __forceinline__ uint32_t tonemap_pack(float sr, float sg, float sb, uint32_t n) {
    const float inv = fdiv(1.0f, (float)n);  // film.rs:46: pixel_sum * (1.0 / num_samples as f32)
    const float r = fmul(sr, inv), g = fmul(sg, inv), b = fmul(sb, inv);
    const float mr = fdiv(r, fadd(1.0f, r)), mg = fdiv(g, fadd(1.0f, g)), mb = fdiv(b, fadd(1.0f, b));  // tonemap.rs:4-10
    return to_u8(mb) | (to_u8(mg) << 8) | (to_u8(mr) << 16) | (255u << 24);
}

// ------------------------------------------------------------------------------------------------------
// one pixel sample: camera ray -> closest hit -> shading with shadow rays -> film -> packed LDR pixel
// ------------------------------------------------------------------------------------------------------
struct LaneCounters {
    uint32_t shadow_rays = 0, prim_hit = 0, blocked = 0, bounce_rays = 0;
};

// shade (mod.rs:207-261): direct light with one closest-hit shadow ray per light the surface faces
template <int ACCEL, int WW>
__device__ __forceinline__ void shade_hit(const TraceParams& P, const V3& o, const V3& d, const HitRec& hit, V3* normal_out, float* out_r,
                                          float* out_g, float* out_b, LaneCounters& cnt) {
    float cr = 0.f, cg = 0.f, cb = 0.f;
    const float4 sh = __ldg(&P.tri_shade[hit.tri]);
    const V3 nrm = {sh.x, sh.y, sh.z};
    *normal_out = nrm;
    const uint32_t geom = __float_as_uint(sh.w);
    const V3 hp = vadd(o, vscale(d, hit.t));  // ray.pos + t * ray.dir  (mod.rs:212)
    for (uint32_t li = 0; li < P.num_lights; ++li) {
        const float4 lp = __ldg(&P.lights[2 * li]), lc = __ldg(&P.lights[2 * li + 1]);
        const V3 L = vsub(V3{lp.x, lp.y, lp.z}, hp);
        const V3 Ln = vunit(L);
        const float ndl = vdot(nrm, Ln);
        if (ndl < 0.0f) continue;
        cnt.shadow_rays += 1;
        const V3 so = vadd(hp, vscale(L, 0.01f));
        if ((ACCEL == 4 && P.lg_start && li < (uint32_t)kGridLights) ? lgrid_shadow_blocked<WW>(P, so, L, li, lp) : shadow_blocked<ACCEL, WW>(P, so, L)) {
            cnt.blocked += 1;
            continue;
        }
        const float4 mat = __ldg(&P.materials[geom]);
        float dr = mat.x, dg = mat.y, db = mat.z;
        const int tex = __float_as_int(mat.w);
        if (tex >= 0) {  // Texture::get_texel(hit.u, hit.v), texture.rs:21-27 (index clamped instead of panicking)
            const DevTexture T = P.textures[tex];
            const float fx = fmul(hit.u, (float)T.width), fy = fmul(hit.v, (float)T.height);
            const size_t x = fx > 0.0f ? (size_t)__float2ull_rz(fx) : 0, y = fy > 0.0f ? (size_t)__float2ull_rz(fy) : 0;
            size_t ti = y * T.width + x;
            const size_t last = (size_t)T.width * T.height - 1;
            if (ti > last) ti = last;
            dr = T.rgb[3 * ti];
            dg = T.rgb[3 * ti + 1];
            db = T.rgb[3 * ti + 2];
        }
        const V3 view = vunit(d);
        const V3 refl = vsub(vscale(nrm, fmul(2.0f, ndl)), Ln);  // 2.0 * ndl * normal - normalize(L)
        const float spec = pow32(vdot(view, refl));
        cr = fadd(cr, fmul(fadd(fmul(dr, ndl), spec), lc.x));
        cg = fadd(cg, fmul(fadd(fmul(dg, ndl), spec), lc.y));
        cb = fadd(cb, fmul(fadd(fmul(db, ndl), spec), lc.z));
    }
    *out_r = cr;
    *out_g = cg;
    *out_b = cb;
}

// compute_radiance (mod.rs:132-176) with bounce rays, as an explicit depth-first walk of the recursion tree.
// Level l traces spread * (recursions - l) bounce rays (mod.rs:150); a bounce direction is the first entry of the
// 65 536-entry unit-vector table, starting at a hashed index and stepping (idx + 1) % 65535, that lies in the
// hemisphere of the normal (mod.rs:186-189, sample_generator.rs:25-33). `path` names the ray inside the tree so its
// random numbers do not depend on evaluation order (shared definition with oracle/rt_oracle.cpp).
constexpr int kMaxRecursions = 4;
struct BounceFrame {
    V3 o, d, normal;
    HitRec hit;
    float rr, rg, rb;  // radiance of this hit
    float sr, sg, sb;  // sum over finished sub rays
    uint32_t k, n, path;
};
template <int ACCEL, int WW>
__device__ __noinline__ void radiance_with_bounces(const TraceParams& P, const V3& o0, const V3& d0, const HitRec& hit0, uint32_t pixel,
                                                   uint32_t sample, float* out_r, float* out_g, float* out_b, LaneCounters& cnt) {
    BounceFrame fr[kMaxRecursions + 1];
    int lvl = 0;
    fr[0].o = o0;
    fr[0].d = d0;
    fr[0].hit = hit0;
    fr[0].path = 0u;
    bool entering = true;
    float ret_r = 0.f, ret_g = 0.f, ret_b = 0.f;
    for (;;) {
        BounceFrame& F = fr[lvl];
        if (entering) {
            shade_hit<ACCEL, WW>(P, F.o, F.d, F.hit, &F.normal, &F.rr, &F.rg, &F.rb, cnt);
            const int rec = P.recursions - lvl;
            F.k = 0u;
            F.n = rec < 1 ? 0u : P.sub_spread * (uint32_t)rec;
            F.sr = F.sg = F.sb = 0.f;
            entering = false;
            if (rec < 1) {  // `if recursions < 1 { return radiance; }`
                ret_r = F.rr;
                ret_g = F.rg;
                ret_b = F.rb;
                if (lvl == 0) break;
                --lvl;
                fr[lvl].sr = fadd(fr[lvl].sr, ret_r);
                fr[lvl].sg = fadd(fr[lvl].sg, ret_g);
                fr[lvl].sb = fadd(fr[lvl].sb, ret_b);
                continue;
            }
        }
        if (F.k < F.n) {
            const uint32_t sub_path = F.path * 31u + F.k + 1u;
            F.k += 1u;
            // normalized_vec_pseudo: random_range(0..NUM_SAMPLES - 1)
            uint32_t idx = (uint32_t)(((unsigned long long)hash4(P.seed ^ 0xb0c0ffeeU, pixel, sample, sub_path) * 65535ull) >> 32);
            V3 rd = {P.sample_table[3 * idx], P.sample_table[3 * idx + 1], P.sample_table[3 * idx + 2]};
            while (vdot(rd, F.normal) <= 0.0f) {
                idx = (idx + 1u) % 65535u;  // normalized_vec_lookup: (sample_idx + 1) % SAMPLE_MAX
                rd = V3{P.sample_table[3 * idx], P.sample_table[3 * idx + 1], P.sample_table[3 * idx + 2]};
            }
            V3 hp = vadd(F.o, vscale(F.d, F.hit.t));       // ray.pos + t * ray.dir       (mod.rs:192)
            hp = vadd(hp, vscale(rd, 0.00001f));           // + 0.00001 * random_dir      (mod.rs:193)
            cnt.bounce_rays += 1;
            HitRec h;
            if (closest_hit<ACCEL, WW>(P, hp, rd, &h) && lvl < kMaxRecursions) {
                BounceFrame& N = fr[lvl + 1];
                N.o = hp;
                N.d = rd;
                N.hit = h;
                N.path = sub_path;
                ++lvl;
                entering = true;
            } else {  // None => RGB::black()
                F.sr = fadd(F.sr, 0.0f);
                F.sg = fadd(F.sg, 0.0f);
                F.sb = fadd(F.sb, 0.0f);
            }
            continue;
        }
        // radiance + fold(sum) * (1.0 / num_sub_rays)
        const float inv = fdiv(1.0f, (float)F.n);
        ret_r = fadd(F.rr, fmul(F.sr, inv));
        ret_g = fadd(F.rg, fmul(F.sg, inv));
        ret_b = fadd(F.rb, fmul(F.sb, inv));
        if (lvl == 0) break;
        --lvl;
        fr[lvl].sr = fadd(fr[lvl].sr, ret_r);
        fr[lvl].sg = fadd(fr[lvl].sg, ret_g);
        fr[lvl].sb = fadd(fr[lvl].sb, ret_b);
    }
    *out_r = ret_r;
    *out_g = ret_g;
    *out_b = ret_b;
}

// camera.rs:80-90 with (u, v) = (idx % width, idx / height)  [sic, mod.rs:96]; `nsamp` numbers the sample of the pixel
// n / d for 32-bit unsigned n with magic = floor(2^32 / d) (0xffffffff for d = 1): the estimate is at most one too small
__device__ __forceinline__ uint32_t udiv_magic(uint32_t n, uint32_t d, uint32_t magic) {
    uint32_t q = __umulhi(n, magic);
    if (n - q * d >= d) ++q;
    return q;
}
// compact row c of a launch over consecutive image rows -> image row (first_row + c) mod height; the host keeps first_row < height and
// c < n_rows <= height (a launch that wraps more than once goes through row_list), so one conditional subtraction is the modulo
__device__ __forceinline__ uint32_t wrap_row(const TraceParams& P, uint32_t c) {
    const uint32_t r = P.first_row + c;
    return r >= P.cam.height ? r - P.cam.height : r;
}
__device__ __forceinline__ V3 camera_ray_dir(const TraceParams& P, uint32_t idx, uint32_t col, uint32_t nsamp) {
    const uint32_t W = P.cam.width, H = P.cam.height;
    float xi1 = 0.5f, xi2 = 0.5f;
    if (P.jitter_mode == 1) {
        xi1 = u01(hash4(P.seed, idx, nsamp, 0));
        xi2 = u01(hash4(P.seed, idx, nsamp, 1));
    }
    const uint32_t pu = col /* = idx % W */, pv = udiv_magic(idx, H, P.magic_h);
    const float dir_x = fadd(-P.cam.max_x, fmul(fmul(2.0f, P.cam.max_x), fdiv(fadd((float)pu, xi1), (float)W)));
    const float dir_y = fadd(-P.cam.max_y, fmul(fmul(2.0f, P.cam.max_y), fdiv(fadd((float)pv, xi2), (float)H)));
    const float ndy = -dir_y;
    const float* R = P.cam.rot;
    V3 d;
    d.x = fadd(fadd(fadd(fmul(dir_x, R[0]), fmul(ndy, R[4])), fmul(1.0f, R[8])), fmul(1.0f, R[12]));
    d.y = fadd(fadd(fadd(fmul(dir_x, R[1]), fmul(ndy, R[5])), fmul(1.0f, R[9])), fmul(1.0f, R[13]));
    d.z = fadd(fadd(fadd(fmul(dir_x, R[2]), fmul(ndy, R[6])), fmul(1.0f, R[10])), fmul(1.0f, R[14]));
    return d;
}

// add_sample (film.rs:20-24) + mean, tonemap, pack of the updated pixel (film.rs:43-48, tonemap.rs:4-10, color.rs:89-95)
__device__ __forceinline__ void film_add_sample(const TraceParams& P, uint32_t idx, float4 fs_, float cr, float cg, float cb) {
    fs_.x = fadd(fs_.x, cr);
    fs_.y = fadd(fs_.y, cg);
    fs_.z = fadd(fs_.z, cb);
    if (cr != 0.0f || cg != 0.0f || cb != 0.0f) {  // a black sample adds +0 to sums of squares that are never -0: nothing to write
        float4 sq = P.film_sq[idx];
        sq.x = fadd(sq.x, fmul(cr, cr));
        sq.y = fadd(sq.y, fmul(cg, cg));
        sq.z = fadd(sq.z, fmul(cb, cb));
        P.film_sq[idx] = sq;
    }
    const uint32_t n_new = __float_as_uint(fs_.w) + 1u;
    fs_.w = __uint_as_float(n_new);
    P.film_sum[idx] = fs_;
    // a pixel whose sums are all zero maps to opaque black whatever n is: 0 * (1/n) = 0, 0 / (1 + 0) = 0
    const uint32_t px = (fs_.x == 0.0f && fs_.y == 0.0f && fs_.z == 0.0f) ? 0xff000000u : tonemap_pack(fs_.x, fs_.y, fs_.z, n_new);
    P.ldr[idx] = px;
    if (P.ldr_remote) P.ldr_remote[idx] = px;
}

// ------------------------------------------------------------------------------------------------------
// Bounce wavefront (compute_radiance with RECURSIONS > 0, mod.rs:132-196, level by level instead of depth first).
// The hits of every level are COMPACTED into a dense node list (warp-aggregated atomic append = ballot/popc stream
// compaction), so the incoherent bounce rays of the next level fill whole warps instead of running one after the
// other in the few lanes of a pixel tile that hit something:
//   trace kernel (BOUNCE = 2)   camera ray -> closest hit -> shade (shadow ray) -> level-0 node {hit point, normal, S, pixel}
//   wf_bounce_kernel, level l   one thread per (node, k): bounce direction k of that node (same hash/table walk as the
//                               depth-first code, so the rays are the same), closest hit -> compacted pending record
//   wf_shade_kernel, level l+1  one thread per pending record: shade (shadow ray) -> level-(l+1) node
//   wf_combine_kernel, level l  bottom up: R = S + (sum over k of R_child_k) * (1 / n) with the children in k order
//                               (a missing child is the +0 the reference adds for RGB::black), handed to the parent's
//                               slot k; level 0 adds R to the film
// Same float operations in the same order as radiance_with_bounces => same film.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t wf_append(unsigned int* counter) {
    const uint32_t m = __activemask();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t leader = (uint32_t)__ffs((int)m) - 1u;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned int)__popc(m));
    base = __shfl_sync(m, base, (int)leader);
    return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}
__device__ __forceinline__ void wf_store_node(const TraceParams& P, uint32_t level, uint32_t slot, const V3& hp, uint32_t pixel, const V3& nrm,
                                              uint32_t path, float sr, float sg, float sb, uint32_t parent, uint32_t k, uint32_t sample) {
    const WfLevel& L = P.wf[level];
    if (slot >= L.cap) return;  // cannot happen: capacities are worst case
    float4* rec = L.rec + kWfRecWords * (size_t)slot;
    rec[0] = make_float4(hp.x, hp.y, hp.z, __uint_as_float(pixel));
    rec[1] = make_float4(nrm.x, nrm.y, nrm.z, __uint_as_float(path));
    rec[2] = make_float4(sr, sg, sb, __uint_as_float(parent | (k << 28)));
    // the sample number of the pixel (its film count when the frame began): the bounce rays of this node hash it, and reading it
    // here saves them a dependent fetch of the film record
    rec[3] = make_float4(__uint_as_float(sample), 0.f, 0.f, 0.f);
    for (uint32_t c = 0; c < 3u * L.n_children; ++c) L.child_r[(size_t)slot * 3u * L.n_children + c] = 0.0f;  // RGB::black
}

// What trace_pixel_radiance leaves for finish_pixel: the radiance of one pixel sample and where it goes.
struct PixelOut {
    float cr, cg, cb;
    uint32_t idx, id, plane_slot;
    uint32_t mode;         // 0 nothing left to do, 1 add to the film, 2 store to a sample plane
};

// first half of a pixel sample: camera ray -> closest hit -> shading (shadow / bounce rays)
template <int ACCEL, int WW, int BOUNCE, bool LAP = false>
__device__ __forceinline__ void trace_pixel_radiance(const TraceParams& P, uint32_t col, uint32_t crow, uint32_t lane_sample, LaneCounters& cnt,
                                                     PixelOut& out) {
    const uint32_t W = P.cam.width;
    out.mode = 0u;
    // sample planes (several samples per pixel in one launch): compact row `crow` = plane * plane_rows + row of the pass;
    // sample lanes (persistent kernel): the lanes of a warp item hold `lane_sample` = 0 .. S-1 of the same pixel
    uint32_t plane = lane_sample, prow = crow;
    if (P.planes) {  // planes are plane_rows_padded (a multiple of the tile height) rows apart; the padding rows are idle
        plane = udiv_magic(crow, P.plane_rows_padded, P.magic_plane_rows);
        prow = crow - plane * P.plane_rows_padded;
        if (prow >= P.plane_rows) return;
    }
    const uint32_t row = P.row_list ? P.row_list[prow] : wrap_row(P, prow);
    const uint32_t idx = row * W + col;
    // the sample's number in the film only matters for the hashed sub-pixel offset and the bounce directions; the film
    // record itself is read when the sample is added (finish_pixel), not kept in registers through the traversal
    const uint32_t nsamp = (P.jitter_mode == 1 || BOUNCE != 0) ? __float_as_uint(P.film_sum[idx].w) + plane : 0u;
    // the pixel's film record is read when the sample is added, after the traversal: ask L2 for it now (one instruction, no register;
    // with a cold film — the benchmark flushes L2 between frames — the add otherwise waits for DRAM: 0.1530 -> 0.1494 ms per step)
    if (P.film_prefetch) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.film_sum + idx));
    const V3 d = camera_ray_dir(P, idx, col, nsamp);
    const V3 o = {P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]};

    float cr = 0.f, cg = 0.f, cb = 0.f;
    HitRec hit;
    uint32_t id = kNoHit;
    if (LAP) {  // lap build: the pixel's rays are booked when its sample is committed (no lane totals, see flush_counters)
        cnt.shadow_rays = 0u;
        cnt.bounce_rays = 0u;
    }
    if (ACCEL == 4 ? pgrid_closest_hit(P, o, d, col, udiv_magic(idx, P.cam.height, P.magic_h), &hit) : closest_hit<ACCEL, WW>(P, o, d, &hit)) {
        cnt.prim_hit += 1;
        id = hit.tri;
        if (P.film_prefetch > 1u) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.film_sq + idx));  // a hit: the sums of squares will be touched too
        if (BOUNCE == 2) {  // wavefront: this hit becomes a level-0 node, the film is updated by wf_combine_kernel
            V3 nrm;
            shade_hit<ACCEL, WW>(P, o, d, hit, &nrm, &cr, &cg, &cb, cnt);
            const uint32_t slot = wf_append(&P.wf_counts[0]);
            wf_store_node(P, 0u, slot, vadd(o, vscale(d, hit.t)), idx, nrm, 0u, cr, cg, cb, 0u, 0u, nsamp);
            P.primary_ids[idx] = id;
            return;
        } else if (BOUNCE == 1) {
            radiance_with_bounces<ACCEL, WW>(P, o, d, hit, idx, nsamp, &cr, &cg, &cb, cnt);
        } else {
            V3 nrm;
            shade_hit<ACCEL, WW>(P, o, d, hit, &nrm, &cr, &cg, &cb, cnt);
        }
    }
    if (LAP) P.lap_rays[idx] = (uint16_t)(min(cnt.shadow_rays, 255u) | (min(cnt.bounce_rays, 255u) << 8));
    out.cr = cr;
    out.cg = cg;
    out.cb = cb;
    out.idx = idx;
    out.id = id;
    out.plane_slot = crow * W + col;
    out.mode = P.planes ? 2u : 1u;
}
// second half: the sample goes to its sample plane (film_accumulate_kernel adds the planes in sample order) or into the film
__device__ __forceinline__ void finish_pixel(const TraceParams& P, const PixelOut& out) {
    if (out.mode == 2u) {
        P.planes[out.plane_slot] = make_float4(out.cr, out.cg, out.cb, __uint_as_float(out.id));
    } else if (out.mode == 1u) {
        P.primary_ids[out.idx] = out.id;
        film_add_sample(P, out.idx, P.film_sum[out.idx], out.cr, out.cg, out.cb);
    }
}
// Sample lanes (multi-sample launches of the persistent kernel): lanes [g * S, (g + 1) * S) of a warp item traced samples
// 0 .. S-1 of ONE pixel. The lane of sample 0 collects them with shuffles and adds them to the film in sample order —
// PixelData::add_sample (film.rs:20-24) S times, then mean, tonemap, pack of the final sums: exactly what S consecutive
// single-sample launches leave (their intermediate LDR / id stores are overwritten), with one film read-modify-write
// per pixel and no sample planes. Called by all 32 lanes; `active` is uniform within a sample group.
__device__ __forceinline__ void finish_sample_lanes(const TraceParams& P, const PixelOut& out, uint32_t lane, bool active) {
    const uint32_t S = 1u << P.lane_samples_log2;
    const uint32_t first = lane & ~(S - 1u);
    const bool leader = active && lane == first && out.mode == 1u;
    float4 fs_ = make_float4(0.f, 0.f, 0.f, 0.f), sq = make_float4(0.f, 0.f, 0.f, 0.f);
    bool sq_loaded = false;
    uint32_t id = kNoHit;
    if (leader) fs_ = P.film_sum[out.idx];
    for (uint32_t s = 0; s < S; ++s) {
        const float cr = __shfl_sync(0xffffffffu, out.cr, (int)(first + s));
        const float cg = __shfl_sync(0xffffffffu, out.cg, (int)(first + s));
        const float cb = __shfl_sync(0xffffffffu, out.cb, (int)(first + s));
        id = __shfl_sync(0xffffffffu, out.id, (int)(first + s));
        if (leader) {
            fs_.x = fadd(fs_.x, cr);
            fs_.y = fadd(fs_.y, cg);
            fs_.z = fadd(fs_.z, cb);
            if (cr != 0.0f || cg != 0.0f || cb != 0.0f) {  // see film_add_sample: a black sample leaves the sums of squares alone
                if (!sq_loaded) {
                    sq = P.film_sq[out.idx];
                    sq_loaded = true;
                }
                sq.x = fadd(sq.x, fmul(cr, cr));
                sq.y = fadd(sq.y, fmul(cg, cg));
                sq.z = fadd(sq.z, fmul(cb, cb));
            }
        }
    }
    if (leader) {
        const uint32_t n_new = __float_as_uint(fs_.w) + S;
        fs_.w = __uint_as_float(n_new);
        P.film_sum[out.idx] = fs_;
        if (sq_loaded) P.film_sq[out.idx] = sq;
        P.primary_ids[out.idx] = id;  // the last sample's, as after S launches
        const uint32_t px = (fs_.x == 0.0f && fs_.y == 0.0f && fs_.z == 0.0f) ? 0xff000000u : tonemap_pack(fs_.x, fs_.y, fs_.z, n_new);
        P.ldr[out.idx] = px;
        if (P.ldr_remote) P.ldr_remote[out.idx] = px;
    }
}
template <int ACCEL, int WW, int BOUNCE>
__device__ __forceinline__ void trace_pixel(const TraceParams& P, uint32_t col, uint32_t crow, LaneCounters& cnt) {
    PixelOut out;
    trace_pixel_radiance<ACCEL, WW, BOUNCE>(P, col, crow, 0u, cnt, out);
    finish_pixel(P, out);
}

__device__ __forceinline__ void flush_counters(const TraceParams& P, LaneCounters c, uint32_t lane) {
    // ray counters: warp reduce, one atomic per warp and counter
    const uint32_t s = __reduce_add_sync(0xffffffffu, c.shadow_rays);
    const uint32_t h = __reduce_add_sync(0xffffffffu, c.prim_hit);
    const uint32_t b = __reduce_add_sync(0xffffffffu, c.blocked);
    const uint32_t r = __reduce_add_sync(0xffffffffu, c.bounce_rays);
    if (lane == 0) {
        unsigned long long* set = P.counters + P.counter_set;  // the per-call counters of this call (device_types.h)
        if (r) {
            atomicAdd(&set[CNT_BOUNCE], (unsigned long long)r);
            atomicAdd(&P.counters[CNT_BOUNCE_TOTAL], (unsigned long long)r);
        }
        if (s) {
            atomicAdd(&set[CNT_SHADOW], (unsigned long long)s);
            atomicAdd(&P.counters[CNT_SHADOW_TOTAL], (unsigned long long)s);
        }
        if (h) atomicAdd(&set[CNT_PRIMARY_HITS], (unsigned long long)h);
        if (b) atomicAdd(&set[CNT_BLOCKED], (unsigned long long)b);
    }
}
// Last warp out of a queue-driven launch (persistent and ray-pool kernels): leaves the device state the NEXT launch needs, so
// that a trace call is one kernel launch and no memset — the tile queue back at zero, the other set of per-call ray counters
// zeroed (the next call counts into it), and, on a multi-GPU run, this rank's "my stores of this frame are done" flag
// published at system scope (every warp fenced its own peer stores before it checked out).
__device__ __forceinline__ void warp_checkout(const TraceParams& P, uint32_t lane, uint32_t total_warps) {
    __syncwarp();
    if (lane == 0) {
        if (P.done_flag) __threadfence_system();  // this warp's stores into rank 0's frame are visible before the count below
        else __threadfence();
        const unsigned long long done = atomicAdd(&P.counters[CNT_WARPS_DONE], 1ull) + 1ull;
        if (done == (unsigned long long)total_warps) {
            P.counters[CNT_TILE_QUEUE] = 0ull;
            P.counters[CNT_WARPS_DONE] = 0ull;
            unsigned long long* other = P.counters + (P.counter_set ^ (CNT_SET_B ^ CNT_SET_A));
            other[CNT_SHADOW] = 0ull;
            other[CNT_PRIMARY_HITS] = 0ull;
            other[CNT_BOUNCE] = 0ull;
            other[CNT_BLOCKED] = 0ull;
            if (P.done_flag) {
                __threadfence_system();
                *(volatile uint32_t*)P.done_flag = P.done_value;
                __threadfence_system();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// variant 0: one thread per pixel, block = 8 warps, each warp an 8x4 pixel tile, block = 32x8 pixels
// ------------------------------------------------------------------------------------------------------
template <int ACCEL, int BOUNCE>
__global__ void __launch_bounds__(256) trace_shade_kernel(const __grid_constant__ TraceParams P) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t col = blockIdx.x * 32u + (warp & 3u) * 8u + (lane & 7u);
    const uint32_t crow = blockIdx.y * 8u + (warp >> 2) * 4u + (lane >> 3);
    LaneCounters cnt;
    if (col < P.cam.width && crow < P.n_rows) trace_pixel<ACCEL, 0, BOUNCE>(P, col, crow, cnt);
    flush_counters(P, cnt, lane);
}

// ------------------------------------------------------------------------------------------------------
// variant 1: persistent warps. Every warp pulls 8x4 pixel tiles from a global atomic queue until the frame is
// done, so a slow tile (deep in the statue) never holds 7 finished warps' registers hostage and the tail of the
// launch is made of the cheap all-miss tiles of the lower image rows. Traversal is while-while.
// ------------------------------------------------------------------------------------------------------
#ifndef RT_PERSISTENT_MIN_BLOCKS
#define RT_PERSISTENT_MIN_BLOCKS 3
#endif
// LAP: the launch traces a lap ahead of the band loop into a frame-aligned sample plane (raytracer.cu, lap_build): per-pixel ray counts go
// to P.lap_rays instead of the ray counters, which the commit books (film_accumulate_kernel). A template flag so that the default
// instantiation, which sits at its register limit, does not carry the extra state.
template <int ACCEL, int BOUNCE, bool SAMPLE_LANES = false, bool LAP = false>
__global__ void __launch_bounds__(256, RT_PERSISTENT_MIN_BLOCKS) trace_shade_persistent_kernel(const __grid_constant__ TraceParams P) {
    const uint32_t lane = threadIdx.x & 31u;
#if RT_TOP_SMEM > 0
    if (ACCEL == 1) stage_top_nodes(P);
#endif
    // A warp item is 32 lanes = item_cols x item_rows pixels x S samples (sample innermost, then column, then row):
    // 8 x 4 x 1 for single-sample launches; with SAMPLE_LANES 8 x 2 x 2, 8 x 1 x 4 or 4 x 1 x 8: the lanes of an item
    // share pixels (finish_sample_lanes). Any 8 consecutive lanes hold whole pixels, which the heavy-item split relies on.
    const uint32_t tiles_x = SAMPLE_LANES ? P.items_x : (P.cam.width + 7u) / 8u;
    const uint32_t n_tiles = SAMPLE_LANES ? P.items_x * P.items_y : tiles_x * ((P.n_rows + 3u) / 4u);
    if ((ACCEL == 1 || ACCEL == 4) && P.film_prefetch > 2u) {
        // a cold L2 (first frame, or a frame after other work went through the cache): ask for the whole tree at once instead of
        // discovering it level by level, one DRAM round trip per level of the first rays
        const uint32_t nt = P.bvh_node_lines + P.bvh_tri_lines, ns = nt + P.tri_shade_lines;
        const uint32_t nq = P.tile_order ? n_tiles / 32u : 0u;  // the queue holds at least one entry per tile
        for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ns + nq; t += gridDim.x * blockDim.x) {
            const char* line = t < P.bvh_node_lines ? (const char*)P.bvh_nodes + 128ull * t
                               : t < nt             ? (const char*)P.bvh_tris + 128ull * (t - P.bvh_node_lines)
                               : t < ns             ? (const char*)P.tri_shade + 128ull * (t - nt)
                                                    : (const char*)P.tile_order + 128ull * (t - ns);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
        }
        if (ACCEL == 4) {  // and the grids: cell offsets and lists of the camera grid, then of the light grids
            const uint32_t a = P.grid_lines[0], b = a + P.grid_lines[1], c = b + P.grid_lines[2], e = c + P.grid_lines[3];
            for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < e; t += gridDim.x * blockDim.x) {
                const char* line = t < a   ? (const char*)P.pg_start + 128ull * t
                                   : t < b ? (const char*)P.pg_tris + 128ull * (t - a)
                                   : t < c ? (const char*)P.lg_start + 128ull * (t - b)
                                           : (const char*)P.lg_tris + 128ull * (t - c);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
            }
        }
        // launches that need a pixel's sample number before they can form its rays (hashed sub-pixel offsets, bounce directions) read the
        // film record first thing, with nothing to hide a cold read behind: ask for the records of all rows of the launch now
        if ((P.jitter_mode == 1 || BOUNCE != 0) && !P.planes && P.film_prefetch_rows) {
            const uint32_t per_row = (P.cam.width + 7u) / 8u;  // 128-byte lines of film_sum per image row
            for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < per_row * P.n_rows; t += gridDim.x * blockDim.x) {
                const uint32_t c = t / per_row;
                const uint32_t row = P.row_list ? P.row_list[c] : wrap_row(P, c);
                asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)(P.film_sum + (size_t)row * P.cam.width) + 128ull * (t - c * per_row)));
            }
        }
    }
    LaneCounters cnt;
    // queue length: n_tiles in image order, or the item count the last tile_sort_kernel produced (heavy tiles are
    // split into four 8-pixel items so that their serial divergent chain is spread over four warps)
    const uint32_t n_items = P.tile_order ? *P.queue_items : n_tiles;
    // The next queue slot is claimed when a tile's rays are done, before its film update: the atomic's round trip
    // (~1 us) overlaps the epilogue instead of sitting in front of the next tile, and the claim is early by so little
    // that the heaviest-first order is not disturbed (claiming a whole tile ahead was measured 10 % slower).
    // In the cheap tail of the queue (with the cost-sorted order: the all-miss tiles, about a microsecond of work
    // each) a warp claims kQueueBatch slots at a time, so the tail of the launch is not 3 552 warps queueing for one counter.
    const uint32_t kQueueBatch = P.queue_batch;
    // (only when every warp can expect several items: on a launch with fewer items than a few per resident warp — a 50-row band,
    // the shard of one rank of eight — a batch claim leaves warps without work while one warp walks through its four items)
    const uint32_t batch_from = (P.tile_order && kQueueBatch > 1u && n_items >= 4u * gridDim.x * (blockDim.x >> 5u))
                                    ? (uint32_t)((unsigned long long)n_items * P.queue_batch_from_pct / 100ull)
                                    : 0xffffffffu;
    // The first item of every warp is assigned statically, block-interleaved: the G heaviest items of a sorted queue go to G different
    // blocks (blocks are spread over the SMs round robin), the next G likewise. Claimed dynamically, the heaviest items would go to
    // whichever blocks start first, i.e. pile up on a few SMs — on a launch with about one item per warp (a 50-row band, one rank's
    // shard of a frame split eight ways) those SMs then run 24 heavy warps while the others idle. The counter serves slots from
    // total_warps on.
    const uint32_t total_warps = gridDim.x * (blockDim.x >> 5);
    uint32_t slot = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, slot_end = slot + 1u;  // this warp owns queue slots [slot, slot_end)
    for (;;) {
        uint32_t item = 0;
        bool last_of_batch = true;
        if (lane == 0) {
            // cost-feedback schedule: queue slot -> work item, heaviest tiles (by last frame's cycle count) first
            if (slot < n_items) item = P.tile_order ? P.tile_order[slot] : slot;
            else item = 0xffffffffu;
            last_of_batch = slot + 1u >= slot_end;
        }
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item == 0xffffffffu) break;
        const uint32_t tile = item & kItemTileMask;
        const uint32_t level = (item >> kItemLevelShift) & 3u;
        const uint32_t part = (item >> kItemPartShift) & 15u;
        const uint32_t tile_y = udiv_magic(tile, tiles_x, P.magic_tiles_x);
        uint32_t col, crow, lane_sample = 0u;
        if (SAMPLE_LANES) {
            // worked out per item from a fresh %laneid: more values kept live through the traversal would spill
            uint32_t lane_now;
            asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_now));
            const uint32_t sl = P.lane_samples_log2, cl = P.item_cols_log2;
            const uint32_t lane_px = lane_now >> sl;
            lane_sample = lane_now & ((1u << sl) - 1u);
            col = ((tile - tile_y * tiles_x) << cl) + (lane_px & ((1u << cl) - 1u));
            crow = tile_y * P.item_rows + (lane_px >> cl);
        } else {
            col = (tile - tile_y * tiles_x) * 8u + (lane & 7u);
            crow = tile_y * 4u + (lane >> 3);
        }
        const bool mine = col < P.cam.width && crow < P.n_rows && (level == 0u || (lane >> (4u - level)) == part);
        const long long t0 = clock64();
#ifdef RT_DEBUG_STEP_COUNTS
        g_dbg_nodes = 0;
        g_dbg_tris = 0;
#endif
        PixelOut pout;
        pout.mode = 0u;
        if (mine) trace_pixel_radiance<ACCEL, 1, BOUNCE, LAP>(P, col, crow, lane_sample, cnt, pout);
        __syncwarp();
        if (lane == 0) {
            if (last_of_batch) {
                const uint32_t k = slot >= batch_from ? kQueueBatch : 1u;
                slot = total_warps + (uint32_t)atomicAdd(&P.counters[CNT_TILE_QUEUE], (unsigned long long)k);
                slot_end = slot + k;
            } else {
                ++slot;
            }
        }
        if (SAMPLE_LANES) finish_sample_lanes(P, pout, lane, mine);
        else if (mine) finish_pixel(P, pout);
        __syncwarp();
#ifdef RT_DEBUG_STEP_COUNTS  // developer build only: (inner nodes visited | triangles tested << 16) instead of the primitive id
        if (mine) {
            const uint32_t row = P.row_list ? P.row_list[crow] : wrap_row(P, crow);
            P.primary_ids[row * P.cam.width + col] = min(g_dbg_nodes, 65535u) | (min(g_dbg_tris, 65535u) << 16);
        }
#endif
        const long long dt = clock64() - t0;
        if (lane == 0 && P.tile_cost) {
            const uint32_t c = dt > 0x3fffffffll ? 0x3fffffffu : (uint32_t)dt;
            // a split tile keeps the largest (parts x part cost) seen (sticky, so it stays split while the view lasts)
            if (level) atomicMax(&P.tile_cost[tile], min(c, 0x3fffffffu >> (level + 1u)) << (level + 1u));
            else P.tile_cost[tile] = c;
        }
#ifdef RT_DEBUG_TILE_CLOCKS  // developer build only (tools/): per-item cycle count instead of the primitive id
        if (mine) {
            const uint32_t row = P.row_list ? P.row_list[crow] : wrap_row(P, crow);
            P.primary_ids[row * P.cam.width + col] = (uint32_t)dt;
        }
#endif
    }
    if (!LAP) flush_counters(P, cnt, lane);
    warp_checkout(P, lane, total_warps);
}

// ------------------------------------------------------------------------------------------------------
// variant 2: ray pool (binary BVH, recursions = 0, one light). A warp is a small wavefront machine whose queues live
// in shared memory; lanes are decoupled from pixels:
//   traverse : every lane owns ONE ray (camera or shadow). A lane whose ray ends pushes the result into a ring and,
//              as soon as `pool_refill` lanes are idle, takes the next ray: a waiting shadow ray first, else the
//              camera ray of the next pixel of the warp's current 8x4 tile. The traversal loops therefore run with
//              (almost) all lanes busy instead of waiting for the slowest pixel of a tile; the inner-node loop is
//              left as soon as fewer than `pool_min_inner` lanes are still descending while others wait at a leaf
//   shade    : 32 camera-ray hits at a time: normal, facing test, Phong/texture -> shadow-ray ring (the colour the
//              pixel gets if the light is visible travels with the ray) or, when the surface faces away, -> film ring
//   film     : 32 finished pixels at a time: add_sample, mean, tonemap, pack, LDR store
// Ring pushes/pops are compactions with __ballot_sync/__popc. Every stage evaluates the same f32 expressions as
// trace_pixel/shade_hit, so results are bit-identical to variants 0/1; only the order in which pixels finish differs.
// ------------------------------------------------------------------------------------------------------
constexpr int kPoolShadowCap = 64, kPoolShadeCap = 96, kPoolFinalCap = 64;
constexpr int kPoolShadowWords = 10; // pixel | hit point xyz | 1/L.xyz | colour if lit rgb
constexpr int kPoolShadeWords = 5;   // pixel | t u v | triangle
constexpr int kPoolFinalWords = 4;   // pixel | rgb
constexpr int kPoolWarpWords = kPoolShadowCap * kPoolShadowWords + kPoolShadeCap * kPoolShadeWords + kPoolFinalCap * kPoolFinalWords;
constexpr int kPoolWarps = 8;
constexpr size_t kPoolSmemBytes = (size_t)kPoolWarps * kPoolWarpWords * 4;
constexpr uint32_t kPoolHitCost = 48u;  // schedule cost of a camera-ray hit (its shade + shadow ray), in traversal steps

template <int CAP>
__device__ __forceinline__ uint32_t ring_wrap(uint32_t x) {
    return x >= (uint32_t)CAP ? x - (uint32_t)CAP : x;
}

__global__ void __launch_bounds__(32 * kPoolWarps, 3) trace_shade_pool_kernel(const __grid_constant__ TraceParams P) {
    extern __shared__ uint32_t pool_smem[];
    const uint32_t full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t* const q_shadow = pool_smem + warp * kPoolWarpWords;
    uint32_t* const q_shade = q_shadow + kPoolShadowCap * kPoolShadowWords;
    uint32_t* const q_final = q_shade + kPoolShadeCap * kPoolShadeWords;
    uint32_t shadow_head = 0, shadow_cnt = 0, shade_head = 0, shade_cnt = 0, final_head = 0, final_cnt = 0;

    const uint32_t W = P.cam.width, H = P.cam.height;
    const uint32_t tiles_x = (W + 7u) / 8u;
    const uint32_t n_tiles = tiles_x * ((P.n_rows + 3u) / 4u);
    const uint32_t n_items = P.tile_order ? *P.queue_items : n_tiles;
    const uint32_t refill_at = P.pool_refill, min_inner = P.pool_min_inner;
    bool queue_done = false;
    uint32_t tile_id = 0, tile_px = 32u;  // the warp's current tile and its next unassigned pixel (32 = used up)
    const V3 cam_o = {P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]};
    const float4 lp4 = __ldg(&P.lights[0]), lc4 = __ldg(&P.lights[1]);
    const V3 light_pos = {lp4.x, lp4.y, lp4.z};
    LaneCounters cnt;

#ifdef RT_DEBUG_WARP_EXIT
    if (lane == 0) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        P.primary_ids[(gridDim.x + blockIdx.x) * kPoolWarps + warp] = (uint32_t)ns;
    }
#endif
    // the ray this lane traverses
    int stack_node[kBvhStack];
    float stack_t[kBvhStack];
    stack_node[0] = kSentinel;
    stack_t[0] = -FLT_MAX;
    bool active = false, is_shadow = false;
    uint32_t pix = 0, ray_tile = 0, steps = 0;
    V3 o = cam_o, d = {0.f, 0.f, 0.f};
    float ix = 0.f, iy = 0.f, iz = 0.f, ox = 0.f, oy = 0.f, oz = 0.f;
    float pen_r = 0.f, pen_g = 0.f, pen_b = 0.f, early_t = -1.0f;
    HitRec best;
    best.t = 0.f;
    best.u = 0.f;
    best.v = 0.f;
    best.tri = kNoHit;
    int cur = kSentinel, sp = 1;

    // ---- film stage: n <= 32 finished pixels ----
    auto film_stage = [&](uint32_t n) {
        if (lane < n) {
            const uint32_t s = ring_wrap<kPoolFinalCap>(final_head + lane);
            const uint32_t idx = q_final[s];
            const float r = __uint_as_float(q_final[kPoolFinalCap + s]);
            const float g = __uint_as_float(q_final[2 * kPoolFinalCap + s]);
            const float b = __uint_as_float(q_final[3 * kPoolFinalCap + s]);
            film_add_sample(P, idx, P.film_sum[idx], r, g, b);
        }
        final_head = ring_wrap<kPoolFinalCap>(final_head + n);
        final_cnt -= n;
        __syncwarp();
    };
    auto push_final = [&](bool want, uint32_t idx, float r, float g, float b) {
        const uint32_t mask = __ballot_sync(full, want);
        if (want) {
            const uint32_t s = ring_wrap<kPoolFinalCap>(ring_wrap<kPoolFinalCap>(final_head + final_cnt) + (uint32_t)__popc(mask & lt));
            q_final[s] = idx;
            q_final[kPoolFinalCap + s] = __float_as_uint(r);
            q_final[2 * kPoolFinalCap + s] = __float_as_uint(g);
            q_final[3 * kPoolFinalCap + s] = __float_as_uint(b);
        }
        final_cnt += (uint32_t)__popc(mask);
    };

    // ---- shade stage: n <= 32 camera-ray hits (shade, mod.rs:207-261, for the single light) ----
    auto shade_stage = [&](uint32_t n) {
        const bool have = lane < n;
        uint32_t idx = 0;
        HitRec hit;
        hit.t = hit.u = hit.v = 0.f;
        hit.tri = 0;
        if (have) {
            const uint32_t s = ring_wrap<kPoolShadeCap>(shade_head + lane);
            idx = q_shade[s];
            hit.t = __uint_as_float(q_shade[kPoolShadeCap + s]);
            hit.u = __uint_as_float(q_shade[2 * kPoolShadeCap + s]);
            hit.v = __uint_as_float(q_shade[3 * kPoolShadeCap + s]);
            hit.tri = q_shade[4 * kPoolShadeCap + s];
        }
        shade_head = ring_wrap<kPoolShadeCap>(shade_head + n);
        shade_cnt -= n;
        bool to_shadow = false;
        float cr = 0.f, cg = 0.f, cb = 0.f;
        V3 hp = {0.f, 0.f, 0.f}, inv = {0.f, 0.f, 0.f};
        if (have) {
            const uint32_t nsamp = P.jitter_mode == 1 ? __float_as_uint(P.film_sum[idx].w) : 0u;
            const V3 d0 = camera_ray_dir(P, idx, idx - udiv_magic(idx, W, P.magic_w) * W, nsamp);
            const float4 sh = __ldg(&P.tri_shade[hit.tri]);
            const V3 nrm = {sh.x, sh.y, sh.z};
            const uint32_t geom = __float_as_uint(sh.w);
            hp = vadd(cam_o, vscale(d0, hit.t));
            const V3 L = vsub(light_pos, hp);
            const V3 Ln = vunit(L);
            const float ndl = vdot(nrm, Ln);
            if (!(ndl < 0.0f)) {
                to_shadow = true;
                cnt.shadow_rays += 1;
                const float4 mat = __ldg(&P.materials[geom]);
                float dr = mat.x, dg = mat.y, db = mat.z;
                const int tex = __float_as_int(mat.w);
                if (tex >= 0) {
                    const DevTexture T = P.textures[tex];
                    const float fx = fmul(hit.u, (float)T.width), fy = fmul(hit.v, (float)T.height);
                    const size_t x = fx > 0.0f ? (size_t)__float2ull_rz(fx) : 0, y = fy > 0.0f ? (size_t)__float2ull_rz(fy) : 0;
                    size_t ti = y * T.width + x;
                    const size_t last = (size_t)T.width * T.height - 1;
                    if (ti > last) ti = last;
                    dr = T.rgb[3 * ti];
                    dg = T.rgb[3 * ti + 1];
                    db = T.rgb[3 * ti + 2];
                }
                const V3 view = vunit(d0);
                const V3 refl = vsub(vscale(nrm, fmul(2.0f, ndl)), Ln);
                const float spec = pow32(vdot(view, refl));
                cr = fadd(0.0f, fmul(fadd(fmul(dr, ndl), spec), lc4.x));
                cg = fadd(0.0f, fmul(fadd(fmul(dg, ndl), spec), lc4.y));
                cb = fadd(0.0f, fmul(fadd(fmul(db, ndl), spec), lc4.z));
                inv = V3{rcp_approx(L.x), rcp_approx(L.y), rcp_approx(L.z)};  // box tests only
            }
        }
        const uint32_t mask = __ballot_sync(full, to_shadow);
        if (to_shadow) {
            const uint32_t s = ring_wrap<kPoolShadowCap>(ring_wrap<kPoolShadowCap>(shadow_head + shadow_cnt) + (uint32_t)__popc(mask & lt));
            q_shadow[s] = idx;
            q_shadow[kPoolShadowCap + s] = __float_as_uint(hp.x);
            q_shadow[2 * kPoolShadowCap + s] = __float_as_uint(hp.y);
            q_shadow[3 * kPoolShadowCap + s] = __float_as_uint(hp.z);
            q_shadow[4 * kPoolShadowCap + s] = __float_as_uint(inv.x);
            q_shadow[5 * kPoolShadowCap + s] = __float_as_uint(inv.y);
            q_shadow[6 * kPoolShadowCap + s] = __float_as_uint(inv.z);
            q_shadow[7 * kPoolShadowCap + s] = __float_as_uint(cr);
            q_shadow[8 * kPoolShadowCap + s] = __float_as_uint(cg);
            q_shadow[9 * kPoolShadowCap + s] = __float_as_uint(cb);
        }
        shadow_cnt += (uint32_t)__popc(mask);
        push_final(have && !to_shadow, idx, 0.f, 0.f, 0.f);  // the surface faces away from the light: black
        __syncwarp();
    };

    for (;;) {
        // ---- full batches of the 32-wide stages ----
        while (final_cnt >= 32u) film_stage(32u);
        while (shade_cnt >= 32u && shadow_cnt <= (uint32_t)kPoolShadowCap - 32u) {
            shade_stage(32u);
            while (final_cnt >= 32u) film_stage(32u);
        }
        // once the tile queue is empty a hit must not wait for a full batch: its shadow ray is the end of the longest
        // dependent chain of the launch (camera ray -> shade -> shadow ray)
        if (queue_done && tile_px >= 32u && shade_cnt > 0u && shadow_cnt == 0u) shade_stage(shade_cnt < 32u ? shade_cnt : 32u);
        // ---- refill idle lanes: waiting shadow rays first, then camera rays of the next pixels ----
        const uint32_t idle_mask = __ballot_sync(full, !active);
        const uint32_t n_idle = (uint32_t)__popc(idle_mask);
        if (n_idle >= refill_at) {
            const uint32_t rank = (uint32_t)__popc(idle_mask & lt);
            const uint32_t n_sh = min(n_idle, shadow_cnt);
            bool fresh = false;
            if (!active && rank < n_sh) {
                const uint32_t s = ring_wrap<kPoolShadowCap>(shadow_head + rank);
                pix = q_shadow[s];
                const V3 hp = {__uint_as_float(q_shadow[kPoolShadowCap + s]), __uint_as_float(q_shadow[2 * kPoolShadowCap + s]),
                               __uint_as_float(q_shadow[3 * kPoolShadowCap + s])};
                ix = __uint_as_float(q_shadow[4 * kPoolShadowCap + s]);
                iy = __uint_as_float(q_shadow[5 * kPoolShadowCap + s]);
                iz = __uint_as_float(q_shadow[6 * kPoolShadowCap + s]);
                pen_r = __uint_as_float(q_shadow[7 * kPoolShadowCap + s]);
                pen_g = __uint_as_float(q_shadow[8 * kPoolShadowCap + s]);
                pen_b = __uint_as_float(q_shadow[9 * kPoolShadowCap + s]);
                d = vsub(light_pos, hp);         // L = light.pos - hit_point           (mod.rs:215)
                o = vadd(hp, vscale(d, 0.01f));  // hit_point + 0.01 * L                (mod.rs:224)
                is_shadow = true;
                best.t = 1.0f;    // hits with t >= 1 can never block ...
                early_t = 0.01f;  // ... and a hit with t <= 0.01 decides "lit" at once (mod.rs:226-230)
                fresh = true;
            }
            shadow_head = ring_wrap<kPoolShadowCap>(shadow_head + n_sh);
            shadow_cnt -= n_sh;
            // camera rays: idle lane number (rank - n_sh) takes the next unassigned pixel of the warp's tile(s)
            const uint32_t need = n_idle - n_sh;
            const uint32_t mine = rank - n_sh;
            const bool want_cam = !active && rank >= n_sh;
            uint32_t given = 0, my_tile = 0, my_px = 0;
            bool take = false;
            while (given < need && !queue_done) {
                if (tile_px >= 32u) {
                    uint32_t item = 0;
                    if (lane == 0) {
                        item = (uint32_t)atomicAdd(&P.counters[CNT_TILE_QUEUE], 1ull);
                        if (item < n_items) item = P.tile_order ? P.tile_order[item] : item;
                        else item = 0xffffffffu;
                    }
                    item = __shfl_sync(full, item, 0);
                    if (item == 0xffffffffu) {
                        queue_done = true;
                        break;
                    }
                    // an order written for the persistent kernel may hold split tiles (one item per part): this kernel
                    // always takes whole tiles, so part 0 stands for the tile and the other parts are skipped
                    if (((item >> kItemPartShift) & 15u) != 0u) continue;
                    tile_id = item & kItemTileMask;
                    tile_px = 0u;
                }
                const uint32_t n = min(32u - tile_px, need - given);
                if (want_cam && mine >= given && mine < given + n) {
                    my_tile = tile_id;
                    my_px = tile_px + (mine - given);
                    take = true;
                }
                tile_px += n;
                given += n;
            }
            if (take) {
                const uint32_t ty = udiv_magic(my_tile, tiles_x, P.magic_tiles_x);
                const uint32_t col = (my_tile - ty * tiles_x) * 8u + (my_px & 7u);
                const uint32_t crow = ty * 4u + (my_px >> 3);
                if (col < W && crow < P.n_rows) {
                    uint32_t row;
                    if (P.row_list) {
                        row = P.row_list[crow];
                    } else {
                        row = P.first_row + crow;
                        if (row >= H) row -= H;
                    }
                    pix = row * W + col;
                    const uint32_t nsamp = P.jitter_mode == 1 ? __float_as_uint(P.film_sum[pix].w) : 0u;
                    d = camera_ray_dir(P, pix, col, nsamp);
                    // 1/d only feeds the (padded, conservative) box tests: the approximate reciprocal is enough
                    ix = rcp_approx(d.x);
                    iy = rcp_approx(d.y);
                    iz = rcp_approx(d.z);
                    o = cam_o;
                    is_shadow = false;
                    ray_tile = my_tile;
                    best.t = FLT_MAX;
                    early_t = -1.0f;
                    fresh = true;
                }
            }
            if (fresh) {
                ox = -o.x * ix;
                oy = -o.y * iy;
                oz = -o.z * iz;
                best.u = 0.f;
                best.v = 0.f;
                best.tri = kNoHit;
                cur = 0;
                sp = 1;
                steps = 0;
                active = true;
            }
            __syncwarp();
        }
        if (!__any_sync(full, active)) {
            // no ray left to traverse: flush partial batches; a shade batch may produce new shadow rays
            if (shade_cnt > 0u) {
                shade_stage(min(shade_cnt, 32u));
                continue;
            }
            if (final_cnt > 0u) {
                film_stage(min(final_cnt, 32u));
                continue;
            }
            if (queue_done) break;
            continue;  // every pixel handed out so far lay outside the image (partial edge tiles): fetch more
        }

        // ---- one round of the BVH traversal (same steps as bvh_closest_hit_ww) ----
        for (;;) {
            const bool inner = (unsigned)cur < (unsigned)kSentinel;
            const uint32_t m_in = __ballot_sync(full, inner);
            if (m_in == 0u) break;
            // few lanes still descend and others wait at a leaf / with a finished ray: serve those first
            if ((uint32_t)__popc(m_in) < min_inner && __any_sync(full, active && !inner)) break;
            if (inner) {
                const float4* n = P.bvh_nodes + 4 * (size_t)cur;
                const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
                const float a0x = fmaf(q0.x, ix, ox), b0x = fmaf(q0.w, ix, ox);
                const float a0y = fmaf(q0.y, iy, oy), b0y = fmaf(q1.x, iy, oy);
                const float a0z = fmaf(q0.z, iz, oz), b0z = fmaf(q1.y, iz, oz);
                const float a1x = fmaf(q1.z, ix, ox), b1x = fmaf(q2.y, ix, ox);
                const float a1y = fmaf(q1.w, iy, oy), b1y = fmaf(q2.z, iy, oy);
                const float a1z = fmaf(q2.x, iz, oz), b1z = fmaf(q2.w, iz, oz);
                const float n0 = fmaxf(fmaxf(fminf(a0x, b0x), fminf(a0y, b0y)), fmaxf(fminf(a0z, b0z), 0.0f));
                const float f0 = fminf(fminf(fmaxf(a0x, b0x), fmaxf(a0y, b0y)), fminf(fmaxf(a0z, b0z), best.t));
                const float n1 = fmaxf(fmaxf(fminf(a1x, b1x), fminf(a1y, b1y)), fmaxf(fminf(a1z, b1z), 0.0f));
                const float f1 = fminf(fminf(fmaxf(a1x, b1x), fmaxf(a1y, b1y)), fminf(fmaxf(a1z, b1z), best.t));
                const bool h0 = n0 <= f0, h1 = n1 <= f1;
                const int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
                const bool go1 = h1 && (!h0 || n1 < n0);
                if (h0 && h1) {
                    stack_node[sp] = go1 ? c0 : c1;
                    stack_t[sp] = go1 ? n0 : n1;
                    ++sp;
                }
                ++steps;
                if (h0 || h1) {
                    cur = go1 ? c1 : c0;
                } else {
                    do {
                        --sp;
                        cur = stack_node[sp];
                    } while (stack_t[sp] > best.t);
                }
            }
        }
        if (cur < 0) {  // leaf
            const uint32_t ref = (uint32_t)~cur;
            const uint32_t count = ref & 15u;
            const float4* tri = P.bvh_tris + 3 * (size_t)(ref >> 4);
            bool ended = false;
            steps += count;
            for (uint32_t i = 0; i < count; ++i) {
                const float4 t0 = __ldg(tri + 3 * i), t1 = __ldg(tri + 3 * i + 1), t2 = __ldg(tri + 3 * i + 2);
                float t, u, v;
                if (!moller_trumbore(o, d, t0, t1, t2, &t, &u, &v)) continue;
                const uint32_t id = __float_as_uint(t2.y);
                if (t < best.t || (t == best.t && id < best.tri)) {
                    best.t = t;
                    best.u = u;
                    best.v = v;
                    best.tri = id;
                    if (t <= early_t) {
                        ended = true;
                        break;
                    }
                }
            }
            if (ended) {
                cur = kSentinel;
            } else {
                do {
                    --sp;
                    cur = stack_node[sp];
                } while (stack_t[sp] > best.t);
            }
        }

        // ---- rays that ended in this round ----
        const bool done = active && cur == kSentinel;
        if (__any_sync(full, done)) {
            bool hit = false;
            if (done) {
                hit = best.tri != kNoHit;
                if (hit) {  // root-cube acceptance rule (oct_tree_intersector.rs:164-169 on the scene AABB)
                    const V3 hp = vadd(o, vscale(d, best.t));
                    if (hp.x < P.root_lo[0] || hp.x > P.root_hi[0] || hp.y < P.root_lo[1] || hp.y > P.root_hi[1] || hp.z < P.root_lo[2] ||
                        hp.z > P.root_hi[2])
                        hit = false;
                }
                active = false;
            }
            const bool prim_done = done && !is_shadow, shadow_done = done && is_shadow;
            if (prim_done) {
                P.primary_ids[pix] = hit ? best.tri : kNoHit;
                if (hit) cnt.prim_hit += 1;
                // cost feedback for the next launch of this view: tiles are started in the order of their longest
                // dependent chain (a ray's steps are serial, the rays of a tile run side by side)
                if (P.tile_cost) atomicMax(&P.tile_cost[ray_tile], steps + (hit ? kPoolHitCost : 0u));
            }
            const bool to_shade = prim_done && hit;
            const uint32_t mask = __ballot_sync(full, to_shade);
            if (to_shade) {
                const uint32_t s = ring_wrap<kPoolShadeCap>(ring_wrap<kPoolShadeCap>(shade_head + shade_cnt) + (uint32_t)__popc(mask & lt));
                q_shade[s] = pix;
                q_shade[kPoolShadeCap + s] = __float_as_uint(best.t);
                q_shade[2 * kPoolShadeCap + s] = __float_as_uint(best.u);
                q_shade[3 * kPoolShadeCap + s] = __float_as_uint(best.v);
                q_shade[4 * kPoolShadeCap + s] = best.tri;
            }
            shade_cnt += (uint32_t)__popc(mask);
            // blocked <=> the closest hit has 0.01 < t < 1.0 (mod.rs:226-230)
            const bool blocked = shadow_done && hit && best.t > 0.01f && best.t < 1.0f;
            if (blocked) cnt.blocked += 1;
            const bool lit = shadow_done && !blocked;
            push_final(done && !to_shade, pix, lit ? pen_r : 0.f, lit ? pen_g : 0.f, lit ? pen_b : 0.f);
            __syncwarp();
        }
    }
#ifdef RT_DEBUG_WARP_EXIT  // developer build only (tools/): when every warp ran out of work, instead of the primitive ids
    if (lane == 0) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        P.primary_ids[blockIdx.x * kPoolWarps + warp] = (uint32_t)ns;
    }
#endif
    flush_counters(P, cnt, lane);
    warp_checkout(P, lane, gridDim.x * kPoolWarps);
}

// ------------------------------------------------------------------------------------------------------
// tile_sort_kernel: order[] = tile ids sorted by descending cost (counting sort on a log-scale key, one block).
// Longest-processing-time-first: the tiles that took most cycles in the previous frame are handed out first, the
// cheap all-miss tiles fill the end of the launch, so no warp is still inside a 500k-cycle tile when the queue
// runs dry. Only the schedule depends on it; pixel results do not.
// ------------------------------------------------------------------------------------------------------
constexpr int kSortBuckets = 256;  // 8 per octave of cost
constexpr int kSortWarps = 32;
// A tile whose cost exceeds the split threshold T = (balanced launch time) * split_quarters / 4 becomes 4, 8 or 16 items
// (the smallest count p with cost / p <= T, limited by max_level): the end of a launch is bounded by the slowest single
// item, and a warp that works on 8, 4 or 2 instead of 32 divergent rays has a much shorter serial chain. The extra items cost
// lanes, not time: on a full GPU only the few heaviest tiles qualify (p = 4), and a launch that cannot fill the GPU — the
// shard of one rank when a frame is split over 8 GPUs, a 50-row band — has the issue slots to spare.
//
// One block. Most tiles of a frame cost about the same (the all-miss tiles), so a shared histogram would see tens of
// thousands of atomics on one address: every warp keeps a PRIVATE histogram (32 x 256 counters) and owns a contiguous
// chunk of the tiles, which also keeps tiles of equal cost in image order (neighbouring tiles stay neighbours in the queue).
__global__ void __launch_bounds__(32 * kSortWarps) tile_sort_kernel(const uint32_t* __restrict__ cost, uint32_t* __restrict__ order, uint32_t n,
                                                                     uint32_t n_warps, uint32_t split_quarters, uint32_t max_level,
                                                                     uint32_t min_split_cycles, uint32_t* __restrict__ queue_items) {
    // split_quarters: 0 = never split; q > 0 = split tiles that cost more than (balanced launch time) * q / 4
    __shared__ uint32_t hist[kSortWarps][kSortBuckets];  // counts, then write cursors, per warp and bucket
    __shared__ uint32_t bucket_base[kSortBuckets];
    __shared__ unsigned long long total_cost;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * kSortBuckets; i += blockDim.x) (&hist[0][0])[i] = 0u;
    if (threadIdx.x == 0) total_cost = 0;
    __syncthreads();
    const uint32_t chunk = (n + kSortWarps - 1u) / kSortWarps;
    const uint32_t begin = min(warp * chunk, n), end = min(begin + chunk, n);
    unsigned long long local_sum = 0;
    for (uint32_t i = begin + lane; i < end; i += 32u) local_sum += cost[i];
    local_sum += __shfl_xor_sync(0xffffffffu, local_sum, 16);
    local_sum += __shfl_xor_sync(0xffffffffu, local_sum, 8);
    local_sum += __shfl_xor_sync(0xffffffffu, local_sum, 4);
    local_sum += __shfl_xor_sync(0xffffffffu, local_sum, 2);
    local_sum += __shfl_xor_sync(0xffffffffu, local_sum, 1);
    if (lane == 0) atomicAdd(&total_cost, local_sum);
    __syncthreads();
    // a perfectly balanced launch would take total / n_warps; only tiles that alone exceed a fraction of that can
    // stretch the tail, so only those are split (a launch of uniformly heavy tiles is left alone)
    const unsigned long long balanced = total_cost / (unsigned long long)max(n_warps, 1u);
    const uint32_t split_above =
        split_quarters ? (uint32_t)min((unsigned long long)0x0fffffffu, max(balanced * split_quarters / 4ull, (unsigned long long)min_split_cycles)) : 0xffffffffu;
    auto split_level = [&](uint32_t c) -> uint32_t {  // 0 whole tile, 1 / 2 / 3 = 4 / 8 / 16 parts
        if (c < split_above || max_level == 0u) return 0u;
        const uint32_t lv = c <= 4u * split_above ? 1u : (c <= 8u * split_above ? 2u : 3u);
        return min(lv, max_level);
    };
    auto bucket = [](uint32_t c) {
        // 8 buckets per octave; reversed so that bucket 0 holds the most expensive items
        const int k = (int)(__log2f((float)c + 1.0f) * 8.0f);
        return (uint32_t)(kSortBuckets - 1 - min(max(k, 0), kSortBuckets - 1));
    };
    for (uint32_t i = begin + lane; i < end; i += 32u) {
        const uint32_t c = cost[i], lv = split_level(c);
        if (lv) atomicAdd(&hist[warp][bucket(c >> (lv + 1u))], 2u << lv);
        else atomicAdd(&hist[warp][bucket(c)], 1u);
    }
    __syncthreads();
    // bucket totals -> exclusive scan over buckets -> per-warp write cursors (warp-major inside a bucket = image order)
    if (threadIdx.x < kSortBuckets) {
        uint32_t t = 0;
        for (int w = 0; w < kSortWarps; ++w) t += hist[w][threadIdx.x];
        bucket_base[threadIdx.x] = t;
    }
    __syncthreads();
    if (warp == 0) {  // 256 buckets: 8 per lane
        uint32_t v[8], sum = 0;
        for (int k = 0; k < 8; ++k) {
            v[k] = bucket_base[lane * 8 + k];
            sum += v[k];
        }
        uint32_t incl = sum;
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= (uint32_t)off) incl += o;
        }
        uint32_t run = incl - sum;
        for (int k = 0; k < 8; ++k) {
            bucket_base[lane * 8 + k] = run;
            run += v[k];
        }
        if (lane == 31) *queue_items = incl;
    }
    __syncthreads();
    if (threadIdx.x < kSortBuckets) {
        uint32_t run = bucket_base[threadIdx.x];
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = hist[w][threadIdx.x];
            hist[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
    for (uint32_t i = begin + lane; i < end; i += 32u) {
        const uint32_t c = cost[i], lv = split_level(c);
        if (lv) {
            const uint32_t parts = 2u << lv;
            const uint32_t at = atomicAdd(&hist[warp][bucket(c >> (lv + 1u))], parts);
            for (uint32_t part = 0; part < parts; ++part) order[at + part] = i | (part << kItemPartShift) | (lv << kItemLevelShift);
        } else {
            order[atomicAdd(&hist[warp][bucket(c)], 1u)] = i;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// film clear / full-frame tonemap / owned-row compaction
// ------------------------------------------------------------------------------------------------------
__global__ void film_clear_kernel(float4* sum, float4* sq, uint32_t* ldr, uint32_t* ids, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sum[i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
    sq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ldr[i] = 0xFFFFFFFFu;  // mean of zero samples is NaN -> every channel 255
    ids[i] = kNoHit;
}

// RayTracer::get_tonemapped_pixels over the whole film (mod.rs:120-128); 4 pixels per thread, 16-byte stores
__global__ void tonemap_pack_kernel(const float4* __restrict__ sum, uint32_t* __restrict__ ldr, uint32_t n) {
    const uint32_t i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (i4 >= n) return;
    uint32_t px[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i4 + k < n) {
            const float4 s = sum[i4 + k];
            px[k] = tonemap_pack(s.x, s.y, s.z, __float_as_uint(s.w));
        } else
            px[k] = 0;
    }
    if (i4 + 3 < n && (n % 4u) == 0u) {
        *reinterpret_cast<uint4*>(ldr + i4) = make_uint4(px[0], px[1], px[2], px[3]);
    } else {
        for (int k = 0; k < 4 && i4 + k < n; ++k) ldr[i4 + k] = px[k];
    }
}

// Film::get_estimated_variances (film.rs:50-67): per pixel and channel (sum_sq / (n (n-1)) - sum * sum / (n * n (n-1))) * 50 with
// `n * (n - 1)` evaluated in u32 (wrapping, as the reference's release build does) before the conversion to f32. n <= 1 gives
// x/0 - y/0 = NaN in every channel, exactly as the reference's arithmetic does. One thread per pixel, same f32 operation order.
__global__ void film_variance_kernel(const float4* __restrict__ sum, const float4* __restrict__ sq, float* __restrict__ out, uint32_t n_px) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    const float4 s = sum[i], q = sq[i];
    const uint32_t n = __float_as_uint(s.w);
    const float nn1 = __uint2float_rn(n * (n - 1u));
    const float n2n1 = fmul(__uint2float_rn(n), nn1);
    out[3 * i] = fmul(fsub(fdiv(q.x, nn1), fdiv(fmul(s.x, s.x), n2n1)), 50.0f);
    out[3 * i + 1] = fmul(fsub(fdiv(q.y, nn1), fdiv(fmul(s.y, s.y), n2n1)), 50.0f);
    out[3 * i + 2] = fmul(fsub(fdiv(q.z, nn1), fdiv(fmul(s.z, s.z), n2n1)), 50.0f);
}

// Adds the sample planes of one launch to one pixel of the film in sample order (PixelData::add_sample,
// film.rs:20-24, once per plane), then mean, tonemap, pack of the final sums: exactly what `n_planes` consecutive
// single-sample launches leave. (Fusing this into the trace kernel — the warp that delivers a tile's last sample adds
// them — was measured slower: the per-item __threadfence + atomic costs more than this 15 us pass.)
__device__ __forceinline__ void accumulate_pixel(const TraceParams& P, uint32_t col, uint32_t prow) {
    const uint32_t W = P.cam.width;
    const uint32_t row = P.row_list ? P.row_list[prow] : wrap_row(P, prow);
    const uint32_t idx = row * W + col;
    float4 fs_ = P.film_sum[idx];
    uint32_t n = __float_as_uint(fs_.w), id = kNoHit;
    // sums of squares: a black sample adds +0 to sums that are never -0, so the Σrgb² record is only touched from the first
    // non-black plane on (most pixels of a frame are black: no read, no write)
    float4 sq = make_float4(0.f, 0.f, 0.f, 0.f);
    bool sq_loaded = false;
    for (uint32_t s = 0; s < P.n_planes; ++s) {
        const float4 c = __ldcg(&P.planes[((size_t)s * P.plane_rows_padded + prow) * W + col]);
        fs_.x = fadd(fs_.x, c.x);
        fs_.y = fadd(fs_.y, c.y);
        fs_.z = fadd(fs_.z, c.z);
        if (c.x != 0.0f || c.y != 0.0f || c.z != 0.0f) {
            if (!sq_loaded) {
                sq = P.film_sq[idx];
                sq_loaded = true;
            }
            sq.x = fadd(sq.x, fmul(c.x, c.x));
            sq.y = fadd(sq.y, fmul(c.y, c.y));
            sq.z = fadd(sq.z, fmul(c.z, c.z));
        }
        n += 1u;
        id = __float_as_uint(c.w);
    }
    fs_.w = __uint_as_float(n);
    P.film_sum[idx] = fs_;
    if (sq_loaded) P.film_sq[idx] = sq;
    P.primary_ids[idx] = id;
    const uint32_t px = (fs_.x == 0.0f && fs_.y == 0.0f && fs_.z == 0.0f) ? 0xff000000u : tonemap_pack(fs_.x, fs_.y, fs_.z, n);
    P.ldr[idx] = px;
    if (P.ldr_remote) P.ldr_remote[idx] = px;
}
__global__ void __launch_bounds__(256) film_accumulate_kernel(const __grid_constant__ TraceParams P) {
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t prow = blockIdx.y;
    const bool mine = col < P.cam.width && prow < P.plane_rows;
    if (mine) accumulate_pixel(P, col, prow);
    if (P.lap_rays) {
        // commit of a band of the lap traced ahead (raytracer.cu, lap_commit): the rays of these samples are booked now, so that
        // the per-call counters and the running totals only ever hold rays whose samples are in the film
        uint32_t rays = 0u;
        if (mine) rays = P.lap_rays[wrap_row(P, prow) * P.cam.width + col];
        const uint32_t sh = __reduce_add_sync(0xffffffffu, rays & 255u), bo = __reduce_add_sync(0xffffffffu, rays >> 8);
        if ((threadIdx.x & 31u) == 0u) {
            unsigned long long* set = P.counters + P.counter_set;
            if (sh) {
                atomicAdd(&set[CNT_SHADOW], (unsigned long long)sh);
                atomicAdd(&P.counters[CNT_SHADOW_TOTAL], (unsigned long long)sh);
            }
            if (bo) {
                atomicAdd(&set[CNT_BOUNCE], (unsigned long long)bo);
                atomicAdd(&P.counters[CNT_BOUNCE_TOTAL], (unsigned long long)bo);
            }
        }
    }
}

__global__ void gather_rows_kernel(const uint32_t* __restrict__ ldr, const uint32_t* __restrict__ row_list, uint32_t n_rows, uint32_t width,
                                   uint32_t* __restrict__ out) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = blockIdx.y;
    if (x < width && r < n_rows) out[(size_t)r * width + x] = ldr[(size_t)row_list[r] * width + x];
}

// ------------------------------------------------------------------------------------------------------
// cross-GPU frame fence for the fused peer-store gather (multi_gpu.py): every rank publishes "my stores of frame k
// are done" by writing k+1 into its slot of a flag array that lives on rank 0 (peer-mapped), rank 0 waits for all
// slots, and publishes "frame k has been read" the same way. Waits are bounded (~2 s): a lost peer turns into an
// error count, never into a hung GPU.
// ------------------------------------------------------------------------------------------------------
__global__ void flag_signal_kernel(volatile uint32_t* flag, uint32_t value) {
    __threadfence_system();  // everything this stream did before (kernel boundaries order it) is visible first
    *flag = value;
    __threadfence_system();
}
// the per-frame fence of a rank other than 0 in one launch: "my stores of frame k are done" (*signal = value), then hold the
// stream until *wait_flag >= target ("the frame that used the next buffer has been read")
__global__ void flag_signal_wait_kernel(volatile uint32_t* signal, uint32_t value, volatile uint32_t* wait_flag, uint32_t target,
                                        unsigned long long timeout_ns, uint32_t* timeouts) {
    __threadfence_system();
    *signal = value;
    __threadfence_system();
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int32_t)(*wait_flag - target) < 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {
            atomicAdd(timeouts, 1u);
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}
// optional signal (flags[signal_slot] = target) -> wait until flags[0..n) >= target -> optional release
// (flags[release_slot] = target): rank 0's whole per-frame fence in one launch
__global__ void flag_wait_kernel(volatile uint32_t* flags, uint32_t n, uint32_t target, int signal_slot, int release_slot,
                                 unsigned long long timeout_ns, uint32_t* timeouts) {
    const uint32_t i = threadIdx.x;
    if (signal_slot >= 0 && i == 0) {
        __threadfence_system();
        flags[signal_slot] = target;
    }
    for (uint32_t f = i; f < n; f += blockDim.x) {  // one lane per flag (lanes take several when there are more flags than lanes)
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        // wrap-safe "flags[f] >= target" for counters that only grow
        while ((int32_t)(flags[f] - target) < 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                atomicAdd(timeouts, 1u);
                break;
            }
            __nanosleep(100);
        }
        __threadfence_system();
    }
    __syncthreads();
    if (release_slot >= 0 && i == 0) {
        flags[release_slot] = target;
        __threadfence_system();
    }
}

// ------------------------------------------------------------------------------------------------------
// bounce wavefront kernels (see trace_pixel, BOUNCE = 2)
// ------------------------------------------------------------------------------------------------------
template <int ACCEL>
__global__ void __launch_bounds__(256) wf_bounce_kernel(const __grid_constant__ TraceParams P) {
#if RT_TOP_SMEM > 0
    if (ACCEL == 1) stage_top_nodes(P);
#endif
    const uint32_t l = P.wf_level;
    const WfLevel& L = P.wf[l];
    const uint32_t n_nodes = min(P.wf_counts[l], L.cap), nch = L.n_children;
    const uint32_t total = n_nodes * nch;
    LaneCounters cnt;
    const uint32_t lane = threadIdx.x & 31u;
    // bounce rays differ a lot in cost: warps pull 32 rays at a time from a queue instead of owning a fixed slice
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&P.wf_counts[kWfLevels + l], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) break;
        const uint32_t j = base + lane;
        if (j >= total) continue;
        const uint32_t parent = j / nch, k = j - parent * nch;
        const float4 r0 = L.rec[kWfRecWords * (size_t)parent], r1 = L.rec[kWfRecWords * (size_t)parent + 1];
        const V3 hp0 = {r0.x, r0.y, r0.z}, normal = {r1.x, r1.y, r1.z};
        const uint32_t pixel = __float_as_uint(r0.w), path = __float_as_uint(r1.w);
        const uint32_t sample = __float_as_uint(P.film_sum[pixel].w);  // the film is not touched before the last combine
        const uint32_t sub_path = path * 31u + k + 1u;
        // normalized_vec_pseudo / normalized_vec_lookup, as in radiance_with_bounces
        uint32_t idx = (uint32_t)(((unsigned long long)hash4(P.seed ^ 0xb0c0ffeeU, pixel, sample, sub_path) * 65535ull) >> 32);
        V3 rd = {P.sample_table[3 * idx], P.sample_table[3 * idx + 1], P.sample_table[3 * idx + 2]};
        while (vdot(rd, normal) <= 0.0f) {
            idx = (idx + 1u) % 65535u;
            rd = V3{P.sample_table[3 * idx], P.sample_table[3 * idx + 1], P.sample_table[3 * idx + 2]};
        }
        const V3 o2 = vadd(hp0, vscale(rd, 0.00001f));  // ray.pos + t * ray.dir + 0.00001 * random_dir (mod.rs:192-193)
        cnt.bounce_rays += 1;
        HitRec h;
        if (closest_hit<ACCEL, 1>(P, o2, rd, &h)) {
            // only ~30 % of the bounce rays hit something: the hits are compacted first and shaded (shadow ray) by
            // wf_shade_kernel with full warps, instead of shading here in the few lanes that hit
            const uint32_t slot = wf_append(&P.wf_counts[l + 1u]);
            const WfLevel& N = P.wf[l + 1u];
            if (slot < N.cap) {
                float4* rec = N.rec + kWfRecWords * (size_t)slot;
                rec[0] = make_float4(o2.x, o2.y, o2.z, __uint_as_float(pixel));
                rec[1] = make_float4(rd.x, rd.y, rd.z, __uint_as_float(sub_path));
                rec[2] = make_float4(h.t, h.u, h.v, __uint_as_float(h.tri));
                rec[3] = make_float4(__uint_as_float(parent | (k << 28)), 0.f, 0.f, 0.f);
            }
        }
    }
    flush_counters(P, cnt, lane);
}

// shades the compacted hits of level P.wf_level (shade, mod.rs:207-261, with its shadow ray) and turns the pending records
// {ray, hit} into node records {hit point, normal, shaded radiance}
template <int ACCEL>
__global__ void __launch_bounds__(256) wf_shade_kernel(const __grid_constant__ TraceParams P) {
#if RT_TOP_SMEM > 0
    if (ACCEL == 1) stage_top_nodes(P);
#endif
    const uint32_t l = P.wf_level;
    const WfLevel& L = P.wf[l];
    const uint32_t total = min(P.wf_counts[l], L.cap);
    LaneCounters cnt;
    const uint32_t lane = threadIdx.x & 31u;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&P.wf_counts[2 * kWfLevels + l], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) break;
        const uint32_t i = base + lane;
        if (i >= total) continue;
        const float4* rec = L.rec + kWfRecWords * (size_t)i;
        const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3];
        const V3 o2 = {r0.x, r0.y, r0.z}, rd = {r1.x, r1.y, r1.z};
        HitRec h;
        h.t = r2.x;
        h.u = r2.y;
        h.v = r2.z;
        h.tri = __float_as_uint(r2.w);
        V3 nrm;
        float cr, cg, cb;
        shade_hit<ACCEL, 1>(P, o2, rd, h, &nrm, &cr, &cg, &cb, cnt);
        const uint32_t pk = __float_as_uint(r3.x);
        wf_store_node(P, l, i, vadd(o2, vscale(rd, h.t)), __float_as_uint(r0.w), nrm, __float_as_uint(r1.w), cr, cg, cb, pk & 0x0fffffffu, pk >> 28,
                      __float_as_uint(P.film_sum[__float_as_uint(r0.w)].w));
    }
    flush_counters(P, cnt, lane);
}

// ------------------------------------------------------------------------------------------------------
// wf_stream_kernel: one bounce level of the wavefront as a RAY STREAM (binary BVH). It replaces wf_bounce_kernel +
// wf_shade_kernel, whose warps traced 32 rays in lockstep: bounce rays (random hemisphere directions from surface points)
// and the shadow rays of their hits differ in length by orders of magnitude, so a warp spent most of its instructions
// waiting for its longest ray — 7.8 of 32 lanes active on thai2 (profiles/r2_ncu_summary.txt). Here lanes are
// decoupled from rays:
//   * every lane owns one ray in flight — a bounce ray of the level (job j = (node, k), same hash / table walk as
//     radiance_with_bounces, so the same ray) or the shadow ray of the hit such a ray found;
//   * traversal runs in rounds like bvh_closest_hit_ww (all lanes descend to a leaf, then all test triangles);
//   * when `refill` lanes have no ray in flight (their ray ended, or they are idle) the warp SERVICES them together:
//     a finished bounce ray with a hit starts shade() (mod.rs:207-261) and becomes that hit's first shadow ray IN PLACE, a
//     finished shadow ray adds its light's contribution and moves to the next light or appends the finished node to
//     level l+1 (warp-aggregated append), idle lanes claim new jobs from the level's queue with one atomic.
// (Tried and dropped: handing the level's rays out in direction-sorted order — counting sort of 8192-job chunks by 96 direction
// bins in shared memory. The sort pass cost ~80 us per level and the sorted stream was no faster than node order, which already
// keeps neighbouring surface points together: 1.04 vs 0.88 ms per thai2 frame, profiles/r2_bounce_stream_ab.log.)
//   * with P.wf_chain (every level from the second on sends ONE bounce ray per hit, the reference's RECURSIONS = 2, SUB_SPREAD = 1:
//     2 rays from a camera-ray hit, 1 from their hits) a lane that completes a node continues IN PLACE with that node's bounce ray,
//     so the whole bounce tree of a frame is one launch: no second level launch with its own ramp-up and tail, no job queue for the
//     deeper levels.
// The state that outlives a ray (hit point, bounce direction, barycentrics, partial radiance) lives in shared memory, 18
// words per lane. Every float operation is the one shade_hit / wf_bounce_kernel perform, in the same order, so the film is
// bit-identical to the lockstep wavefront and to the depth-first walk.
// ------------------------------------------------------------------------------------------------------
constexpr int kWsWords = 18;  // per-lane context words (see the enum below)
enum WsField { WS_HPX = 0, WS_HPY, WS_HPZ, WS_RDX, WS_RDY, WS_RDZ, WS_TRI, WS_U, WS_V, WS_LIGHT, WS_CR, WS_CG, WS_CB, WS_PIXEL, WS_PATH, WS_PARENT,
               WS_LEVEL /* level of the node the lane's bounce ray left from; its hit becomes a node of the next level */,
               WS_SAMPLE /* sample number of the pixel */ };
enum WsState { WS_IDLE = 0, WS_BOUNCE = 1, WS_SHADOW = 2 };

// The bounce ray of job j = (node, k) of level P.wf_level: same hash / table walk as radiance_with_bounces (mod.rs:186-189), so the same ray.
// Bounce direction of the sub ray `sub_path` of a hit: the first entry of the unit-vector table, from a hashed start index on, that lies in
// the hemisphere of the normal (normalized_vec_pseudo / normalized_vec_lookup, mod.rs:186-189) — as in radiance_with_bounces, so the same ray.
__device__ __forceinline__ V3 wf_bounce_dir(const TraceParams& P, const V3& normal, uint32_t pixel, uint32_t sample, uint32_t sub_path) {
    uint32_t idx = (uint32_t)(((unsigned long long)hash4(P.seed ^ 0xb0c0ffeeU, pixel, sample, sub_path) * 65535ull) >> 32);
    // The walk takes the first table entry from idx on that lies in the hemisphere; half of the entries do, so it usually ends at the
    // first or second. Three consecutive candidates are fetched per round trip (they share one or two lines): the service phase of
    // the ray stream runs with a few lanes, and a chain of dependent fetches there stalls the lanes that are traversing.
    for (;;) {
        const uint32_t i1 = (idx + 1u) % 65535u, i2 = (i1 + 1u) % 65535u;
        const V3 a = {P.sample_table[3 * idx], P.sample_table[3 * idx + 1], P.sample_table[3 * idx + 2]};
        const V3 b = {P.sample_table[3 * i1], P.sample_table[3 * i1 + 1], P.sample_table[3 * i1 + 2]};
        const V3 c = {P.sample_table[3 * i2], P.sample_table[3 * i2 + 1], P.sample_table[3 * i2 + 2]};
        if (vdot(a, normal) > 0.0f) return a;
        if (vdot(b, normal) > 0.0f) return b;
        if (vdot(c, normal) > 0.0f) return c;
        idx = (i2 + 1u) % 65535u;
    }
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(256, MIN_BLOCKS) wf_stream_kernel(const __grid_constant__ TraceParams P) {
    __shared__ uint32_t ctx[kWsWords][256];
    const uint32_t full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t l = P.wf_level;
    const WfLevel& L = P.wf[l];
    const uint32_t n_nodes = min(P.wf_counts[l], L.cap), nch = L.n_children;
    const uint32_t total = n_nodes * nch;
    const uint32_t refill = max(P.pool_refill, 1u), min_inner = P.pool_min_inner;
    LaneCounters cnt;
    bool queue_empty = false;

    // the ray this lane traverses. Its stack: the first RT_STREAM_STACK_SMEM entries in shared memory, deeper ones in local memory.
    // With unrelated rays in the lanes the L1 is full of node lines that are used once, and the stack lines of 32 warps (a warp's
    // entry i is 256 bytes) do not survive between a push and its pop: in the first version of this kernel the two pop loops,
    // running at 2-4 lanes, collected 16 % of all stall samples waiting for local-memory loads that had gone to L2 — as much
    // as the node fetches themselves (profiles/r2_bounce_stream_blocks.txt).
#ifndef RT_STREAM_STACK_SMEM
#define RT_STREAM_STACK_SMEM 8
#endif
#if RT_STREAM_STACK_SMEM > 0
    __shared__ int2 s_stack[RT_STREAM_STACK_SMEM][256];
    int2 l_stack[kBvhStack > RT_STREAM_STACK_SMEM ? kBvhStack - RT_STREAM_STACK_SMEM : 1];
#define WS_STK_ST(i, v)                                           \
    {                                                             \
        if ((i) < RT_STREAM_STACK_SMEM) s_stack[(i)][tid] = (v);  \
        else l_stack[(i) - RT_STREAM_STACK_SMEM] = (v);           \
    }
#define WS_STK_LD(i) ((i) < RT_STREAM_STACK_SMEM ? s_stack[(i)][tid] : l_stack[(i) - RT_STREAM_STACK_SMEM])
#else
    int2 stack[kBvhStack];
#define WS_STK_ST(i, v) stack[(i)] = (v)
#define WS_STK_LD(i) stack[(i)]
#endif
    WS_STK_ST(0, make_int2(kSentinel, __float_as_int(-FLT_MAX)));
    int state = WS_IDLE, cur = kSentinel, sp = 1;
    V3 o = {0.f, 0.f, 0.f}, d = {0.f, 0.f, 0.f};
    float ix = 0.f, iy = 0.f, iz = 0.f, ox = 0.f, oy = 0.f, oz = 0.f, early_t = -1.0f;
    bool early = false;
    HitRec best;
    best.t = 0.f;
    best.u = 0.f;
    best.v = 0.f;
    best.tri = kNoHit;

    auto start_ray = [&](const V3& ro, const V3& rd, float t_limit, float t_early) {
        o = ro;
        d = rd;
        ix = rcp_approx(rd.x);  // box tests only (padded boxes)
        iy = rcp_approx(rd.y);
        iz = rcp_approx(rd.z);
        ox = -ro.x * ix;
        oy = -ro.y * iy;
        oz = -ro.z * iz;
        best.t = t_limit;
        best.u = 0.f;
        best.v = 0.f;
        best.tri = kNoHit;
        early_t = t_early;
        early = false;
        cur = 0;
        sp = 1;
    };
    // shade()'s loop over the lights from light ctx[WS_LIGHT] on: the first light the surface faces becomes a shadow ray
    // (returns true); no such light left: the node is complete (returns false)
    auto next_shadow_ray = [&]() -> bool {
        const V3 hp = {__uint_as_float(ctx[WS_HPX][tid]), __uint_as_float(ctx[WS_HPY][tid]), __uint_as_float(ctx[WS_HPZ][tid])};
        const float4 sh = __ldg(&P.tri_shade[ctx[WS_TRI][tid]]);
        const V3 nrm = {sh.x, sh.y, sh.z};
        for (uint32_t li = ctx[WS_LIGHT][tid]; li < P.num_lights; ++li) {
            const float4 lp = __ldg(&P.lights[2 * li]);
            const V3 Lv = vsub(V3{lp.x, lp.y, lp.z}, hp);
            const float ndl = vdot(nrm, vunit(Lv));
            if (ndl < 0.0f) continue;
            cnt.shadow_rays += 1;
            ctx[WS_LIGHT][tid] = li;
            // hits with t >= 1 can never block, a hit with t <= 0.01 decides "lit" at once (mod.rs:226-230)
            start_ray(vadd(hp, vscale(Lv, 0.01f)), Lv, 1.0f, 0.01f);
            return true;
        }
        return false;
    };

    for (;;) {
        // ---- service: lanes without a ray in flight that have something to do (a finished ray, or jobs left to claim) ----
        const bool waiting = cur == kSentinel;
        const uint32_t m_trav = __ballot_sync(full, !waiting);
        const uint32_t m_srv = __ballot_sync(full, waiting && (state != WS_IDLE || !queue_empty));
        if (m_trav == 0u && m_srv == 0u) break;  // nothing in flight, nothing finished, the queue is empty
        if (m_trav == 0u || (uint32_t)__popc(m_srv) >= refill) {
            if (waiting && state == WS_BOUNCE) {
                // closest hit of a bounce ray, root-cube acceptance rule as in bvh_closest_hit_ww
                bool hit = best.tri != kNoHit;
                V3 hp = {0.f, 0.f, 0.f};
                if (hit) {
                    hp = vadd(o, vscale(d, best.t));
                    if (hp.x < P.root_lo[0] || hp.x > P.root_hi[0] || hp.y < P.root_lo[1] || hp.y > P.root_hi[1] || hp.z < P.root_lo[2] ||
                        hp.z > P.root_hi[2])
                        hit = false;
                }
                state = WS_IDLE;
                if (hit) {  // shade_hit for this hit: ray.pos + t * ray.dir is the hit point (mod.rs:212)
                    ctx[WS_HPX][tid] = __float_as_uint(hp.x);
                    ctx[WS_HPY][tid] = __float_as_uint(hp.y);
                    ctx[WS_HPZ][tid] = __float_as_uint(hp.z);
                    ctx[WS_RDX][tid] = __float_as_uint(d.x);
                    ctx[WS_RDY][tid] = __float_as_uint(d.y);
                    ctx[WS_RDZ][tid] = __float_as_uint(d.z);
                    ctx[WS_TRI][tid] = best.tri;
                    ctx[WS_U][tid] = __float_as_uint(best.u);
                    ctx[WS_V][tid] = __float_as_uint(best.v);
                    ctx[WS_LIGHT][tid] = 0u;
                    ctx[WS_CR][tid] = 0u;  // +0.0f
                    ctx[WS_CG][tid] = 0u;
                    ctx[WS_CB][tid] = 0u;
                    state = WS_SHADOW + 1;  // shading in progress, no shadow ray chosen yet
                }
            } else if (waiting && state == WS_SHADOW) {
                // blocked <=> the closest hit has 0.01 < t < 1.0 (mod.rs:226-230); the root-cube rule applies to a hit that did
                // not end the search early
                bool hit = best.tri != kNoHit;
                if (hit && !early) {
                    const V3 sp_ = vadd(o, vscale(d, best.t));
                    if (sp_.x < P.root_lo[0] || sp_.x > P.root_hi[0] || sp_.y < P.root_lo[1] || sp_.y > P.root_hi[1] || sp_.z < P.root_lo[2] ||
                        sp_.z > P.root_hi[2])
                        hit = false;
                }
                const bool blocked = hit && best.t > 0.01f && best.t < 1.0f;
                const uint32_t li = ctx[WS_LIGHT][tid];
                if (blocked) {
                    cnt.blocked += 1;
                } else {  // Phong + texture for light li, the operations of shade_hit in its order
                    const V3 hp = {__uint_as_float(ctx[WS_HPX][tid]), __uint_as_float(ctx[WS_HPY][tid]), __uint_as_float(ctx[WS_HPZ][tid])};
                    const V3 rd = {__uint_as_float(ctx[WS_RDX][tid]), __uint_as_float(ctx[WS_RDY][tid]), __uint_as_float(ctx[WS_RDZ][tid])};
                    const float4 sh = __ldg(&P.tri_shade[ctx[WS_TRI][tid]]);
                    const V3 nrm = {sh.x, sh.y, sh.z};
                    const uint32_t geom = __float_as_uint(sh.w);
                    const float4 lp = __ldg(&P.lights[2 * li]), lc = __ldg(&P.lights[2 * li + 1]);
                    const V3 Ln = vunit(vsub(V3{lp.x, lp.y, lp.z}, hp));
                    const float ndl = vdot(nrm, Ln);
                    const float4 mat = __ldg(&P.materials[geom]);
                    float dr = mat.x, dg = mat.y, db = mat.z;
                    const int tex = __float_as_int(mat.w);
                    if (tex >= 0) {  // Texture::get_texel(hit.u, hit.v), texture.rs:21-27 (index clamped instead of panicking)
                        const DevTexture T = P.textures[tex];
                        const float fx = fmul(__uint_as_float(ctx[WS_U][tid]), (float)T.width), fy = fmul(__uint_as_float(ctx[WS_V][tid]), (float)T.height);
                        const size_t x = fx > 0.0f ? (size_t)__float2ull_rz(fx) : 0, y = fy > 0.0f ? (size_t)__float2ull_rz(fy) : 0;
                        size_t ti = y * T.width + x;
                        const size_t last = (size_t)T.width * T.height - 1;
                        if (ti > last) ti = last;
                        dr = T.rgb[3 * ti];
                        dg = T.rgb[3 * ti + 1];
                        db = T.rgb[3 * ti + 2];
                    }
                    const V3 view = vunit(rd);
                    const V3 refl = vsub(vscale(nrm, fmul(2.0f, ndl)), Ln);
                    const float spec = pow32(vdot(view, refl));
                    ctx[WS_CR][tid] = __float_as_uint(fadd(__uint_as_float(ctx[WS_CR][tid]), fmul(fadd(fmul(dr, ndl), spec), lc.x)));
                    ctx[WS_CG][tid] = __float_as_uint(fadd(__uint_as_float(ctx[WS_CG][tid]), fmul(fadd(fmul(dg, ndl), spec), lc.y)));
                    ctx[WS_CB][tid] = __float_as_uint(fadd(__uint_as_float(ctx[WS_CB][tid]), fmul(fadd(fmul(db, ndl), spec), lc.z)));
                }
                ctx[WS_LIGHT][tid] = li + 1u;
                state = WS_SHADOW + 1;
            }
            // lanes in the middle of shade(): the next light the surface faces, or the node is complete
            uint32_t emit = 0u;  // level the lane's finished node belongs to (0 = none: level 0 is written by the trace kernel)
            if (state == WS_SHADOW + 1) {
                if (next_shadow_ray()) state = WS_SHADOW;
                else emit = ctx[WS_LEVEL][tid] + 1u;
            }
            __syncwarp();
            // warp-aggregated append, one per level (the lanes of a warp may be at different levels when rays are chained). Done in
            // warp-uniform control flow with explicit ballots: inside divergent code the set of lanes an __activemask() reports is up to
            // the compiler (an if-converted branch counts lanes that do not append, which leaves holes in the level's node list).
            uint32_t slot = 0u;
            for (uint32_t lv = l + 1u; lv <= (uint32_t)P.recursions; ++lv) {
                const uint32_t m_emit = __ballot_sync(full, emit == lv);
                if (m_emit != 0u) {
                    const uint32_t leader = (uint32_t)__ffs((int)m_emit) - 1u;
                    uint32_t base = 0u;
                    if (lane == leader) base = atomicAdd(&P.wf_counts[lv], (unsigned int)__popc(m_emit));
                    base = __shfl_sync(full, base, (int)leader);
                    if (emit == lv) slot = base + (uint32_t)__popc(m_emit & ((1u << lane) - 1u));
                }
            }
            if (emit != 0u) {
                const float4 sh = __ldg(&P.tri_shade[ctx[WS_TRI][tid]]);
                const V3 nrm = {sh.x, sh.y, sh.z};
                const V3 hp = {__uint_as_float(ctx[WS_HPX][tid]), __uint_as_float(ctx[WS_HPY][tid]), __uint_as_float(ctx[WS_HPZ][tid])};
                const uint32_t pk = ctx[WS_PARENT][tid], pixel = ctx[WS_PIXEL][tid], path = ctx[WS_PATH][tid];
                wf_store_node(P, emit, slot, hp, pixel, nrm, path, __uint_as_float(ctx[WS_CR][tid]), __uint_as_float(ctx[WS_CG][tid]),
                              __uint_as_float(ctx[WS_CB][tid]), pk & 0x0fffffffu, pk >> 28, ctx[WS_SAMPLE][tid]);
                state = WS_IDLE;
                if (P.wf_chain && P.wf[emit].n_children == 1u && slot < P.wf[emit].cap) {
                    // the node's only bounce ray (k = 0) continues in this lane
                    const uint32_t sub_path = path * 31u + 1u;
                    const V3 rd = wf_bounce_dir(P, nrm, pixel, ctx[WS_SAMPLE][tid], sub_path);
                    cnt.bounce_rays += 1;
                    ctx[WS_PATH][tid] = sub_path;
                    ctx[WS_PARENT][tid] = slot;  // | (0 << 28)
                    ctx[WS_LEVEL][tid] = emit;
                    start_ray(vadd(hp, vscale(rd, 0.00001f)), rd, FLT_MAX, -1.0f);
                    state = WS_BOUNCE;
                }
            }
            __syncwarp();
            // idle lanes claim the next jobs of the level: one atomic per warp
            const uint32_t m_idle = __ballot_sync(full, state == WS_IDLE);
            if (m_idle != 0u && !queue_empty) {
                const uint32_t n_idle = (uint32_t)__popc(m_idle);
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&P.wf_counts[kWfLevels + l], n_idle);
                base = __shfl_sync(full, base, 0);
                if (base + n_idle >= total) queue_empty = true;
                const uint32_t jj = base + (uint32_t)__popc(m_idle & ((1u << lane) - 1u));
                if (state == WS_IDLE && jj < total) {  // job jj = (node, k) of level l
                    const uint32_t parent = jj / nch, k = jj - parent * nch;
                    const float4 r0 = L.rec[kWfRecWords * (size_t)parent], r1 = L.rec[kWfRecWords * (size_t)parent + 1];
                    const uint32_t sample = __float_as_uint(L.rec[kWfRecWords * (size_t)parent + 3].x);
                    const uint32_t pixel = __float_as_uint(r0.w), sub_path = __float_as_uint(r1.w) * 31u + k + 1u;
                    const V3 rd = wf_bounce_dir(P, V3{r1.x, r1.y, r1.z}, pixel, sample, sub_path);
                    cnt.bounce_rays += 1;
                    ctx[WS_PIXEL][tid] = pixel;
                    ctx[WS_SAMPLE][tid] = sample;
                    ctx[WS_PATH][tid] = sub_path;
                    ctx[WS_PARENT][tid] = parent | (k << 28);
                    ctx[WS_LEVEL][tid] = l;
                    // ray.pos + t * ray.dir + 0.00001 * random_dir (mod.rs:192-193)
                    start_ray(vadd(V3{r0.x, r0.y, r0.z}, vscale(rd, 0.00001f)), rd, FLT_MAX, -1.0f);
                    state = WS_BOUNCE;
                }
            }
            __syncwarp();
        }

        // ---- one traversal round (the steps of bvh_closest_hit_ww) ----
        // The rays of a warp are unrelated, so the number of inner nodes between two leaves differs widely from lane to lane:
        // the inner-node loop is left as soon as fewer than `min_inner` lanes still descend while others wait at a leaf (those
        // are served first; the descending lanes resume in the next round). One vote per iteration.
        for (;;) {
            const bool inner = (unsigned)cur < (unsigned)kSentinel;
            const uint32_t m_in = __ballot_sync(full, inner);
            if (m_in == 0u) break;
            if ((uint32_t)__popc(m_in) < min_inner && __any_sync(full, cur < 0)) break;
            if (inner) {
                const float4* n = P.bvh_nodes + 4 * (size_t)cur;
                const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
                const float a0x = fmaf(q0.x, ix, ox), b0x = fmaf(q0.w, ix, ox);
                const float a0y = fmaf(q0.y, iy, oy), b0y = fmaf(q1.x, iy, oy);
                const float a0z = fmaf(q0.z, iz, oz), b0z = fmaf(q1.y, iz, oz);
                const float a1x = fmaf(q1.z, ix, ox), b1x = fmaf(q2.y, ix, ox);
                const float a1y = fmaf(q1.w, iy, oy), b1y = fmaf(q2.z, iy, oy);
                const float a1z = fmaf(q2.x, iz, oz), b1z = fmaf(q2.w, iz, oz);
                const float n0 = fmaxf(fmaxf(fminf(a0x, b0x), fminf(a0y, b0y)), fmaxf(fminf(a0z, b0z), 0.0f));
                const float f0 = fminf(fminf(fmaxf(a0x, b0x), fmaxf(a0y, b0y)), fminf(fmaxf(a0z, b0z), best.t));
                const float n1 = fmaxf(fmaxf(fminf(a1x, b1x), fminf(a1y, b1y)), fmaxf(fminf(a1z, b1z), 0.0f));
                const float f1 = fminf(fminf(fmaxf(a1x, b1x), fmaxf(a1y, b1y)), fminf(fmaxf(a1z, b1z), best.t));
                const bool h0 = n0 <= f0, h1 = n1 <= f1;
                const int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
                const bool go1 = h1 && (!h0 || n1 < n0);
                if (h0 && h1) {
                    WS_STK_ST(sp, make_int2(go1 ? c0 : c1, __float_as_int(go1 ? n0 : n1)));
                    ++sp;
                }
                if (h0 || h1) {
                    cur = go1 ? c1 : c0;
                } else {
                    int2 e;
                    do {
                        --sp;
                        e = WS_STK_LD(sp);
                    } while (__int_as_float(e.y) > best.t);
                    cur = e.x;
                }
            }
        }
        if (cur < 0) {  // leaf
            const uint32_t ref = (uint32_t)~cur;
            const uint32_t count = ref & 15u;
            const float4* tri = P.bvh_tris + 3 * (size_t)(ref >> 4);
            for (uint32_t i = 0; i < count; ++i) {
                const float4 t0 = __ldg(tri + 3 * i), t1 = __ldg(tri + 3 * i + 1), t2 = __ldg(tri + 3 * i + 2);
                float t, u, v;
                if (!moller_trumbore(o, d, t0, t1, t2, &t, &u, &v)) continue;
                const uint32_t id = __float_as_uint(t2.y);
                if (t < best.t || (t == best.t && id < best.tri)) {
                    best.t = t;
                    best.u = u;
                    best.v = v;
                    best.tri = id;
                    if (t <= early_t) {
                        early = true;
                        break;
                    }
                }
            }
            if (early) {
                cur = kSentinel;
            } else {
                int2 e;
                do {
                    --sp;
                    e = WS_STK_LD(sp);
                } while (__int_as_float(e.y) > best.t);
                cur = e.x;
            }
        }
        __syncwarp();
    }
#undef WS_STK_ST
#undef WS_STK_LD
    flush_counters(P, cnt, lane);
}

__global__ void __launch_bounds__(256) wf_combine_kernel(const __grid_constant__ TraceParams P) {
    const uint32_t l = P.wf_level;
    const WfLevel& L = P.wf[l];
    const uint32_t n_nodes = min(P.wf_counts[l], L.cap), nch = L.n_children;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
        const float4 r2 = L.rec[kWfRecWords * (size_t)i + 2];
        float rr = r2.x, rg = r2.y, rb = r2.z;
        if (nch > 0u) {  // radiance + fold(sum) * (1.0 / num_sub_rays)   (mod.rs:154-175)
            float sr = 0.f, sg = 0.f, sb = 0.f;
            const float* c = L.child_r + (size_t)i * 3u * nch;
            for (uint32_t k = 0; k < nch; ++k) {
                sr = fadd(sr, c[3 * k]);
                sg = fadd(sg, c[3 * k + 1]);
                sb = fadd(sb, c[3 * k + 2]);
            }
            const float inv = fdiv(1.0f, (float)nch);
            rr = fadd(rr, fmul(sr, inv));
            rg = fadd(rg, fmul(sg, inv));
            rb = fadd(rb, fmul(sb, inv));
        }
        if (l == 0u) {
            const uint32_t pixel = __float_as_uint(L.rec[kWfRecWords * (size_t)i].w);
            film_add_sample(P, pixel, P.film_sum[pixel], rr, rg, rb);
        } else {
            const uint32_t pk = __float_as_uint(r2.w), parent = pk & 0x0fffffffu, k = pk >> 28;
            const WfLevel& U = P.wf[l - 1u];
            float* dst = U.child_r + ((size_t)parent * U.n_children + k) * 3u;
            dst[0] = rr;
            dst[1] = rg;
            dst[2] = rb;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------------
template <int ACCEL, int BOUNCE>
static void launch_trace_t(const TraceParams& p, int variant, uint32_t blocks, cudaStream_t stream) {
    if (variant == 2 && ACCEL == 1 && BOUNCE == 0 && p.num_lights == 1 && !p.planes) {
        // > 48 KB of dynamic shared memory needs the opt-in (set per device by pool_blocks_per_sm, which every handle calls first)
        trace_shade_pool_kernel<<<blocks, 32 * kPoolWarps, kPoolSmemBytes, stream>>>(p);
        return;
    }
    if (variant == 0) {
        dim3 grid((p.cam.width + 31u) / 32u, (p.n_rows + 7u) / 8u);
        trace_shade_kernel<ACCEL, BOUNCE><<<grid, 256, 0, stream>>>(p);
    } else {
#ifdef RT_DEBUG_STEP_COUNTS
        trace_shade_persistent_kernel<ACCEL, BOUNCE><<<blocks, 256, 2048, stream>>>(p);
#else
        if (BOUNCE == 0 && p.lane_samples_log2 != 0u) trace_shade_persistent_kernel<ACCEL, 0, true><<<blocks, 256, 0, stream>>>(p);
        else if (p.lap_rays) trace_shade_persistent_kernel<ACCEL, BOUNCE, false, true><<<blocks, 256, 0, stream>>>(p);
        else trace_shade_persistent_kernel<ACCEL, BOUNCE><<<blocks, 256, 0, stream>>>(p);
#endif
    }
}
// bounce_mode: 0 none, 1 depth first in the thread, 2 wavefront (the trace kernel only emits level-0 nodes)
cudaError_t launch_trace(const TraceParams& p, int accel, int variant, int persistent_blocks, cudaStream_t stream) {
    if (p.n_rows == 0) return cudaSuccess;
    const uint32_t tiles = p.items_x * p.items_y;  // = 8x4 pixel tiles unless the launch uses sample lanes
    uint32_t blocks = (uint32_t)persistent_blocks;
    if (blocks * 8u > tiles) blocks = (tiles + 7u) / 8u;
    const int bounce = p.recursions > 0 ? (p.wf_counts ? 2 : 1) : 0;
    if (bounce == 2 && variant == 0) return cudaErrorInvalidValue;  // the wavefront trace kernel is the persistent one
    switch (accel * 3 + bounce) {
        case 0: launch_trace_t<0, 0>(p, variant, blocks, stream); break;
        case 1: launch_trace_t<0, 1>(p, variant, blocks, stream); break;
        case 2: trace_shade_persistent_kernel<0, 2><<<blocks, 256, 0, stream>>>(p); break;
        case 3: launch_trace_t<1, 0>(p, variant, blocks, stream); break;
        case 4: launch_trace_t<1, 1>(p, variant, blocks, stream); break;
        case 5: trace_shade_persistent_kernel<1, 2><<<blocks, 256, 0, stream>>>(p); break;
        case 6: launch_trace_t<2, 0>(p, variant, blocks, stream); break;
        case 7: launch_trace_t<2, 1>(p, variant, blocks, stream); break;
        case 8: trace_shade_persistent_kernel<2, 2><<<blocks, 256, 0, stream>>>(p); break;
        case 9: launch_trace_t<3, 0>(p, variant, blocks, stream); break;
        case 10: launch_trace_t<3, 1>(p, variant, blocks, stream); break;
        case 11: trace_shade_persistent_kernel<3, 2><<<blocks, 256, 0, stream>>>(p); break;
        case 12: launch_trace_t<4, 0>(p, 1, blocks, stream); break;  // persistent kernel only (raytracer.cu selects 4 for variant 1)
        case 13: launch_trace_t<4, 1>(p, 1, blocks, stream); break;
        case 14: trace_shade_persistent_kernel<4, 2><<<blocks, 256, 0, stream>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_wf_bounce(const TraceParams& p, int accel, int blocks, cudaStream_t stream) {
    switch (accel) {
        case 0: wf_bounce_kernel<0><<<blocks, 256, 0, stream>>>(p); break;
        case 1: wf_bounce_kernel<1><<<blocks, 256, 0, stream>>>(p); break;
        case 2: wf_bounce_kernel<2><<<blocks, 256, 0, stream>>>(p); break;
        case 3: wf_bounce_kernel<3><<<blocks, 256, 0, stream>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_wf_shade(const TraceParams& p, int accel, int blocks, cudaStream_t stream) {
    switch (accel) {
        case 0: wf_shade_kernel<0><<<blocks, 256, 0, stream>>>(p); break;
        case 1: wf_shade_kernel<1><<<blocks, 256, 0, stream>>>(p); break;
        case 2: wf_shade_kernel<2><<<blocks, 256, 0, stream>>>(p); break;
        case 3: wf_shade_kernel<3><<<blocks, 256, 0, stream>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_wf_stream(const TraceParams& p, int blocks_per_sm, int num_sms, cudaStream_t stream) {
    switch (blocks_per_sm) {
        case 3: wf_stream_kernel<3><<<3 * num_sms, 256, 0, stream>>>(p); break;
        case 4: wf_stream_kernel<4><<<4 * num_sms, 256, 0, stream>>>(p); break;
        case 5: wf_stream_kernel<5><<<5 * num_sms, 256, 0, stream>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
// resident 256-thread blocks per SM of the lockstep wavefront kernels (kind 0 = wf_bounce_kernel, 1 = wf_shade_kernel)
int wf_blocks_per_sm(int kind, int accel) {
    int n = 0;
#define RT_OCC(K) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, K, 256, 0)
    switch (kind * 4 + accel) {
        case 0: RT_OCC(wf_bounce_kernel<0>); break;
        case 1: RT_OCC(wf_bounce_kernel<1>); break;
        case 2: RT_OCC(wf_bounce_kernel<2>); break;
        case 3: RT_OCC(wf_bounce_kernel<3>); break;
        case 4: RT_OCC(wf_shade_kernel<0>); break;
        case 5: RT_OCC(wf_shade_kernel<1>); break;
        case 6: RT_OCC(wf_shade_kernel<2>); break;
        case 7: RT_OCC(wf_shade_kernel<3>); break;
        default: break;
    }
#undef RT_OCC
    return n > 0 ? n : 1;
}
cudaError_t launch_wf_combine(const TraceParams& p, int blocks, cudaStream_t stream) {
    wf_combine_kernel<<<blocks, 256, 0, stream>>>(p);
    return cudaGetLastError();
}
int pool_blocks_per_sm() {
    int n = 0;
    cudaFuncSetAttribute(trace_shade_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoolSmemBytes);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_shade_pool_kernel, 32 * kPoolWarps, kPoolSmemBytes);
    return n > 0 ? n : 1;
}
// resident 256-thread blocks per SM of the persistent kernel (for sizing its grid)
int persistent_blocks_per_sm(int accel, int bounce) {
    int n = 0;
    switch (accel * 2 + (bounce ? 1 : 0)) {
        case 0: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_shade_persistent_kernel<0, 0>, 256, 0); break;
        case 1: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_shade_persistent_kernel<0, 1>, 256, 0); break;
        case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_shade_persistent_kernel<1, 0>, 256, 0); break;
        case 3: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_shade_persistent_kernel<1, 1>, 256, 0); break;
        case 4: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_shade_persistent_kernel<2, 0>, 256, 0); break;
        case 5: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_shade_persistent_kernel<2, 1>, 256, 0); break;
        case 6: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_shade_persistent_kernel<3, 0>, 256, 0); break;
        case 7: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_shade_persistent_kernel<3, 1>, 256, 0); break;
        case 8: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_shade_persistent_kernel<4, 0>, 256, 0); break;
        case 9: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_shade_persistent_kernel<4, 1>, 256, 0); break;
        default: break;
    }
    return n > 0 ? n : 1;
}
cudaError_t launch_tile_sort(const uint32_t* cost, uint32_t* order, uint32_t n, uint32_t n_warps, uint32_t split_quarters, uint32_t max_level,
                             uint32_t min_split_cycles, uint32_t* queue_items, cudaStream_t stream) {
    tile_sort_kernel<<<1, 32 * kSortWarps, 0, stream>>>(cost, order, n, n_warps, split_quarters, max_level, min_split_cycles, queue_items);
    return cudaGetLastError();
}
cudaError_t launch_film_clear(float4* sum, float4* sq, uint32_t* ldr, uint32_t* ids, uint32_t n, cudaStream_t stream) {
    film_clear_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(sum, sq, ldr, ids, n);
    return cudaGetLastError();
}
cudaError_t launch_film_variance(const float4* sum, const float4* sq, float* out, uint32_t n, cudaStream_t stream) {
    film_variance_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(sum, sq, out, n);
    return cudaGetLastError();
}
cudaError_t launch_tonemap(const float4* sum, uint32_t* ldr, uint32_t n, cudaStream_t stream) {
    const uint32_t threads = (n + 3u) / 4u;
    tonemap_pack_kernel<<<(threads + 255u) / 256u, 256, 0, stream>>>(sum, ldr, n);
    return cudaGetLastError();
}
cudaError_t launch_flag_signal(uint32_t* flag, uint32_t value, cudaStream_t stream) {
    flag_signal_kernel<<<1, 1, 0, stream>>>(flag, value);
    return cudaGetLastError();
}
cudaError_t launch_flag_signal_wait(uint32_t* signal, uint32_t value, uint32_t* wait_flag, uint32_t target, uint32_t* timeouts, cudaStream_t stream) {
    flag_signal_wait_kernel<<<1, 1, 0, stream>>>(signal, value, wait_flag, target, 2000000000ull, timeouts);
    return cudaGetLastError();
}
cudaError_t launch_flag_wait(uint32_t* flags, uint32_t n, uint32_t target, int signal_slot, int release_slot, uint32_t* timeouts,
                             cudaStream_t stream) {
    if (n == 0) return cudaErrorInvalidValue;
    flag_wait_kernel<<<1, 32, 0, stream>>>(flags, n, target, signal_slot, release_slot, 2000000000ull, timeouts);
    return cudaGetLastError();
}
cudaError_t launch_film_accumulate(const TraceParams& p, cudaStream_t stream) {
    if (p.plane_rows == 0 || p.n_planes == 0) return cudaSuccess;
    dim3 grid((p.cam.width + 255u) / 256u, p.plane_rows);
    film_accumulate_kernel<<<grid, 256, 0, stream>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_gather_rows(const uint32_t* ldr, const uint32_t* row_list, uint32_t n_rows, uint32_t width, uint32_t* out, cudaStream_t stream) {
    if (n_rows == 0) return cudaSuccess;
    dim3 grid((width + 255u) / 256u, n_rows);
    gather_rows_kernel<<<grid, 256, 0, stream>>>(ldr, row_list, n_rows, width, out);
    return cudaGetLastError();
}

}  // namespace rtb
