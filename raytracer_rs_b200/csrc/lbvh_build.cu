// lbvh_build.cu — the binary BVH of the trace kernels, built ON THE GPU (the north star allows
// "built on host or as a GPU LBVH"). Replaces, for RT_ACCEL_LBVH, the host SAH build of bvh_build.cpp (which itself
// stands in for OctTreeIntersector::with_triangles_per_leaf, oct_tree_intersector.rs:66-146).
//
//   1. lbvh_prims_kernel   : triangle AABB + 30-bit Morton code of its centre inside the scene AABB
//   2. radix sort          : 4 stable LSD passes of 8 bits (histogram / scan / scatter), values = triangle ids
//   3. lbvh_hierarchy_kernel: Karras 2012, one thread per internal node: range, split, children, parents
//   4. lbvh_refit_kernel   : bottom-up boxes (second arrival at a node continues), writes the 64-byte traversal nodes
//                            with the same outward padding as the host builder; subtrees of <= 4 triangles become
//                            one leaf (their triangles are contiguous in Morton order)
//   5. lbvh_pack_tris_kernel: 48-byte triangle records {v0, e1 = v1-v0, e2 = v2-v0, id} in leaf order, with the same
//                            f32 subtractions the reference performs per ray (intersect.rs:66-67)
// The tree only decides WHICH triangles a ray tests; the hit rules (closest hit, lowest-id tie break, root-cube
// acceptance) live in the traversal, so the image is the same as with the SAH tree.
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "device_types.h"
#include "kernels.h"

namespace rtb {
namespace {

constexpr int kSortChunk = 256;  // keys per (one-warp) block of the radix passes
constexpr uint32_t kLeafMax = 4;

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {  // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void lbvh_prims_kernel(const float* __restrict__ verts, uint32_t n, float3 lo, float3 inv_ext, float4* __restrict__ tri_lo,
                                  float4* __restrict__ tri_hi, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float* v = verts + 9 * (size_t)t;
    const float bx0 = fminf(fminf(v[0], v[3]), v[6]), bx1 = fmaxf(fmaxf(v[0], v[3]), v[6]);
    const float by0 = fminf(fminf(v[1], v[4]), v[7]), by1 = fmaxf(fmaxf(v[1], v[4]), v[7]);
    const float bz0 = fminf(fminf(v[2], v[5]), v[8]), bz1 = fmaxf(fmaxf(v[2], v[5]), v[8]);
    tri_lo[t] = make_float4(bx0, by0, bz0, 0.f);
    tri_hi[t] = make_float4(bx1, by1, bz1, 0.f);
    const float cx = (0.5f * (bx0 + bx1) - lo.x) * inv_ext.x, cy = (0.5f * (by0 + by1) - lo.y) * inv_ext.y,
                cz = (0.5f * (bz0 + bz1) - lo.z) * inv_ext.z;
    const uint32_t qx = (uint32_t)fminf(fmaxf(cx * 1024.0f, 0.0f), 1023.0f);
    const uint32_t qy = (uint32_t)fminf(fmaxf(cy * 1024.0f, 0.0f), 1023.0f);
    const uint32_t qz = (uint32_t)fminf(fmaxf(cz * 1024.0f, 0.0f), 1023.0f);
    keys[t] = (expand_bits10(qx) << 2) | (expand_bits10(qy) << 1) | expand_bits10(qz);
    vals[t] = t;
}

// ---- stable LSD radix sort, 8 bits per pass; one warp per chunk of kSortChunk keys -------------------------------
__global__ void __launch_bounds__(32) radix_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ hist,
                                                        uint32_t n_blocks) {
    __shared__ uint32_t h[256];
    for (int b = threadIdx.x; b < 256; b += 32) h[b] = 0;
    __syncwarp();
    const uint32_t base = blockIdx.x * kSortChunk;
    for (uint32_t i = threadIdx.x; i < kSortChunk && base + i < n; i += 32) atomicAdd(&h[(keys[base + i] >> shift) & 255u], 1u);
    __syncwarp();
    for (int b = threadIdx.x; b < 256; b += 32) hist[(size_t)b * n_blocks + blockIdx.x] = h[b];  // digit-major
}
// exclusive scan of 256 * n_blocks counters (digit-major order = global order of the stable scatter), one block
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* __restrict__ hist, uint32_t total) {
    __shared__ uint32_t part[1024];
    const uint32_t per = (total + 1023u) / 1024u;
    const uint32_t b = threadIdx.x * per, e = min(b + per, total);
    uint32_t s = 0;
    for (uint32_t i = b; i < e; ++i) s += hist[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const uint32_t v = threadIdx.x >= (unsigned)off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - s;
    for (uint32_t i = b; i < e; ++i) {
        const uint32_t c = hist[i];
        hist[i] = run;
        run += c;
    }
}
__global__ void __launch_bounds__(32) radix_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n,
                                                           int shift, const uint32_t* __restrict__ hist, uint32_t n_blocks,
                                                           uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t offs[256];
    const uint32_t lane = threadIdx.x, lt = (1u << lane) - 1u;
    for (int b = lane; b < 256; b += 32) offs[b] = hist[(size_t)b * n_blocks + blockIdx.x];
    __syncwarp();
    const uint32_t base = blockIdx.x * kSortChunk;
    for (uint32_t r = 0; r < kSortChunk && base + r < n; r += 32) {
        const uint32_t i = base + r + lane;
        const bool have = r + lane < kSortChunk && i < n;
        const uint32_t live = __ballot_sync(0xffffffffu, have);
        if (have) {
            const uint32_t k = keys[i], d = (k >> shift) & 255u;
            const uint32_t same = __match_any_sync(live, d);  // lanes of this round with the same digit
            const uint32_t dst = offs[d] + (uint32_t)__popc(same & lt);  // earlier lanes first: stable
            __syncwarp(live);
            if ((same & lt) == 0u) offs[d] += (uint32_t)__popc(same);  // the first lane of every digit group
            keys_out[dst] = k;
            vals_out[dst] = vals[i];
        }
        __syncwarp();
    }
}

// ---- Karras 2012: maximising the common prefix of (code, position) pairs --------------------------------------------
__device__ __forceinline__ int prefix_len(const uint32_t* __restrict__ codes, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint32_t a = codes[i], b = codes[j];
    if (a != b) return __clz(a ^ b);
    return 32 + __clz((uint32_t)i ^ (uint32_t)j);  // equal codes: fall back to the position
}
// internal node i in [0, n-2]; leaves are sorted positions. child encoding here: >= 0 internal, < 0 leaf ~position
__global__ void lbvh_hierarchy_kernel(const uint32_t* __restrict__ codes, int n, int2* __restrict__ children, int2* __restrict__ range,
                                      int* __restrict__ parent_internal, int* __restrict__ parent_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = prefix_len(codes, n, i, i + 1) - prefix_len(codes, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = prefix_len(codes, n, i, i - d);
    int lmax = 2;
    while (prefix_len(codes, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (prefix_len(codes, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = prefix_len(codes, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (prefix_len(codes, n, i, i + (s + t) * d) > dnode) s += t;
        if (t <= 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    int2 ch;
    if (first == gamma) {
        ch.x = ~gamma;
        parent_leaf[gamma] = i;
    } else {
        ch.x = gamma;
        parent_internal[gamma] = i;
    }
    if (last == gamma + 1) {
        ch.y = ~(gamma + 1);
        parent_leaf[gamma + 1] = i;
    } else {
        ch.y = gamma + 1;
        parent_internal[gamma + 1] = i;
    }
    children[i] = ch;
    range[i] = make_int2(first, last);
    if (i == 0) parent_internal[0] = -1;
}

struct Box {
    float lo[3], hi[3];
};
__device__ __forceinline__ Box leaf_box(const float4* tri_lo, const float4* tri_hi, const uint32_t* vals, int pos) {
    const float4 a = tri_lo[vals[pos]], b = tri_hi[vals[pos]];
    return Box{{a.x, a.y, a.z}, {b.x, b.y, b.z}};
}
__device__ __forceinline__ int leaf_ref(int first, int count) { return ~(int)(((uint32_t)first << 4) | (uint32_t)count); }

// One thread per leaf walks towards the root; the second thread to arrive at a node owns it (both children are final).
__global__ void lbvh_refit_kernel(int n, const int2* __restrict__ children, const int2* __restrict__ range, const int* __restrict__ parent_internal,
                                  const int* __restrict__ parent_leaf, const float4* __restrict__ tri_lo, const float4* __restrict__ tri_hi,
                                  const uint32_t* __restrict__ vals, float pad_abs, float4* __restrict__ node_lo, float4* __restrict__ node_hi,
                                  uint32_t* __restrict__ visits, uint32_t* __restrict__ depth, float4* __restrict__ out_nodes) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int node = parent_leaf[leaf];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&visits[node], 1u) == 0u) return;  // the sibling subtree is not finished yet
        __threadfence();
        const int2 ch = children[node];
        Box cb[2];
        int ref[2];
        uint32_t dep = 0;
        for (int k = 0; k < 2; ++k) {
            const int c = k ? ch.y : ch.x;
            if (c < 0) {
                cb[k] = leaf_box(tri_lo, tri_hi, vals, ~c);
                ref[k] = leaf_ref(~c, 1);
            } else {
                const float4 a = __ldcg(&node_lo[c]), b = __ldcg(&node_hi[c]);
                cb[k] = Box{{a.x, a.y, a.z}, {b.x, b.y, b.z}};
                const int2 r = range[c];
                const uint32_t cnt = (uint32_t)(r.y - r.x + 1);
                if (cnt <= kLeafMax) {
                    ref[k] = leaf_ref(r.x, (int)cnt);  // a small subtree is one leaf: its triangles are contiguous
                } else {
                    ref[k] = c;
                    dep = max(dep, __ldcg(&depth[c]));
                }
            }
        }
        float4 q[3];
        float* f = reinterpret_cast<float*>(q);  // lo0 xyz hi0 xyz lo1 xyz hi1 xyz
        for (int k = 0; k < 2; ++k)
            for (int a = 0; a < 3; ++a) {
                // same outward padding as the host builder (bvh_build.cpp store_child)
                f[6 * k + a] = cb[k].lo[a] - (fabsf(cb[k].lo[a]) * 9.5367431640625e-7f + pad_abs);
                f[6 * k + 3 + a] = cb[k].hi[a] + (fabsf(cb[k].hi[a]) * 9.5367431640625e-7f + pad_abs);
            }
        out_nodes[4 * (size_t)node + 0] = q[0];
        out_nodes[4 * (size_t)node + 1] = q[1];
        out_nodes[4 * (size_t)node + 2] = q[2];
        out_nodes[4 * (size_t)node + 3] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), 0.f, 0.f);
        node_lo[node] = make_float4(fminf(cb[0].lo[0], cb[1].lo[0]), fminf(cb[0].lo[1], cb[1].lo[1]), fminf(cb[0].lo[2], cb[1].lo[2]), 0.f);
        node_hi[node] = make_float4(fmaxf(cb[0].hi[0], cb[1].hi[0]), fmaxf(cb[0].hi[1], cb[1].hi[1]), fmaxf(cb[0].hi[2], cb[1].hi[2]), 0.f);
        depth[node] = dep + 1u;
        node = parent_internal[node];
    }
}

// scenes of 0 or 1 triangles: a root whose children are (the triangle | nothing) and nothing
__global__ void lbvh_tiny_root_kernel(int n, const float4* __restrict__ tri_lo, const float4* __restrict__ tri_hi, float pad_abs,
                                      float4* __restrict__ out_nodes, uint32_t* __restrict__ depth) {
    float f[12];
    for (int k = 0; k < 2; ++k)
        for (int a = 0; a < 3; ++a) {
            f[6 * k + a] = FLT_MAX;
            f[6 * k + 3 + a] = -FLT_MAX;
        }
    if (n == 1) {
        const float lo[3] = {tri_lo[0].x, tri_lo[0].y, tri_lo[0].z}, hi[3] = {tri_hi[0].x, tri_hi[0].y, tri_hi[0].z};
        for (int a = 0; a < 3; ++a) {
            f[a] = lo[a] - (fabsf(lo[a]) * 9.5367431640625e-7f + pad_abs);
            f[3 + a] = hi[a] + (fabsf(hi[a]) * 9.5367431640625e-7f + pad_abs);
        }
    }
    out_nodes[0] = make_float4(f[0], f[1], f[2], f[3]);
    out_nodes[1] = make_float4(f[4], f[5], f[6], f[7]);
    out_nodes[2] = make_float4(f[8], f[9], f[10], f[11]);
    out_nodes[3] = make_float4(__int_as_float(leaf_ref(0, n == 1 ? 1 : 0)), __int_as_float(leaf_ref(0, 0)), 0.f, 0.f);
    depth[0] = 1u;
}

__global__ void lbvh_pack_tris_kernel(const float* __restrict__ verts, const uint32_t* __restrict__ vals, uint32_t n, float4* __restrict__ out) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t t = vals[s];
    const float* v = verts + 9 * (size_t)t;
    // e1 = v1 - v0, e2 = v2 - v0: the subtractions intersect.rs:66-67 performs per ray (round to nearest, never fused)
    out[3 * (size_t)s + 0] = make_float4(v[0], v[1], v[2], __fsub_rn(v[3], v[0]));
    out[3 * (size_t)s + 1] = make_float4(__fsub_rn(v[4], v[1]), __fsub_rn(v[5], v[2]), __fsub_rn(v[6], v[0]), __fsub_rn(v[7], v[1]));
    out[3 * (size_t)s + 2] = make_float4(__fsub_rn(v[8], v[2]), __uint_as_float(t), 0.f, 0.f);
}

}  // namespace

size_t lbvh_scratch_bytes(uint32_t n) {
    const size_t nn = n ? n : 1;
    const size_t n_blocks = (nn + kSortChunk - 1) / kSortChunk;
    // tri_lo, tri_hi, node_lo, node_hi (float4) | keys x2, vals x2 | hist | children, range (int2) | parents x2 | visits | depth
    return 4 * nn * 16 + 4 * nn * 4 + 256 * n_blocks * 4 + 2 * nn * 8 + 2 * nn * 4 + 2 * nn * 4 + 256;
}

// Builds nodes (4 float4 per node, max(n - 1, 1) nodes, root = node 0), triangle records (3 float4 per triangle, leaf
// order) and the slot -> triangle table on `stream`. `scratch` must hold lbvh_scratch_bytes(n) bytes. d_depth receives
// the depth of the tree (the caller checks it against the traversal stack).
cudaError_t build_lbvh_device(const float* d_verts, uint32_t n, const float root_lo[3], const float root_hi[3], void* scratch, float4* d_nodes,
                              float4* d_tris, uint32_t* d_tri_order, uint32_t* d_depth, cudaStream_t stream) {
    const size_t nn = n ? n : 1;
    const uint32_t n_blocks = (uint32_t)((nn + kSortChunk - 1) / kSortChunk);
    char* p = static_cast<char*>(scratch);
    auto take = [&](size_t bytes) {
        char* r = p;
        p += (bytes + 15) & ~size_t(15);
        return r;
    };
    float4* tri_lo = (float4*)take(nn * 16);
    float4* tri_hi = (float4*)take(nn * 16);
    float4* node_lo = (float4*)take(nn * 16);
    float4* node_hi = (float4*)take(nn * 16);
    uint32_t* keys_a = (uint32_t*)take(nn * 4);
    uint32_t* keys_b = (uint32_t*)take(nn * 4);
    uint32_t* vals_a = (uint32_t*)take(nn * 4);
    uint32_t* vals_c = (uint32_t*)take(nn * 4);
    uint32_t* hist = (uint32_t*)take(256 * (size_t)n_blocks * 4);
    int2* children = (int2*)take(nn * 8);
    int2* range = (int2*)take(nn * 8);
    int* parent_internal = (int*)take(nn * 4);
    int* parent_leaf = (int*)take(nn * 4);
    uint32_t* visits = (uint32_t*)take(nn * 4);

    float ext = 0.f;
    for (int a = 0; a < 3; ++a) ext = fmaxf(ext, root_hi[a] - root_lo[a]);
    const float pad_abs = n ? 1e-5f * ext : 0.f;  // bvh_build.cpp: pad_abs = 1e-5 * scene extent
    const float3 lo = make_float3(root_lo[0], root_lo[1], root_lo[2]);
    const float3 inv_ext = make_float3(root_hi[0] > root_lo[0] ? 1.0f / (root_hi[0] - root_lo[0]) : 0.f,
                                       root_hi[1] > root_lo[1] ? 1.0f / (root_hi[1] - root_lo[1]) : 0.f,
                                       root_hi[2] > root_lo[2] ? 1.0f / (root_hi[2] - root_lo[2]) : 0.f);
    cudaError_t e;
    if (n == 0) {
        lbvh_tiny_root_kernel<<<1, 1, 0, stream>>>(0, tri_lo, tri_hi, pad_abs, d_nodes, d_depth);
        return cudaGetLastError();
    }
    lbvh_prims_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(d_verts, n, lo, inv_ext, tri_lo, tri_hi, keys_a, vals_a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    uint32_t *kin = keys_a, *kout = keys_b, *vin = vals_a, *vout = vals_c;
    for (int pass = 0; pass < 4; ++pass) {
        radix_hist_kernel<<<n_blocks, 32, 0, stream>>>(kin, n, 8 * pass, hist, n_blocks);
        radix_scan_kernel<<<1, 1024, 0, stream>>>(hist, 256u * n_blocks);
        radix_scatter_kernel<<<n_blocks, 32, 0, stream>>>(kin, vin, n, 8 * pass, hist, n_blocks, kout, vout);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        uint32_t* t = kin;
        kin = kout;
        kout = t;
        t = vin;
        vin = vout;
        vout = t;
    }
    // after four passes the sorted codes are in `kin` (= keys_a) and the triangle ids in `vin` (= vals_a)
    if ((e = cudaMemcpyAsync(d_tri_order, vin, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
    lbvh_pack_tris_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(d_verts, vin, n, d_tris);
    if (n == 1) {
        lbvh_tiny_root_kernel<<<1, 1, 0, stream>>>(1, tri_lo, tri_hi, pad_abs, d_nodes, d_depth);
        return cudaGetLastError();
    }
    if ((e = cudaMemsetAsync(visits, 0, (size_t)n * 4, stream)) != cudaSuccess) return e;
    lbvh_hierarchy_kernel<<<(n + 254u) / 256u, 256, 0, stream>>>(kin, (int)n, children, range, parent_internal, parent_leaf);
    // the per-node depth array doubles as the output: depth of node 0 = depth of the tree
    uint32_t* depth = (uint32_t*)take(nn * 4);
    lbvh_refit_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>((int)n, children, range, parent_internal, parent_leaf, tri_lo, tri_hi, vin, pad_abs,
                                                            node_lo, node_hi, visits, depth, d_nodes);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return cudaMemcpyAsync(d_depth, depth, 4, cudaMemcpyDeviceToDevice, stream);
}

}  // namespace rtb
