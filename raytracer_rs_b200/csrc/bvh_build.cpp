// bvh_build.cpp — host build of the flat binary BVH used by the fast traversal kernel.
//
// The reference has no BVH (it ships the octree of oct_tree_intersector.rs); this structure is the B200-first
// replacement the north star asks for. Top-down full-sweep SAH over triangle centroids, leaves of at most
// `max_leaf_size` triangles, both child boxes stored in the parent so one node visit = one 64-byte record.
// Child boxes are padded outward (a few ulps plus 1e-5 of the scene extent) so that box culling can never reject a
// triangle that the reference's Moller-Trumbore arithmetic (intersect.rs:62-98) would accept: the BVH only
// changes WHICH triangles are tested, never the result of a test.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <numeric>

#include "accel_build.h"

namespace rtb {
namespace {

struct Aabb {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    float hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const float* p) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], p[a]);
            hi[a] = std::max(hi[a], p[a]);
        }
    }
    void grow(const Aabb& b) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    float half_area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0.f) return 0.f;
        return dx * dy + dy * dz + dz * dx;
    }
};

struct Builder {
    const HostScene& scene;
    uint32_t max_leaf;
    std::vector<Aabb> tri_box;
    std::vector<float> centroid;  // 3 per triangle
    std::vector<uint32_t> order;  // working permutation
    FlatBvh out;
    float pad_abs = 0.f;

    static constexpr float kTraversalCost = 1.0f;
    float kTriangleCost = 0.8f;  // measured (tools/gpu_bvh_params.py): 0.5-0.8 are 2 % faster than 1.2 on thai2, 4 % on ico3_tex bounce frames; developer override: RT_BVH_TRI_COST

    Builder(const HostScene& s, uint32_t ml) : scene(s), max_leaf(ml) {
        if (const char* e = std::getenv("RT_BVH_TRI_COST")) kTriangleCost = (float)std::atof(e);
    }

    Aabb range_box(uint32_t b, uint32_t e) const {
        Aabb r;
        for (uint32_t i = b; i < e; ++i) r.grow(tri_box[order[i]]);
        return r;
    }

    void store_child(FlatBvh::Node& n, int slot, const Aabb& box, int32_t child, int32_t count) {
        for (int a = 0; a < 3; ++a) {
            // outward padding: relative 2^-20 of the coordinate magnitude plus an absolute scene-scaled margin
            const float plo = std::fabs(box.lo[a]) * 9.5367431640625e-7f + pad_abs;
            const float phi = std::fabs(box.hi[a]) * 9.5367431640625e-7f + pad_abs;
            n.lo[slot][a] = box.lo[a] - plo;
            n.hi[slot][a] = box.hi[a] + phi;
        }
        n.child[slot] = child;
        n.count[slot] = count;
    }

    // returns (child reference, count) for the range [b, e): a leaf reference or the index of a new inner node
    std::pair<int32_t, int32_t> build_range(uint32_t b, uint32_t e, const Aabb& box, uint32_t level) {
        const uint32_t n = e - b;
        if (level > out.depth) out.depth = level;
        uint32_t best_axis = 3, best_split = 0;
        float best_cost = FLT_MAX;
        if (n > 1) {
            std::vector<float> right_area(n);
            const float inv_parent = 1.0f / std::max(box.half_area(), 1e-30f);
            for (uint32_t axis = 0; axis < 3; ++axis) {
                std::sort(order.begin() + b, order.begin() + e, [&](uint32_t x, uint32_t y) {
                    const float cx = centroid[3 * x + axis], cy = centroid[3 * y + axis];
                    return cx < cy || (cx == cy && x < y);
                });
                Aabb acc;
                for (uint32_t i = n; i-- > 1;) {
                    acc.grow(tri_box[order[b + i]]);
                    right_area[i] = acc.half_area();
                }
                acc = Aabb();
                for (uint32_t i = 1; i < n; ++i) {
                    acc.grow(tri_box[order[b + i - 1]]);
                    const float cost = kTraversalCost + kTriangleCost * inv_parent * (acc.half_area() * (float)i + right_area[i] * (float)(n - i));
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_axis = axis;
                        best_split = i;
                    }
                }
            }
        }
        const float leaf_cost = kTriangleCost * (float)n;
        if (n <= max_leaf && (best_axis == 3 || leaf_cost <= best_cost)) {
            const int32_t first = (int32_t)out.tri_order.size();
            std::sort(order.begin() + b, order.begin() + e);  // ascending global index inside a leaf (tie-break order)
            for (uint32_t i = b; i < e; ++i) out.tri_order.push_back(order[i]);
            out.max_leaf = std::max(out.max_leaf, n);
            out.num_leaves++;
            return {~first, (int32_t)n};
        }
        if (best_axis == 3) {  // cannot happen for n > 1, kept for safety
            best_axis = 0;
            best_split = n / 2;
        }
        std::sort(order.begin() + b, order.begin() + e, [&](uint32_t x, uint32_t y) {
            const float cx = centroid[3 * x + best_axis], cy = centroid[3 * y + best_axis];
            return cx < cy || (cx == cy && x < y);
        });
        const uint32_t mid = b + best_split;
        const Aabb lbox = range_box(b, mid), rbox = range_box(mid, e);
        const int32_t idx = (int32_t)out.nodes.size();
        out.nodes.emplace_back();
        auto l = build_range(b, mid, lbox, level + 1);
        auto r = build_range(mid, e, rbox, level + 1);
        FlatBvh::Node& node = out.nodes[idx];
        store_child(node, 0, lbox, l.first, l.second);
        store_child(node, 1, rbox, r.first, r.second);
        return {idx, 0};
    }

    void run() {
        const uint32_t n = scene.num_triangles();
        tri_box.resize(n);
        centroid.resize(3 * (size_t)n);
        order.resize(n);
        std::iota(order.begin(), order.end(), 0u);
        Aabb all;
        for (uint32_t t = 0; t < n; ++t) {
            const float* v = &scene.vertices[9 * (size_t)t];
            for (int c = 0; c < 3; ++c) tri_box[t].grow(v + 3 * c);
            for (int a = 0; a < 3; ++a) centroid[3 * t + a] = 0.5f * (tri_box[t].lo[a] + tri_box[t].hi[a]);
            all.grow(tri_box[t]);
        }
        float extent = 0.f;
        for (int a = 0; a < 3; ++a) {
            out.root_lo[a] = all.lo[a];
            out.root_hi[a] = all.hi[a];
            if (n) extent = std::max(extent, all.hi[a] - all.lo[a]);
        }
        pad_abs = 1e-5f * extent;
        out.nodes.reserve(2 * (size_t)n + 2);
        out.tri_order.reserve(n);
        if (n == 0) {
            out.nodes.emplace_back();
            store_child(out.nodes[0], 0, Aabb(), ~0, 0);
            store_child(out.nodes[0], 1, Aabb(), ~0, 0);
            return;
        }
        auto root = build_range(0, n, all, 0);
        if (root.first < 0) {  // the whole scene fits one leaf: wrap it in a root node with an empty second child
            out.nodes.emplace_back();
            store_child(out.nodes[0], 0, all, root.first, root.second);
            store_child(out.nodes[0], 1, Aabb(), ~0, 0);
        }
    }
};

}  // namespace

FlatBvh build_bvh(const HostScene& scene, uint32_t max_leaf_size) {
    Builder b(scene, std::max(1u, max_leaf_size));
    b.run();
    return std::move(b.out);
}

}  // namespace rtb
