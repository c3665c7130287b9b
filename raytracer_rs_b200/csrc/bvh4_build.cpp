// bvh4_build.cpp — 4-wide BVH with full-precision child boxes, collapsed from the binary SAH BVH of bvh_build.cpp.
//
// Why: on the cache-resident scenes of this renderer a ray's time is the length of its chain of dependent node
// visits. A 4-wide node halves that chain for about the same number of instructions per box, and
// one node is exactly one 128-byte line: six float4 (child boxes, structure of arrays: lo.x[4] lo.y[4] lo.z[4] hi.x[4]
// hi.y[4] hi.z[4]) + one int4 of child references + 16 bytes of padding.
// The collapse is the optimal one of Ylitie et al. (HPG 2017, section 3.1) for width 4: cost[n][i] = least SAH cost of
// representing the subtree of binary node n by at most i children of a wide node.
// The structure only decides WHICH triangles are tested; boxes are the binary builder's padded boxes.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "accel_build.h"

namespace rtb {
namespace {

struct Ref4 {
    float lo[3], hi[3];
    int32_t child;  // FlatBvh reference: >= 0 binary inner node, < 0 leaf (~first)
    int32_t count;
};

struct Builder4 {
    static constexpr int kWidth = 4;
    uint32_t kMaxLeaf = 4;                  // developer override: RT_BVH4_MAX_LEAF (<= 15)
    static constexpr float kNodeCost = 1.0f;
    float kPrimCost = 0.6f;                 // developer override: RT_BVH4_PRIM_COST
    enum : uint8_t { kLeaf = 0, kInternal = 1, kDistribute = 2, kFewer = 3 };
    struct Dp {
        float cost[kWidth];            // [1 .. kWidth-1]
        uint8_t choice[kWidth];        // [1 .. kWidth-1]
        uint8_t split[kWidth + 1];     // [2 .. kWidth]: slots given to child 0 when j are distributed
    };
    const FlatBvh& bvh;
    FlatBvh4 out;
    std::vector<uint32_t> sub_first, sub_count;
    std::vector<Dp> dp;

    explicit Builder4(const FlatBvh& b) : bvh(b), sub_first(b.nodes.size(), 0u), sub_count(b.nodes.size(), 0u), dp(b.nodes.size()) {
        if (const char* e = std::getenv("RT_BVH4_MAX_LEAF")) kMaxLeaf = (uint32_t)std::min(15, std::max(1, std::atoi(e)));
        if (const char* e = std::getenv("RT_BVH4_PRIM_COST")) kPrimCost = (float)std::atof(e);
        solve(0);
    }

    static float box_area(const float* lo, const float* hi) {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0.f) return 0.f;
        return dx * dy + dy * dz + dz * dx;
    }

    void solve(int32_t node) {
        const FlatBvh::Node& n = bvh.nodes[node];
        uint32_t first = 0xffffffffu, count = 0;
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int k = 0; k < 2; ++k) {
            const int32_t c = n.child[k];
            if (c >= 0) {
                solve(c);
                first = std::min(first, sub_first[c]);
                count += sub_count[c];
            } else if (n.count[k] > 0) {
                first = std::min(first, (uint32_t)(~c));
                count += (uint32_t)n.count[k];
            }
            if (c >= 0 || n.count[k] > 0)
                for (int a = 0; a < 3; ++a) {
                    lo[a] = std::min(lo[a], n.lo[k][a]);
                    hi[a] = std::max(hi[a], n.hi[k][a]);
                }
        }
        sub_first[node] = first;
        sub_count[node] = count;
        const float area = box_area(lo, hi);
        auto child_cost = [&](int k, int i) {
            if (n.child[k] >= 0) return dp[n.child[k]].cost[i];
            return box_area(n.lo[k], n.hi[k]) * (float)n.count[k] * kPrimCost;
        };
        Dp& d = dp[node];
        float dist[kWidth + 1];
        for (int j = 2; j <= kWidth; ++j) {
            dist[j] = FLT_MAX;
            d.split[j] = 1;
            for (int a = 1; a < j; ++a) {
                if (a > kWidth - 1 || j - a > kWidth - 1) continue;
                const float c = child_cost(0, a) + child_cost(1, j - a);
                if (c < dist[j]) {
                    dist[j] = c;
                    d.split[j] = (uint8_t)a;
                }
            }
        }
        const float internal = dist[kWidth] + area * kNodeCost;
        const float leaf = count <= kMaxLeaf ? area * (float)count * kPrimCost : FLT_MAX;
        d.cost[1] = std::min(leaf, internal);
        d.choice[1] = leaf <= internal ? kLeaf : kInternal;
        for (int i = 2; i <= kWidth - 1; ++i) {
            if (dist[i] < d.cost[i - 1]) {
                d.cost[i] = dist[i];
                d.choice[i] = kDistribute;
            } else {
                d.cost[i] = d.cost[i - 1];
                d.choice[i] = kFewer;
            }
        }
    }

    void gather(int32_t node, int budget, std::vector<Ref4>& kids) const {
        const FlatBvh::Node& n = bvh.nodes[node];
        const int a = dp[node].split[budget];
        emit(n, 0, a, kids);
        emit(n, 1, budget - a, kids);
    }
    void emit(const FlatBvh::Node& parent, int k, int budget, std::vector<Ref4>& kids) const {
        Ref4 r;
        std::memcpy(r.lo, parent.lo[k], 12);
        std::memcpy(r.hi, parent.hi[k], 12);
        r.child = parent.child[k];
        r.count = parent.count[k];
        if (r.child < 0) {
            if (r.count > 0) kids.push_back(r);
            return;
        }
        const Dp& d = dp[r.child];
        while (budget > 1 && d.choice[budget] == kFewer) --budget;
        if (budget > 1) {
            gather(r.child, budget, kids);
        } else if (d.choice[1] == kLeaf) {
            r.count = (int32_t)sub_count[r.child];
            r.child = ~(int32_t)sub_first[r.child];
            kids.push_back(r);
        } else {
            kids.push_back(r);
        }
    }

    void fill(uint32_t index, int32_t bnode, uint32_t level) {
        out.depth = std::max(out.depth, level);
        std::vector<Ref4> kids;
        gather(bnode, kWidth, kids);
        FlatBvh4::Node node;
        const float inf = std::numeric_limits<float>::infinity();
        int32_t inner_bnode[kWidth];
        uint32_t inner_slot[kWidth], n_inner = 0;
        for (int s = 0; s < kWidth; ++s) {
            if (s < (int)kids.size()) {
                const Ref4& r = kids[s];
                std::memcpy(node.lo[s], r.lo, 12);
                std::memcpy(node.hi[s], r.hi, 12);
                if (r.child >= 0) {
                    inner_bnode[n_inner] = r.child;
                    inner_slot[n_inner++] = (uint32_t)s;
                    node.child[s] = 0;  // patched below
                    node.count[s] = 0;
                } else {
                    // leaf: copy its triangles (ascending global index inside a leaf, as in the binary tree's merged run)
                    const uint32_t first = (uint32_t)(~r.child), slot = (uint32_t)out.tri_order.size();
                    std::vector<uint32_t> tris(bvh.tri_order.begin() + first, bvh.tri_order.begin() + first + r.count);
                    std::sort(tris.begin(), tris.end());
                    out.tri_order.insert(out.tri_order.end(), tris.begin(), tris.end());
                    node.child[s] = ~(int32_t)slot;
                    node.count[s] = r.count;
                    out.max_leaf = std::max(out.max_leaf, (uint32_t)r.count);
                    out.num_leaves++;
                }
            } else {  // empty slot: a box no ray can hit (lo = hi = +inf), see the traversal kernel
                for (int a = 0; a < 3; ++a) node.lo[s][a] = node.hi[s][a] = inf;
                node.child[s] = ~0;
                node.count[s] = 0;
            }
        }
        const uint32_t child_base = (uint32_t)out.nodes.size();
        for (uint32_t k = 0; k < n_inner; ++k) node.child[inner_slot[k]] = (int32_t)(child_base + k);
        out.nodes[index] = node;
        out.nodes.resize(out.nodes.size() + n_inner);
        for (uint32_t k = 0; k < n_inner; ++k) fill(child_base + k, inner_bnode[k], level + 1);
    }
};

}  // namespace

FlatBvh4 build_bvh4(const HostScene& scene) {
    uint32_t base_leaf = 4;
    if (const char* e = std::getenv("RT_BVH4_BASE_LEAF")) base_leaf = (uint32_t)std::min(15, std::max(1, std::atoi(e)));  // developer override
    const FlatBvh bvh = build_bvh(scene, base_leaf);
    Builder4 b(bvh);
    b.out.nodes.resize(1);
    b.out.tri_order.reserve(scene.num_triangles());
    std::memcpy(b.out.root_lo, bvh.root_lo, 12);
    std::memcpy(b.out.root_hi, bvh.root_hi, 12);
    b.fill(0, 0, 0);
    return std::move(b.out);
}

}  // namespace rtb
