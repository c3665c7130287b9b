// cwbvh_build.cpp — compressed 8-wide BVH (after Ylitie, Karras, Laine: "Efficient Incoherent Ray Traversal on GPUs
// Through Compressed Wide BVHs", HPG 2017), built on the host by collapsing the binary SAH BVH of bvh_build.cpp.
//
// Why: the binary-BVH kernel is bound by the chain of dependent node fetches: a heavy ray does
// ~150 node steps, each waiting for memory. An 8-wide node cuts the chain ~3x, and quantising the child boxes to
// 8 bits relative to the node (80 bytes for 8 children) makes the whole node set of thai2 fit the L1 cache.
// The structure only decides WHICH triangles are tested; every triangle test is the reference's Moller-Trumbore
// arithmetic, so results are unchanged. Quantised boxes are rounded outward and verified with the same float
// expressions the traversal evaluates.
//
// Node (5 x 16 bytes):
//   w0: origin p.xyz (f32) | ex, ey, ez (biased exponent bytes, scale = 2^(e-127)), imask (bit i: child i is a node)
//   w1: first child node index | first triangle index | meta[0..3] | meta[4..7]
//   w2: qlo_x[0..3] qlo_x[4..7] qlo_y[0..3] qlo_y[4..7]
//   w3: qlo_z[0..3] qlo_z[4..7] qhi_x[0..3] qhi_x[4..7]
//   w4: qhi_y[0..3] qhi_y[4..7] qhi_z[0..3] qhi_z[4..7]
//   meta: 0 = empty; node child: 0b001_11sss (sss = slot); leaf child: unary triangle count (1, 3, 7) << 5 | offset from
//   the node's first triangle (0..23).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

#include "accel_build.h"

namespace rtb {
namespace {

struct Ref {
    float lo[3], hi[3];
    int32_t child;  // FlatBvh reference: >= 0 inner node, < 0 leaf
    int32_t count;
    float area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

struct WideBuilder {
    const FlatBvh& bvh;
    FlatCwbvh out;

    // triangles below every binary inner node: leaves are emitted depth first, so a subtree owns one contiguous run
    std::vector<uint32_t> sub_first, sub_count;

    // Optimal collapse of the binary tree (Ylitie et al., section 3.1): cost[n][i] = least SAH cost of representing
    // the subtree of binary node n by at most i children of a wide node. i = 1: either one leaf child (<= kMaxLeaf
    // triangles) or one inner wide node whose 8 slots are distributed over n's two children.
    static constexpr uint32_t kMaxLeaf = 3;  // unary count in 3 bits
    static constexpr float kNodeCost = 1.0f, kPrimCost = 0.3f;
    enum : uint8_t { kLeaf = 0, kInternal = 1, kDistribute = 2, kFewer = 3 };
    struct Dp {
        float cost[8];     // [1..7]
        uint8_t choice[8]; // [1..7]
        uint8_t split[9];  // [2..8]: roots given to child 0 when j are distributed
    };
    std::vector<Dp> dp;

    explicit WideBuilder(const FlatBvh& b) : bvh(b), sub_first(b.nodes.size(), 0u), sub_count(b.nodes.size(), 0u), dp(b.nodes.size()) {
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
        float dist[9];
        for (int j = 2; j <= 8; ++j) {
            dist[j] = FLT_MAX;
            d.split[j] = 1;
            for (int a = 1; a < j && a <= 7; ++a) {
                if (j - a > 7) continue;
                const float c = child_cost(0, a) + child_cost(1, j - a);
                if (c < dist[j]) {
                    dist[j] = c;
                    d.split[j] = (uint8_t)a;
                }
            }
        }
        const float internal = dist[8] + area * kNodeCost;
        const float leaf = count <= kMaxLeaf ? area * (float)count * kPrimCost : FLT_MAX;
        d.cost[1] = std::min(leaf, internal);
        d.choice[1] = leaf <= internal ? kLeaf : kInternal;
        for (int i = 2; i <= 7; ++i) {
            if (dist[i] < d.cost[i - 1]) {
                d.cost[i] = dist[i];
                d.choice[i] = kDistribute;
            } else {
                d.cost[i] = d.cost[i - 1];
                d.choice[i] = kFewer;
            }
        }
    }

    // children of the wide node that replaces binary node `node`, which may use up to `budget` slots
    void gather(int32_t node, int budget, std::vector<Ref>& kids) const {
        const FlatBvh::Node& n = bvh.nodes[node];
        const int a = dp[node].split[budget];
        emit(n, 0, a, kids);
        emit(n, 1, budget - a, kids);
    }
    void emit(const FlatBvh::Node& parent, int k, int budget, std::vector<Ref>& kids) const {
        Ref r;
        std::memcpy(r.lo, parent.lo[k], 12);
        std::memcpy(r.hi, parent.hi[k], 12);
        r.child = parent.child[k];
        r.count = parent.count[k];
        if (r.child < 0) {
            kids.push_back(r);
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

    // fills wide node `index` from the binary subtree rooted at inner node `bnode`
    void fill(uint32_t index, int32_t bnode, uint32_t level) {
        out.depth = std::max(out.depth, level);
        // 1. children chosen by the optimal collapse
        std::vector<Ref> kids;
        gather(bnode, 8, kids);
        kids.erase(std::remove_if(kids.begin(), kids.end(), [](const Ref& r) { return r.child < 0 && r.count == 0; }), kids.end());

        // 2. node box and quantisation frame
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (const Ref& r : kids)
            for (int a = 0; a < 3; ++a) {
                lo[a] = std::min(lo[a], r.lo[a]);
                hi[a] = std::max(hi[a], r.hi[a]);
            }
        if (kids.empty()) lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0.f;
        int e[3];
        for (int a = 0; a < 3; ++a) {
            const float ext = std::max(hi[a] - lo[a], 1e-30f);
            e[a] = (int)std::ceil(std::log2(ext / 255.0f));
            e[a] = std::min(std::max(e[a], -126), 127);
        }

        // 3. slot assignment: greedy maximum of (centroid offset) . (slot direction); slot s is visited first by rays
        //    whose direction is negative on the axes whose bit is set in s
        float nc[3];
        for (int a = 0; a < 3; ++a) nc[a] = 0.5f * (lo[a] + hi[a]);
        int slot_of[8];
        bool slot_used[8] = {false, false, false, false, false, false, false, false};
        std::vector<bool> kid_done(kids.size(), false);
        for (size_t round = 0; round < kids.size(); ++round) {
            float bestc = -FLT_MAX;
            int bk = -1, bs = -1;
            for (size_t k = 0; k < kids.size(); ++k) {
                if (kid_done[k]) continue;
                for (int s = 0; s < 8; ++s) {
                    if (slot_used[s]) continue;
                    float c = 0.f;
                    for (int a = 0; a < 3; ++a) c += (0.5f * (kids[k].lo[a] + kids[k].hi[a]) - nc[a]) * (((s >> a) & 1) ? 1.f : -1.f);
                    if (c > bestc) {
                        bestc = c;
                        bk = (int)k;
                        bs = s;
                    }
                }
            }
            kid_done[bk] = true;
            slot_used[bs] = true;
            slot_of[bk] = bs;
        }
        int kid_in_slot[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
        for (size_t k = 0; k < kids.size(); ++k) kid_in_slot[slot_of[k]] = (int)k;

        // 4. quantise child boxes outward; grow the exponent if a coordinate does not fit 8 bits
        uint8_t q[6][8];
        for (int a = 0; a < 3; ++a) {
            for (;;) {
                const float scale = std::ldexp(1.0f, e[a]);
                bool fits = true;
                for (int s = 0; s < 8 && fits; ++s) {
                    q[a][s] = 0;
                    q[3 + a][s] = 0;
                    if (kid_in_slot[s] < 0) continue;
                    const Ref& r = kids[kid_in_slot[s]];
                    int ql = (int)std::floor((r.lo[a] - lo[a]) / scale);
                    int qh = (int)std::ceil((r.hi[a] - lo[a]) / scale);
                    ql = std::max(ql, 0);
                    // verify with the float expression the traversal uses: plane = p + q * scale
                    while (ql > 0 && lo[a] + (float)ql * scale > r.lo[a]) --ql;
                    while (lo[a] + (float)qh * scale < r.hi[a]) ++qh;
                    if (qh > 255) {
                        fits = false;
                        break;
                    }
                    q[a][s] = (uint8_t)ql;
                    q[3 + a][s] = (uint8_t)qh;
                }
                if (fits) break;
                ++e[a];
            }
        }

        // 5. children: nodes get consecutive indices in slot order, leaf triangles go to one contiguous run
        uint8_t imask = 0, meta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const uint32_t child_base = (uint32_t)out.nodes.size() / 5;
        const uint32_t tri_base = (uint32_t)out.tri_order.size();
        uint32_t n_inner = 0, tri_off = 0;
        int32_t inner_bnode[8];
        for (int s = 0; s < 8; ++s) {
            if (kid_in_slot[s] < 0) continue;
            const Ref& r = kids[kid_in_slot[s]];
            if (r.child >= 0) {
                imask |= (uint8_t)(1u << s);
                meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
                inner_bnode[n_inner++] = r.child;
            } else {
                const uint32_t first = (uint32_t)(~r.child);
                meta[s] = (uint8_t)((((1u << r.count) - 1u) << 5) | tri_off);
                for (int t = 0; t < r.count; ++t) out.tri_order.push_back(bvh.tri_order[first + t]);
                tri_off += (uint32_t)r.count;
                out.num_leaves++;
            }
        }
        out.nodes.resize(out.nodes.size() + 5 * (size_t)n_inner);

        // 6. pack
        auto word = [](const uint8_t* b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); };
        auto f2u = [](float f) {
            uint32_t u;
            std::memcpy(&u, &f, 4);
            return u;
        };
        CwWord* w = &out.nodes[5 * (size_t)index];
        w[0] = CwWord{f2u(lo[0]), f2u(lo[1]), f2u(lo[2]),
                      (uint32_t)(e[0] + 127) | ((uint32_t)(e[1] + 127) << 8) | ((uint32_t)(e[2] + 127) << 16) | ((uint32_t)imask << 24)};
        w[1] = CwWord{child_base, tri_base, word(meta), word(meta + 4)};
        w[2] = CwWord{word(q[0]), word(q[0] + 4), word(q[1]), word(q[1] + 4)};
        w[3] = CwWord{word(q[2]), word(q[2] + 4), word(q[3]), word(q[3] + 4)};
        w[4] = CwWord{word(q[4]), word(q[4] + 4), word(q[5]), word(q[5] + 4)};

        for (uint32_t k = 0; k < n_inner; ++k) fill(child_base + k, inner_bnode[k], level + 1);
    }
};

}  // namespace

FlatCwbvh build_cwbvh(const HostScene& scene) {
    const FlatBvh bvh = build_bvh(scene, 3);  // a leaf child carries at most 3 triangles (unary count in 3 bits)
    WideBuilder b(bvh);
    b.out.nodes.resize(5);
    b.out.tri_order.reserve(scene.num_triangles());
    std::memcpy(b.out.root_lo, bvh.root_lo, 12);
    std::memcpy(b.out.root_hi, bvh.root_hi, 12);
    b.fill(0, 0, 0);
    return std::move(b.out);
}

}  // namespace rtb
