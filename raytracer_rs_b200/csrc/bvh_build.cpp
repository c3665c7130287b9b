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
#include <cstdio>
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

// ------------------------------------------------------------------------------------------------------
// Insertion-based optimisation of the finished tree (after Bittner, Hapala, Havran: "Fast insertion-based optimization of
// bounding volume hierarchies", CGF 2013): a subtree is cut out together with its parent (the sibling moves up), the position
// where putting it back adds the least surface area is found by branch and bound over the whole tree, and the freed parent
// node is re-used there. The position it came from is one of the candidates, so the summed area of the inner nodes — the
// part of the SAH cost a fixed set of leaves leaves open — never grows. Leaves (and with them tri_order and the tie-break
// order inside a leaf) are untouched; only which boxes a ray has to walk through changes, never the outcome of a triangle test.
// ------------------------------------------------------------------------------------------------------
struct Reinserter {
    // ids [0, n_inner) are inner nodes, [n_inner, n_inner + n_leaf) leaves
    std::vector<Aabb> box;
    std::vector<int32_t> parent, kid[2];
    std::vector<int32_t> leaf_ref, leaf_count;  // per leaf id - n_inner
    int32_t n_inner = 0, root = 0;

    static Aabb unite(const Aabb& a, const Aabb& b) {
        Aabb r = a;
        r.grow(b);
        return r;
    }
    bool is_leaf(int32_t id) const { return id >= n_inner; }

    void load(const FlatBvh& t) {
        n_inner = (int32_t)t.nodes.size();
        size_t leaves = 0;
        for (const auto& n : t.nodes) leaves += (n.child[0] < 0) + (n.child[1] < 0);
        const size_t total = (size_t)n_inner + leaves;
        box.assign(total, Aabb());
        parent.assign(total, -1);
        kid[0].assign(total, -1);
        kid[1].assign(total, -1);
        int32_t next_leaf = n_inner;
        for (int32_t i = 0; i < n_inner; ++i) {
            const FlatBvh::Node& n = t.nodes[(size_t)i];
            for (int k = 0; k < 2; ++k) {
                int32_t id = n.child[k];
                if (id < 0) {
                    id = next_leaf++;
                    leaf_ref.push_back(n.child[k]);
                    leaf_count.push_back(n.count[k]);
                }
                kid[k][(size_t)i] = id;
                parent[(size_t)id] = i;
                for (int a = 0; a < 3; ++a) {  // the stored (padded) child box
                    box[(size_t)id].lo[a] = n.lo[k][a];
                    box[(size_t)id].hi[a] = n.hi[k][a];
                }
            }
        }
        box[0] = unite(box[(size_t)kid[0][0]], box[(size_t)kid[1][0]]);
        root = 0;
    }

    double inner_area() const {
        double s = 0.0;
        for (int32_t i = 0; i < n_inner; ++i) s += box[(size_t)i].half_area();
        return s;
    }

    void refit_up(int32_t id) {
        while (id >= 0) {
            box[(size_t)id] = unite(box[(size_t)kid[0][(size_t)id]], box[(size_t)kid[1][(size_t)id]]);
            id = parent[(size_t)id];
        }
    }

    // best node to pair `n` with (n is detached): minimises area(new parent) + the growth of every ancestor
    int32_t find_target(int32_t n) {
        const Aabb& nb = box[(size_t)n];
        const float n_area = nb.half_area();
        float best = FLT_MAX;
        int32_t best_id = root;
        std::vector<std::pair<float, int32_t>> heap;  // (induced cost, node), smallest induced cost first
        auto cmp = [](const std::pair<float, int32_t>& a, const std::pair<float, int32_t>& b) { return a.first > b.first; };
        heap.emplace_back(0.f, root);
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            const auto [induced, x] = heap.back();
            heap.pop_back();
            if (induced + n_area >= best) break;  // nothing cheaper is left
            const float direct = unite(box[(size_t)x], nb).half_area();
            const float total = induced + direct;
            if (total < best) {
                best = total;
                best_id = x;
            }
            if (!is_leaf(x)) {
                const float child_induced = total - box[(size_t)x].half_area();
                if (child_induced + n_area < best)
                    for (int k = 0; k < 2; ++k) {
                        heap.emplace_back(child_induced, kid[k][(size_t)x]);
                        std::push_heap(heap.begin(), heap.end(), cmp);
                    }
            }
        }
        return best_id;
    }

    // cut n (and its parent p) out, put it back at the best place; returns false when n cannot move (child of the root)
    bool reinsert(int32_t n) {
        const int32_t p = parent[(size_t)n];
        if (p < 0 || p == root) return false;
        const int32_t g = parent[(size_t)p];
        const int side = kid[0][(size_t)p] == n ? 0 : 1;
        const int32_t s = kid[1 - side][(size_t)p];
        kid[kid[0][(size_t)g] == p ? 0 : 1][(size_t)g] = s;
        parent[(size_t)s] = g;
        refit_up(g);
        const int32_t x = find_target(n);
        const int32_t xp = parent[(size_t)x];
        kid[0][(size_t)p] = x;
        kid[1][(size_t)p] = n;
        parent[(size_t)x] = p;
        parent[(size_t)n] = p;
        parent[(size_t)p] = xp;
        if (xp < 0) root = p;
        else kid[kid[0][(size_t)xp] == x ? 0 : 1][(size_t)xp] = p;
        refit_up(p);
        return true;
    }

    uint32_t depth_of(int32_t id) const {  // levels of inner nodes below and including id (iterative: the tree may be deep)
        uint32_t deepest = 0;
        std::vector<std::pair<int32_t, uint32_t>> st{{id, 0u}};
        while (!st.empty()) {
            const auto [x, d] = st.back();
            st.pop_back();
            if (is_leaf(x)) continue;
            deepest = std::max(deepest, d);
            st.emplace_back(kid[0][(size_t)x], d + 1);
            st.emplace_back(kid[1][(size_t)x], d + 1);
        }
        return deepest;
    }

    void store(FlatBvh& t) const {
        std::vector<int32_t> new_index((size_t)n_inner, -1), order;
        order.reserve((size_t)n_inner);
        order.push_back(root);
        new_index[(size_t)root] = 0;
        for (size_t q = 0; q < order.size(); ++q)
            for (int k = 0; k < 2; ++k) {
                const int32_t c = kid[k][(size_t)order[q]];
                if (!is_leaf(c)) {
                    new_index[(size_t)c] = (int32_t)order.size();
                    order.push_back(c);
                }
            }
        // triangle slots are handed out again in depth-first order, so that every subtree owns one contiguous run of tri_order (the 4- and
        // 8-wide builders merge subtrees into leaves by run)
        std::vector<int32_t> new_ref(leaf_ref.size(), 0);
        std::vector<uint32_t> tri_order;
        tri_order.reserve(t.tri_order.size());
        {
            std::vector<int32_t> st{root};
            while (!st.empty()) {
                const int32_t x = st.back();
                st.pop_back();
                if (is_leaf(x)) {
                    const size_t l = (size_t)(x - n_inner);
                    const uint32_t first = (uint32_t)~leaf_ref[l];
                    new_ref[l] = ~(int32_t)tri_order.size();
                    for (int32_t i = 0; i < leaf_count[l]; ++i) tri_order.push_back(t.tri_order[first + (uint32_t)i]);
                } else {
                    st.push_back(kid[1][(size_t)x]);
                    st.push_back(kid[0][(size_t)x]);
                }
            }
        }
        std::vector<FlatBvh::Node> nodes(order.size());
        for (size_t i = 0; i < order.size(); ++i) {
            const int32_t id = order[i];
            for (int k = 0; k < 2; ++k) {
                const int32_t c = kid[k][(size_t)id];
                for (int a = 0; a < 3; ++a) {
                    nodes[i].lo[k][a] = box[(size_t)c].lo[a];
                    nodes[i].hi[k][a] = box[(size_t)c].hi[a];
                }
                nodes[i].child[k] = is_leaf(c) ? new_ref[(size_t)(c - n_inner)] : new_index[(size_t)c];
                nodes[i].count[k] = is_leaf(c) ? leaf_count[(size_t)(c - n_inner)] : 0;
            }
        }
        t.nodes.swap(nodes);
        t.tri_order.swap(tri_order);
        t.depth = depth_of(root) + 1;  // the builder counts the level of the deepest leaf
    }
};

// returns the summed inner-node area after / before (1 = nothing gained)
float optimize_by_reinsertion(FlatBvh& t, int max_passes, uint32_t max_depth) {
    if (t.nodes.size() < 3) return 1.f;
    Reinserter r;
    r.load(t);
    const double before = r.inner_area();
    double prev = before;
    const int32_t total = (int32_t)r.box.size();
    std::vector<int32_t> cand;
    for (int pass = 0; pass < max_passes; ++pass) {
        cand.clear();
        for (int32_t id = 0; id < total; ++id)
            if (id != r.root) cand.push_back(id);
        // large nodes first: moving them changes the most area, and later (smaller) moves then see the improved upper tree
        std::stable_sort(cand.begin(), cand.end(), [&](int32_t a, int32_t b) { return r.box[(size_t)a].half_area() > r.box[(size_t)b].half_area(); });
        for (int32_t id : cand) r.reinsert(id);
        const double now = r.inner_area();
        if (now > prev * 0.998) {
            prev = now;
            break;
        }
        prev = now;
    }
    if (r.depth_of(r.root) + 1 > max_depth) return 1.f;  // deeper than the traversal stacks allow: keep the builder's tree
    r.store(t);
    return (float)(prev / before);
}

}  // namespace

FlatBvh build_bvh(const HostScene& scene, uint32_t max_leaf_size) {
    Builder b(scene, std::max(1u, max_leaf_size));
    b.run();
    int passes = 8;
    if (const char* e = std::getenv("RT_BVH_REINSERT_PASSES")) passes = std::atoi(e);  // developer override, 0 = the builder's tree as is
    if (passes > 0) {
        const float ratio = optimize_by_reinsertion(b.out, passes, 40u);
        if (std::getenv("RT_BVH_VERBOSE")) std::fprintf(stderr, "bvh: reinsertion left %.4f of the inner-node area, depth %u\n", ratio, b.out.depth);
    }
    return std::move(b.out);
}

}  // namespace rtb
