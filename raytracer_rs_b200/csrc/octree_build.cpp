// octree_build.cpp — host build of the reference's loose octree, emitted directly in flattened form.
//
// Semantics follow raytracer_lib/src/raytracer/accel_intersect/oct_tree_intersector.rs:
//   root cube = scene AABB (calc_extents :315-330); a leaf with more than `triangles_per_leaf` references and
//   recurse_level <= 8 is split into 8 half cubes around mid = 0.5*(max+min) (generate_child_cubes :274-313,
//   child bit0 -> x, bit1 -> y, bit2 -> z upper half); a child references every parent triangle that passes the
//   13-axis separating-axis test with inclusive comparisons (triangle_cube_intersection :393-458).
// The implementation is iterative (explicit work stack) but allocates child indices in the same order as the
// reference's recursion (split_node :94-146) so node numbers are identical.
#include <cfloat>
#include <cmath>

#include "accel_build.h"

namespace rtb {
namespace {

struct Box {
    f3 lo, hi;
};

inline void span_on_axis(const f3* pts, int n, f3 axis, float* lo, float* hi) {  // project_points_on_axis :460-469
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int i = 0; i < n; ++i) {
        const float d = dot3(axis, pts[i]);
        mn = std::fmin(mn, d);
        mx = std::fmax(mx, d);
    }
    *lo = mn;
    *hi = mx;
}

bool triangle_touches_box(const Box& b, const f3 tri[3]) {
    const f3 ax[3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
    const float blo[3] = {b.lo.x, b.lo.y, b.lo.z}, bhi[3] = {b.hi.x, b.hi.y, b.hi.z};
    float tlo, thi, clo, chi;
    for (int a = 0; a < 3; ++a) {  // box face normals
        span_on_axis(tri, 3, ax[a], &tlo, &thi);
        if (thi < blo[a] || tlo > bhi[a]) return false;
    }
    const f3 corner[8] = {b.lo,
                          {b.hi.x, b.lo.y, b.lo.z},
                          {b.lo.x, b.hi.y, b.lo.z},
                          {b.lo.x, b.lo.y, b.hi.z},
                          {b.lo.x, b.hi.y, b.hi.z},
                          {b.hi.x, b.lo.y, b.hi.z},
                          {b.hi.x, b.hi.y, b.lo.z},
                          b.hi};
    const f3 edge[3] = {tri[0] - tri[1], tri[1] - tri[2], tri[2] - tri[0]};
    const f3 nrm = cross3(edge[0], edge[1]);  // triangle plane
    const float plane = dot3(nrm, tri[0]);
    span_on_axis(corner, 8, nrm, &clo, &chi);
    if (chi < plane || clo > plane) return false;
    for (int e = 0; e < 3; ++e)  // 9 edge x axis directions, edge-major like the reference's `axes` array
        for (int a = 0; a < 3; ++a) {
            const f3 dir = cross3(edge[e], ax[a]);
            span_on_axis(corner, 8, dir, &clo, &chi);
            span_on_axis(tri, 3, dir, &tlo, &thi);
            if (chi < tlo || clo > thi) return false;
        }
    return true;
}

struct BuildNode {
    Box box;
    std::vector<uint32_t> tris;
    int32_t first_child = -1;
    uint32_t level = 0;
};

}  // namespace

FlatOctree build_octree(const HostScene& scene, uint32_t triangles_per_leaf) {
    const uint32_t ntri = scene.num_triangles();
    auto vertex = [&](uint32_t t, int c) { return f3{scene.vertices[9 * (size_t)t + 3 * c], scene.vertices[9 * (size_t)t + 3 * c + 1], scene.vertices[9 * (size_t)t + 3 * c + 2]}; };

    std::vector<BuildNode> nodes(1);
    nodes[0].box.lo = f3{FLT_MAX, FLT_MAX, FLT_MAX};
    nodes[0].box.hi = f3{-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t t = 0; t < ntri; ++t) {
        for (int c = 0; c < 3; ++c) {
            const f3 p = vertex(t, c);
            Box& b = nodes[0].box;
            b.lo = f3{std::fmin(b.lo.x, p.x), std::fmin(b.lo.y, p.y), std::fmin(b.lo.z, p.z)};
            b.hi = f3{std::fmax(b.hi.x, p.x), std::fmax(b.hi.y, p.y), std::fmax(b.hi.z, p.z)};
        }
        nodes[0].tris.push_back(t);
    }

    uint32_t depth = 0;
    std::vector<uint32_t> work{0};
    while (!work.empty()) {
        const uint32_t n = work.back();
        work.pop_back();
        if (nodes[n].tris.size() <= triangles_per_leaf || nodes[n].level > 8) continue;
        const Box pb = nodes[n].box;
        const f3 mid = 0.5f * (pb.hi + pb.lo);
        const uint32_t base = (uint32_t)nodes.size();
        const uint32_t child_level = nodes[n].level + 1;
        std::vector<uint32_t> parent_tris;
        parent_tris.swap(nodes[n].tris);
        nodes.resize(nodes.size() + 8);
        for (int c = 0; c < 8; ++c) {
            BuildNode& ch = nodes[base + c];
            ch.level = child_level;
            ch.box.lo = f3{(c & 1) ? mid.x : pb.lo.x, (c & 2) ? mid.y : pb.lo.y, (c & 4) ? mid.z : pb.lo.z};
            ch.box.hi = f3{(c & 1) ? pb.hi.x : mid.x, (c & 2) ? pb.hi.y : mid.y, (c & 4) ? pb.hi.z : mid.z};
            for (uint32_t t : parent_tris) {
                const f3 tri[3] = {vertex(t, 0), vertex(t, 1), vertex(t, 2)};
                if (triangle_touches_box(ch.box, tri)) ch.tris.push_back(t);
            }
        }
        nodes[n].first_child = (int32_t)base;
        if (child_level > depth) depth = child_level;
        for (int c = 7; c >= 0; --c) work.push_back(base + c);  // child 0 is split first, as in the recursion
    }

    FlatOctree out;
    out.depth = depth;
    out.max_stack = 7 * depth + 1;
    out.cubes.reserve(nodes.size() * 6);
    out.first_child.reserve(nodes.size());
    out.leaf_offset.reserve(nodes.size() + 1);
    for (const BuildNode& bn : nodes) {
        const float c[6] = {bn.box.lo.x, bn.box.lo.y, bn.box.lo.z, bn.box.hi.x, bn.box.hi.y, bn.box.hi.z};
        out.cubes.insert(out.cubes.end(), c, c + 6);
        out.first_child.push_back(bn.first_child);
        out.leaf_offset.push_back((uint32_t)out.leaf_tris.size());
        if (bn.first_child < 0) out.leaf_tris.insert(out.leaf_tris.end(), bn.tris.begin(), bn.tris.end());
    }
    out.leaf_offset.push_back((uint32_t)out.leaf_tris.size());
    return out;
}

}  // namespace rtb
