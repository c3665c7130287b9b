// accel_build.h — host-side acceleration structure builders.
#pragma once
#include <cstdint>
#include <vector>

#include "host_scene.h"

namespace rtb {

// The reference's loose octree (oct_tree_intersector.rs:66-146), flattened. Node i and cube i coincide, the 8
// children of an inner node are consecutive (first_child .. first_child+7) and numbered exactly as the
// reference numbers them (depth-first order of splitting), so an export can be compared index by index.
struct FlatOctree {
    std::vector<float> cubes;           // 6 per node: min xyz, max xyz
    std::vector<int32_t> first_child;   // -1 for a leaf
    std::vector<uint32_t> leaf_offset;  // nodes + 1 entries; leaf i references leaf_tris[leaf_offset[i] .. leaf_offset[i+1])
    std::vector<uint32_t> leaf_tris;    // global triangle indices, ascending inside a leaf
    uint32_t depth = 0;                 // deepest level that holds a node (root = 0)
    uint32_t max_stack = 1;             // upper bound of the traversal stack: 7 * depth + 1
    size_t num_nodes() const { return first_child.size(); }
};
FlatOctree build_octree(const HostScene& scene, uint32_t triangles_per_leaf);

// Binary SAH BVH over the triangle soup (the "flat GPU BVH" of the north star). Node layout is chosen for
// 16-byte loads; see DeviceBvhNode in device_types.h.
struct FlatBvh {
    struct Node {
        float lo[2][3], hi[2][3];  // child boxes (padded conservatively)
        int32_t child[2];          // >= 0: inner node index; < 0: leaf, ~child = first triangle slot
        int32_t count[2];          // triangles in the leaf (0 for inner)
    };
    std::vector<Node> nodes;         // nodes[0] = root (a scene with one leaf gets a root with an empty second child)
    std::vector<uint32_t> tri_order; // triangle slot -> global triangle index
    uint32_t depth = 0, max_leaf = 0, num_leaves = 0;
    float root_lo[3], root_hi[3];
};
FlatBvh build_bvh(const HostScene& scene, uint32_t max_leaf_size);

// 4-wide BVH with full-precision child boxes (bvh4_build.cpp): one node = one 128-byte line on the device.
struct FlatBvh4 {
    struct Node {
        float lo[4][3], hi[4][3];  // child boxes (the binary builder's padded boxes); empty slot: lo = hi = +inf
        int32_t child[4];          // >= 0: inner node index; < 0: leaf, ~child = first triangle slot
        int32_t count[4];          // triangles in the leaf (0 for inner / empty)
    };
    std::vector<Node> nodes;          // nodes[0] = root
    std::vector<uint32_t> tri_order;  // triangle slot -> global triangle index
    uint32_t depth = 0, max_leaf = 0, num_leaves = 0;
    float root_lo[3], root_hi[3];
};
FlatBvh4 build_bvh4(const HostScene& scene);

// Compressed 8-wide BVH (layout documented in cwbvh_build.cpp): 5 x 16-byte words per node.
struct CwWord {
    uint32_t x, y, z, w;
};
struct FlatCwbvh {
    std::vector<CwWord> nodes;        // 5 words per node, node 0 = root
    std::vector<uint32_t> tri_order;  // triangle slot -> global triangle index (a node's leaf triangles are contiguous)
    uint32_t depth = 0, num_leaves = 0;
    float root_lo[3], root_hi[3];
    size_t num_nodes() const { return nodes.size() / 5; }
};
FlatCwbvh build_cwbvh(const HostScene& scene);

}  // namespace rtb
