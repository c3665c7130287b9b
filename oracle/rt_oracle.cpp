// rt_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// This file is a plain C++ restatement of the per-pixel render loop of
// Andreas-Edling/raytracer-rs (reference paths below are relative to /root/reference).
// It is the correctness checker for the CUDA path and the CPU baseline of bench.py.
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
// load it; the product library (raytracer_rs_b200/librt_b200.so) never links or calls it.
//
// PARITY STATUS: the reference cannot be compiled here (no Rust toolchain, un-vendored
// crates), and its own tests hold golden vectors only for the ray/AABB slab test
// (oct_tree_intersector.rs:475-512), the Collada matrix conversion (collada_types.rs:98-125)
// and two matrix identities (vecmath.rs:343-359). Those vectors are checked in
// tests/test_oracle_golden.py. Everything else on the path (Moller-Trumbore, octree
// build/traversal, shading, tonemap, packing, pixel->ray mapping) is "parity unpinned":
// no reference vector or reference run exists for it; this restatement follows the source
// line by line. It is cross-checked bit for bit against a second restatement written
// independently from the Rust sources in numpy float32 (tests/render_ref.py,
// tests/test_oracle_golden.py::test_oracle_matches_independent_numpy_restatement), which guards
// against transcription slips but is not a run of the reference: the status stays "unpinned".
//
// Arithmetic rules: IEEE binary32 everywhere, no FMA contraction (-ffp-contract=off), sums
// left to right exactly as the Rust source writes them, glibc tanf/sinf/cosf/powf/sqrtf (what
// Rust's f32::{tan,sin,cos,powf,sqrt} lower to on x86_64-unknown-linux-gnu).
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC)

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cfloat>
#include <vector>
#include <algorithm>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// ---------------------------------------------------------------------------------------
// vecmath  (raytracer_lib/src/vecmath.rs)
// ---------------------------------------------------------------------------------------
struct Vec3 {
    float x, y, z;
};
struct Vec4 {
    float x, y, z, w;
};
struct Ray {
    Vec3 pos, dir;
};

static inline Vec3 v3(float x, float y, float z) { return Vec3{x, y, z}; }
static inline Vec3 add(const Vec3& a, const Vec3& b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }   // vecmath.rs:35-38
static inline Vec3 sub(const Vec3& a, const Vec3& b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }   // vecmath.rs:40-43
static inline Vec3 mul(const Vec3& a, float s) { return v3(a.x * s, a.y * s, a.z * s); }               // vecmath.rs:45-46
static inline Vec3 mul(float s, const Vec3& a) { return v3(s * a.x, s * a.y, s * a.z); }               // vecmath.rs:47-48
static inline float dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }    // vecmath.rs:74-76
static inline Vec3 cross(const Vec3& a, const Vec3& b) {                                               // vecmath.rs:79-85
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline Vec3 normalized(const Vec3& v) {                                                         // vecmath.rs:23-26
    float len = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    return v3(v.x / len, v.y / len, v.z / len);
}

struct Matrix {
    float e[16];
};
static Matrix mat_ident() {  // vecmath.rs:107-114
    Matrix m;
    for (int i = 0; i < 16; ++i) m.e[i] = 0.0f;
    m.e[0] = m.e[5] = m.e[10] = m.e[15] = 1.0f;
    return m;
}
static Matrix mat_rot_x(float r) {  // vecmath.rs:116-123
    Matrix m = mat_ident();
    m.e[5] = cosf(r);
    m.e[6] = -sinf(r);
    m.e[9] = sinf(r);
    m.e[10] = cosf(r);
    return m;
}
static Matrix mat_rot_y(float r) {  // vecmath.rs:124-131
    Matrix m = mat_ident();
    m.e[0] = cosf(r);
    m.e[2] = sinf(r);
    m.e[8] = -sinf(r);
    m.e[10] = cosf(r);
    return m;
}
static Matrix mat_translate(const Vec3& v) {  // vecmath.rs:133-139
    Matrix m = mat_ident();
    m.e[12] = v.x;
    m.e[13] = v.y;
    m.e[14] = v.z;
    return m;
}
static Matrix mat_transpose(const Matrix& s) {  // vecmath.rs:141-159
    Matrix m = s;
    m.e[1] = s.e[4];
    m.e[2] = s.e[8];
    m.e[3] = s.e[12];
    m.e[4] = s.e[1];
    m.e[6] = s.e[9];
    m.e[7] = s.e[13];
    m.e[8] = s.e[2];
    m.e[9] = s.e[6];
    m.e[11] = s.e[14];
    m.e[12] = s.e[3];
    m.e[13] = s.e[7];
    m.e[14] = s.e[11];
    return m;
}
// row-vector * matrix, vecmath.rs:200-211
static inline Vec4 mat_mul_vec4(const Matrix& m, const Vec4& v) {
    Vec4 r;
    r.x = v.x * m.e[0] + v.y * m.e[4] + v.z * m.e[8] + v.w * m.e[12];
    r.y = v.x * m.e[1] + v.y * m.e[5] + v.z * m.e[9] + v.w * m.e[13];
    r.z = v.x * m.e[2] + v.y * m.e[6] + v.z * m.e[10] + v.w * m.e[14];
    r.w = v.x * m.e[3] + v.y * m.e[7] + v.z * m.e[11] + v.w * m.e[15];
    return r;
}
// 4x4 product, vecmath.rs:237-313 (row i of lhs times column j of rhs, 4-term sum left to right)
static Matrix mat_mul(const Matrix& a, const Matrix& b) {
    Matrix r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.e[4 * i + j] = a.e[4 * i + 0] * b.e[0 + j] + a.e[4 * i + 1] * b.e[4 + j] +
                             a.e[4 * i + 2] * b.e[8 + j] + a.e[4 * i + 3] * b.e[12 + j];
    return r;
}

// collada_types.rs:76-90 : reflect_z * transpose(collada) * swap_yz
static Matrix collada_to_vecmath(const float elems[16]) {
    Matrix cm;
    memcpy(cm.e, elems, sizeof(cm.e));
    Matrix row_major = mat_transpose(cm);
    Matrix swap_yx = {{1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1}};
    Matrix reflect_z = {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1}};
    return mat_mul(mat_mul(reflect_z, row_major), swap_yx);
}

// ---------------------------------------------------------------------------------------
// camera  (raytracer_lib/src/scene/camera.rs)
// ---------------------------------------------------------------------------------------
struct Camera {
    float x_angle = 0.0f, y_angle = 0.0f;
    Vec3 pos{0, 0, 0};
    size_t width = 0, height = 0;
    Matrix base_orientation, base_rotation, orientation, rotation;
    float max_x = 0, max_y = 0;

    void update_matrices() {  // camera.rs:92-98
        rotation = mat_mul(mat_mul(mat_rot_x(x_angle), mat_rot_y(y_angle)), base_rotation);
        orientation = mat_mul(mat_mul(rotation, mat_translate(pos)), base_orientation);
    }
    void init(size_t w, size_t h, const Matrix& orient, float fov_deg) {  // camera.rs:22-61
        Matrix rot = orient;
        rot.e[3] = 0.0f;
        rot.e[7] = 0.0f;
        rot.e[11] = 0.0f;
        rot.e[12] = 0.0f;
        rot.e[13] = 0.0f;
        rot.e[14] = 0.0f;
        rot.e[15] = 1.0f;
        float fov = fov_deg * 3.14159265358979323846f / 180.0f;  // std::f32::consts::PI
        float half_fov = 0.5f * fov;
        max_x = 1.0f * tanf(half_fov);
        max_y = 1.0f * tanf(half_fov);
        x_angle = y_angle = 0.0f;
        pos = v3(0, 0, 0);
        width = w;
        height = h;
        base_orientation = orient;
        base_rotation = rot;
        update_matrices();
    }
    // camera.rs:80-90, with the two random numbers passed in
    Ray get_ray(size_t u, size_t v, float xi1, float xi2) const {
        float dir_x = -max_x + 2.0f * max_x * (((float)u + xi1) / (float)width);
        float dir_y = -max_y + 2.0f * max_y * (((float)v + xi2) / (float)height);
        Vec4 d = mat_mul_vec4(rotation, Vec4{dir_x, -dir_y, 1.0f, 1.0f});
        Vec4 p = mat_mul_vec4(orientation, Vec4{0.0f, 0.0f, 0.0f, 1.0f});
        return Ray{v3(p.x, p.y, p.z), v3(d.x, d.y, d.z)};
    }
};

// ---------------------------------------------------------------------------------------
// scene  (raytracer_lib/src/scene/{mod,color,texture}.rs)
// ---------------------------------------------------------------------------------------
struct RGB {
    float r, g, b;
};
struct Texture {
    size_t width, height;
    std::vector<RGB> data;
    // texture.rs:21-27 ; the reference panics on an out-of-range index (v == 1.0), the oracle
    // clamps to the last texel instead (same rule in the CUDA path).
    const RGB& get_texel(float u, float v) const {
        float fx = u * (float)width, fy = v * (float)height;
        size_t x = (fx > 0.0f) ? (size_t)fx : 0;  // Rust `as usize`: saturating, NaN -> 0
        size_t y = (fy > 0.0f) ? (size_t)fy : 0;
        size_t idx = y * width + x;
        if (idx >= data.size()) idx = data.size() - 1;
        return data[idx];
    }
};
struct Material {
    int kind;  // 0 = Diffuse::Color, 1 = Diffuse::TextureId
    RGB color;
    size_t texture_id;
};
struct Geometry {
    std::vector<Vec3> vertices;  // == transformed_vertices (scene/mod.rs:53-55)
    Material material;
    uint32_t first_triangle;  // global triangle index of triangle 0
};
struct Light {
    Vec3 pos;
    RGB color;
};
struct Scene {
    std::vector<Geometry> geometries;
    std::vector<Light> lights;
    std::vector<Texture> textures;
};

struct HitInfo {
    float t, u, v;
};
struct Hit {
    HitInfo hit_info;
    size_t geometry_index, vertex_index;
};

// per-ray-kind work counters (exact replacements for the survey's sampled figures)
struct Counters {
    uint64_t rays[3] = {0, 0, 0};        // 0 primary, 1 shadow, 2 bounce
    uint64_t cube_tests[3] = {0, 0, 0};  // intersect_cube_inverse_ray calls
    uint64_t tri_tests[3] = {0, 0, 0};   // Moller-Trumbore calls
    uint64_t inner_nodes[3] = {0, 0, 0};
    uint64_t leaves[3] = {0, 0, 0};
    uint64_t leaf_rejects[3] = {0, 0, 0};  // hits discarded by the in-cube test (Q6)
    uint64_t primary_hits = 0;
    uint64_t shadow_blocked = 0;
    void add(const Counters& o) {
        for (int k = 0; k < 3; ++k) {
            rays[k] += o.rays[k];
            cube_tests[k] += o.cube_tests[k];
            tri_tests[k] += o.tri_tests[k];
            inner_nodes[k] += o.inner_nodes[k];
            leaves[k] += o.leaves[k];
            leaf_rejects[k] += o.leaf_rejects[k];
        }
        primary_hits += o.primary_hits;
        shadow_blocked += o.shadow_blocked;
    }
};

// ---------------------------------------------------------------------------------------
// Moller-Trumbore  (raytracer_lib/src/raytracer/intersect.rs:62-98, intersect_late_out)
// ---------------------------------------------------------------------------------------
static inline bool moller_trumbore(const Ray& ray, const Vec3& v0, const Vec3& v1, const Vec3& v2, HitInfo* out) {
    Vec3 v0v1 = sub(v1, v0);
    Vec3 v0v2 = sub(v2, v0);
    Vec3 pvec = cross(ray.dir, v0v2);
    float det = dot(v0v1, pvec);
    if (fabsf(det) < FLT_EPSILON) return false;  // std::f32::EPSILON = 1.1920929e-7
    float inv_det = 1.0f / det;
    Vec3 tvec = sub(ray.pos, v0);
    float u = dot(tvec, pvec) * inv_det;
    Vec3 qvec = cross(tvec, v0v1);
    float v = dot(ray.dir, qvec) * inv_det;
    float t = dot(v0v2, qvec) * inv_det;
    if (u < 0.0f || u > 1.0f) return false;
    if (v < 0.0f || u + v > 1.0f) return false;
    if (t < 0.0f) return false;
    out->t = t;
    out->u = u;
    out->v = v;
    return true;
}

// ---------------------------------------------------------------------------------------
// octree  (raytracer_lib/src/raytracer/accel_intersect/oct_tree_intersector.rs)
// ---------------------------------------------------------------------------------------
struct Cube {
    Vec3 min, max;
    bool contains(const Vec3& v) const {  // :34-45
        if (v.x < min.x || v.x > max.x || v.y < min.y || v.y > max.y || v.z < min.z || v.z > max.z) return false;
        return true;
    }
};
struct TriangleIndex {
    size_t geom_idx, tri_idx;  // tri_idx = index of the triangle's first vertex (3 * triangle number)
};
struct OctNode {
    bool is_leaf;
    size_t cube_index;                            // Leaf
    std::vector<TriangleIndex> triangle_indices;  // Leaf
    size_t children[8];                           // Node
};

// :348-372 ; Rust f32::min/max ignore NaN exactly like fminf/fmaxf
static inline bool intersect_cube_inverse_ray(const Ray& inv_ray, const Cube& cube, float* t_out) {
    float tx1 = (cube.min.x - inv_ray.pos.x) * inv_ray.dir.x;
    float tx2 = (cube.max.x - inv_ray.pos.x) * inv_ray.dir.x;
    float tmin = fminf(tx1, tx2);
    float tmax = fmaxf(tx1, tx2);
    float ty1 = (cube.min.y - inv_ray.pos.y) * inv_ray.dir.y;
    float ty2 = (cube.max.y - inv_ray.pos.y) * inv_ray.dir.y;
    tmin = fmaxf(tmin, fminf(ty1, ty2));
    tmax = fminf(tmax, fmaxf(ty1, ty2));
    float tz1 = (cube.min.z - inv_ray.pos.z) * inv_ray.dir.z;
    float tz2 = (cube.max.z - inv_ray.pos.z) * inv_ray.dir.z;
    tmin = fmaxf(tmin, fminf(tz1, tz2));
    tmax = fminf(tmax, fmaxf(tz1, tz2));
    if (tmax >= tmin && tmax > 0.0f) {
        *t_out = tmin;
        return true;
    }
    return false;
}

static void project_points_on_axis(const Vec3* pts, size_t n, const Vec3& axis, float* mn, float* mx) {  // :460-469
    float lo = FLT_MAX, hi = -FLT_MAX;  // std::f32::MAX / std::f32::MIN
    for (size_t i = 0; i < n; ++i) {
        float val = dot(axis, pts[i]);
        lo = fminf(lo, val);
        hi = fmaxf(hi, val);
    }
    *mn = lo;
    *mx = hi;
}

static bool triangle_cube_intersection(const Cube& cube, const Vec3* tri) {  // :393-458
    Vec3 x_axis = v3(1, 0, 0), y_axis = v3(0, 1, 0), z_axis = v3(0, 0, 1);
    float tmin, tmax;
    project_points_on_axis(tri, 3, x_axis, &tmin, &tmax);
    if (tmax < cube.min.x || tmin > cube.max.x) return false;
    project_points_on_axis(tri, 3, y_axis, &tmin, &tmax);
    if (tmax < cube.min.y || tmin > cube.max.y) return false;
    project_points_on_axis(tri, 3, z_axis, &tmin, &tmax);
    if (tmax < cube.min.z || tmin > cube.max.z) return false;

    Vec3 cv[8] = {cube.min,
                  v3(cube.max.x, cube.min.y, cube.min.z),
                  v3(cube.min.x, cube.max.y, cube.min.z),
                  v3(cube.min.x, cube.min.y, cube.max.z),
                  v3(cube.min.x, cube.max.y, cube.max.z),
                  v3(cube.max.x, cube.min.y, cube.max.z),
                  v3(cube.max.x, cube.max.y, cube.min.z),
                  cube.max};
    Vec3 e1 = sub(tri[0], tri[1]);
    Vec3 e2 = sub(tri[1], tri[2]);
    Vec3 n = cross(e1, e2);
    float tri_offset = dot(n, tri[0]);
    float cmin, cmax;
    project_points_on_axis(cv, 8, n, &cmin, &cmax);
    if (cmax < tri_offset || cmin > tri_offset) return false;

    Vec3 e3 = sub(tri[2], tri[0]);
    Vec3 axes[9] = {cross(e1, x_axis), cross(e1, y_axis), cross(e1, z_axis), cross(e2, x_axis), cross(e2, y_axis),
                    cross(e2, z_axis), cross(e3, x_axis), cross(e3, y_axis), cross(e3, z_axis)};
    for (int a = 0; a < 9; ++a) {
        project_points_on_axis(cv, 8, axes[a], &cmin, &cmax);
        project_points_on_axis(tri, 3, axes[a], &tmin, &tmax);
        if (cmax < tmin || cmin > tmax) return false;
    }
    return true;
}

static void generate_child_cubes(const Cube& c, Cube out[8]) {  // :274-313
    Vec3 mid = mul(0.5f, add(c.max, c.min));
    Vec3 mn = c.min, mx = c.max;
    out[0] = Cube{v3(mn.x, mn.y, mn.z), v3(mid.x, mid.y, mid.z)};
    out[1] = Cube{v3(mid.x, mn.y, mn.z), v3(mx.x, mid.y, mid.z)};
    out[2] = Cube{v3(mn.x, mid.y, mn.z), v3(mid.x, mx.y, mid.z)};
    out[3] = Cube{v3(mid.x, mid.y, mn.z), v3(mx.x, mx.y, mid.z)};
    out[4] = Cube{v3(mn.x, mn.y, mid.z), v3(mid.x, mid.y, mx.z)};
    out[5] = Cube{v3(mid.x, mn.y, mid.z), v3(mx.x, mid.y, mx.z)};
    out[6] = Cube{v3(mn.x, mid.y, mid.z), v3(mid.x, mx.y, mx.z)};
    out[7] = Cube{v3(mid.x, mid.y, mid.z), v3(mx.x, mx.y, mx.z)};
}

struct OctTree {
    std::vector<Cube> cubes;
    std::vector<OctNode> nodes;
    size_t trunk = 0;
    size_t max_level = 0;

    void build(const Scene& scene, size_t triangles_per_leaf) {  // :66-81
        Cube trunk_cube;                                           // calc_extents :315-330
        trunk_cube.min = v3(FLT_MAX, FLT_MAX, FLT_MAX);
        trunk_cube.max = v3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
        for (const Geometry& g : scene.geometries)
            for (const Vec3& p : g.vertices) {
                trunk_cube.min.x = fminf(trunk_cube.min.x, p.x);
                trunk_cube.min.y = fminf(trunk_cube.min.y, p.y);
                trunk_cube.min.z = fminf(trunk_cube.min.z, p.z);
                trunk_cube.max.x = fmaxf(trunk_cube.max.x, p.x);
                trunk_cube.max.y = fmaxf(trunk_cube.max.y, p.y);
                trunk_cube.max.z = fmaxf(trunk_cube.max.z, p.z);
            }
        OctNode root;
        root.is_leaf = true;
        root.cube_index = 0;
        for (size_t g = 0; g < scene.geometries.size(); ++g)  // all_triangle_indices :332-342
            for (size_t t = 0; t < scene.geometries[g].vertices.size() / 3; ++t)
                root.triangle_indices.push_back(TriangleIndex{g, t * 3});
        cubes.clear();
        nodes.clear();
        cubes.push_back(trunk_cube);
        nodes.push_back(root);
        trunk = 0;
        max_level = 0;
        split_node(trunk, triangles_per_leaf, scene, 0);
    }

    void split_node(size_t node_idx, size_t num_triangles, const Scene& scene, size_t level) {  // :94-146
        if (!nodes[node_idx].is_leaf) return;
        if (nodes[node_idx].triangle_indices.size() <= num_triangles || level > 8) return;
        std::vector<OctNode> new_nodes;
        size_t new_child_list[8];
        {
            const OctNode& leaf = nodes[node_idx];
            Cube child_cubes[8];
            generate_child_cubes(cubes[leaf.cube_index], child_cubes);
            for (int i = 0; i < 8; ++i) {
                OctNode child;
                child.is_leaf = true;
                for (const TriangleIndex& ti : leaf.triangle_indices) {  // triangles_intersecting_cube :374-391
                    if (triangle_cube_intersection(child_cubes[i], &scene.geometries[ti.geom_idx].vertices[ti.tri_idx]))
                        child.triangle_indices.push_back(ti);
                }
                cubes.push_back(child_cubes[i]);
                size_t child_idx = cubes.size() - 1;
                child.cube_index = child_idx;
                new_nodes.push_back(std::move(child));
                new_child_list[i] = child_idx;
            }
        }
        OctNode inner;
        inner.is_leaf = false;
        inner.cube_index = 0;
        memcpy(inner.children, new_child_list, sizeof(new_child_list));
        nodes[node_idx] = std::move(inner);
        size_t start = nodes.size();
        for (auto& n : new_nodes) nodes.push_back(std::move(n));
        if (level + 1 > max_level) max_level = level + 1;
        for (size_t c = start; c < start + 8; ++c) split_node(c, num_triangles, scene, level + 1);
    }

    // intersect_leaf_triangles :249-272
    bool intersect_leaf(const Scene& scene, const Ray& ray, const OctNode& leaf, Hit* out, Counters& cnt, int kind) const {
        bool have = false;
        Hit best{};
        for (const TriangleIndex& ti : leaf.triangle_indices) {
            const Vec3* tv = &scene.geometries[ti.geom_idx].vertices[ti.tri_idx];
            HitInfo hi;
            cnt.tri_tests[kind]++;
            if (!moller_trumbore(ray, tv[0], tv[1], tv[2], &hi)) continue;
            if (!have || hi.t < best.hit_info.t) {
                have = true;
                best = Hit{hi, ti.geom_idx, ti.tri_idx};
            }
        }
        if (have) *out = best;
        return have;
    }

    // intersect_node :148-196
    bool intersect_node(const Scene& scene, const Ray& ray, const Ray& inv_ray, size_t node_idx, Hit* out, Counters& cnt,
                        int kind) const {
        const OctNode& node = nodes[node_idx];
        if (node.is_leaf) {
            cnt.leaves[kind]++;
            Hit h;
            if (!intersect_leaf(scene, ray, node, &h, cnt, kind)) return false;
            Vec3 hit_point = add(ray.pos, mul(ray.dir, h.hit_info.t));
            if (cubes[node_idx].contains(hit_point)) {
                *out = h;
                return true;
            }
            cnt.leaf_rejects[kind]++;
            return false;
        }
        cnt.inner_nodes[kind]++;
        std::pair<size_t, float> dist[8];
        int n = 0;
        for (int c = 0; c < 8; ++c) {
            float t;
            cnt.cube_tests[kind]++;
            if (intersect_cube_inverse_ray(inv_ray, cubes[node.children[c]], &t)) dist[n++] = {node.children[c], t};
        }
        // slice::sort_by is a stable sort; partial_cmp().unwrap() cannot see NaN here because a
        // NaN tmin fails `tmax >= tmin`
        std::stable_sort(dist, dist + n, [](const std::pair<size_t, float>& a, const std::pair<size_t, float>& b) {
            return a.second < b.second;
        });
        for (int i = 0; i < n; ++i)
            if (intersect_node(scene, ray, inv_ray, dist[i].first, out, cnt, kind)) return true;
        return false;
    }

    bool intersect_ray(const Scene& scene, const Ray& ray, Hit* out, Counters& cnt, int kind) const {  // :240-246
        Ray inv_ray{ray.pos, v3(1.0f / ray.dir.x, 1.0f / ray.dir.y, 1.0f / ray.dir.z)};
        cnt.rays[kind]++;
        return intersect_node(scene, ray, inv_ray, trunk, out, cnt, kind);
    }
};

// no_acceleration_intersector.rs:13-41
static bool brute_force_intersect(const Scene& scene, const Ray& ray, Hit* out, Counters& cnt, int kind) {
    bool have = false;
    Hit best{};
    cnt.rays[kind]++;
    for (size_t g = 0; g < scene.geometries.size(); ++g) {
        const std::vector<Vec3>& vs = scene.geometries[g].vertices;
        for (size_t t = 0; t + 2 < vs.size(); t += 3) {
            HitInfo hi;
            cnt.tri_tests[kind]++;
            if (!moller_trumbore(ray, vs[t], vs[t + 1], vs[t + 2], &hi)) continue;
            if (!have || hi.t < best.hit_info.t) {
                have = true;
                best = Hit{hi, g, t};
            }
        }
    }
    if (have) *out = best;
    return have;
}

// ---------------------------------------------------------------------------------------
// deterministic stand-ins for the reference's OS-seeded RNGs (mod.rs:84,152; camera.rs:82,84;
// sample_generator.rs:32,37-44). The reference is not reproducible; oracle and
// CUDA path share this counter-based generator instead. Spec (also in DESIGN.md):
//   mix(h): h ^= h>>16; h *= 0x7feb352d; h ^= h>>15; h *= 0x846ca68b; h ^= h>>16
//   hash(a,b,c,d) = mix(mix(mix(mix(a) ^ b) ^ c) ^ d)  (with +golden-ratio offsets, below)
//   uniform [0,1) = (hash >> 8) * 2^-24
// ---------------------------------------------------------------------------------------
static inline uint32_t mix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x7feb352dU;
    h ^= h >> 15;
    h *= 0x846ca68bU;
    h ^= h >> 16;
    return h;
}
static inline uint32_t hash4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t h = mix32(a + 0x9e3779b9U);
    h = mix32(h ^ (b + 0x85ebca6bU));
    h = mix32(h ^ (c + 0xc2b2ae35U));
    h = mix32(h ^ (d + 0x27d4eb2fU));
    return h;
}
static inline float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

enum JitterMode { JITTER_FIXED_HALF = 0, JITTER_HASHED = 1 };
enum IntersectorKind { ISECT_OCTREE = 0, ISECT_BRUTE = 1 };

// sample_generator.rs:9-53 with the table drawn from hash4(seed ^ 0x5a17ab1e, i, attempt, axis)
struct SampleGenerator {
    std::vector<Vec3> normalized_vecs;
    void init(uint32_t seed) {
        normalized_vecs.resize(65536);
        for (uint32_t i = 0; i < 65536; ++i) {
            for (uint32_t attempt = 0;; ++attempt) {
                Vec3 d = v3(u01(hash4(seed ^ 0x5a17ab1eU, i, attempt, 0)) * 2.0f - 1.0f,
                            u01(hash4(seed ^ 0x5a17ab1eU, i, attempt, 1)) * 2.0f - 1.0f,
                            u01(hash4(seed ^ 0x5a17ab1eU, i, attempt, 2)) * 2.0f - 1.0f);
                if (dot(d, d) < 1.0f && dot(d, d) > 0.0f) {
                    normalized_vecs[i] = normalized(d);
                    break;
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------
// film / tonemap / pack  (film.rs:3-48, tonemap.rs:4-10, color.rs:85-95)
// ---------------------------------------------------------------------------------------
struct PixelData {
    RGB sum{0, 0, 0};
    RGB sum_sq{0, 0, 0};
    uint32_t n = 0;
};
static inline uint32_t to_u8(float x) {
    float c = fmaxf(fminf(x, 1.0f), 0.0f) * 255.0f;  // Rust min/max drop NaN -> NaN becomes 1.0
    return (uint32_t)(uint8_t)c;                     // `as u8`: truncation (value already in [0,255])
}
static inline uint32_t tonemap_pack(const RGB& mean) {
    RGB m{mean.r / (1.0f + mean.r), mean.g / (1.0f + mean.g), mean.b / (1.0f + mean.b)};  // simple_map
    uint32_t r = to_u8(m.r), g = to_u8(m.g), b = to_u8(m.b), a = to_u8(1.0f);
    return b | (g << 8) | (r << 16) | (a << 24);
}

// ---------------------------------------------------------------------------------------
// the render loop  (raytracer_lib/src/raytracer/mod.rs)
// ---------------------------------------------------------------------------------------
struct RayTracer {
    size_t width = 0, height = 0;
    Camera camera;
    Scene scene;
    OctTree octree;
    SampleGenerator sample_generator;
    std::vector<PixelData> film;
    std::vector<uint32_t> primary_ids;
    size_t current_row = 0;
    // pinned-mode switches (reference: RECURSIONS=2, SUB_SPREAD=1, OS-seeded jitter; mod.rs:81-84)
    int recursions = 2;
    uint32_t sub_spread = 1;
    int jitter_mode = JITTER_HASHED;
    uint32_t seed = 0;
    int intersector = ISECT_OCTREE;
    Counters counters;

    bool intersect(const Ray& ray, Hit* out, Counters& cnt, int kind) const {
        if (intersector == ISECT_BRUTE) return brute_force_intersect(scene, ray, out, cnt, kind);
        return octree.intersect_ray(scene, ray, out, cnt, kind);
    }

    Vec3 calc_normal(const Hit& hit) const {  // mod.rs:198-205
        const std::vector<Vec3>& gv = scene.geometries[hit.geometry_index].vertices;
        Vec3 n = cross(sub(gv[hit.vertex_index + 1], gv[hit.vertex_index]), sub(gv[hit.vertex_index + 2], gv[hit.vertex_index]));
        return normalized(n);
    }

    RGB shade(const Ray& ray, const Hit& hit, const Vec3& normal, Counters& cnt) const {  // mod.rs:207-261
        RGB accum{0, 0, 0};
        Vec3 hit_point = add(ray.pos, mul(hit.hit_info.t, ray.dir));
        for (const Light& light : scene.lights) {
            Ray ray_to_light{hit_point, sub(light.pos, hit_point)};
            Vec3 ldir_n = normalized(ray_to_light.dir);
            float dot_light_normal = dot(normal, ldir_n);
            if (dot_light_normal < 0.0f) continue;
            bool blocked = false;
            Ray offs{add(ray_to_light.pos, mul(ray_to_light.dir, 0.01f)), ray_to_light.dir};
            Hit sh;
            if (intersect(offs, &sh, cnt, 1)) {
                if (sh.hit_info.t > 0.01f && sh.hit_info.t < 1.0f) blocked = true;
            }
            if (blocked) {
                cnt.shadow_blocked++;
                continue;
            }
            const Material& mat = scene.geometries[hit.geometry_index].material;
            RGB diffuse = mat.kind == 0 ? mat.color : scene.textures[mat.texture_id].get_texel(hit.hit_info.u, hit.hit_info.v);
            Vec3 view_ray = normalized(ray.dir);
            // 2.0 * dot_light_normal * normal - ray_to_light.dir.normalized()
            Vec3 reflected = sub(mul(2.0f * dot_light_normal, normal), normalized(ray_to_light.dir));
            float spec = powf(dot(view_ray, reflected), 32.0f);
            // (diffuse_rgb * ndl + SPECULAR * spec) * light.color ; SPECULAR = white
            RGB term{(diffuse.r * dot_light_normal + 1.0f * spec) * light.color.r,
                     (diffuse.g * dot_light_normal + 1.0f * spec) * light.color.g,
                     (diffuse.b * dot_light_normal + 1.0f * spec) * light.color.b};
            accum.r += term.r;
            accum.g += term.g;
            accum.b += term.b;
        }
        return accum;
    }

    // mod.rs:178-196 ; `path` identifies the bounce ray inside the sample's recursion tree so
    // that its random numbers do not depend on evaluation order
    Ray randomize_reflection_ray(const Hit& hit, const Ray& ray, const Vec3& normal, uint32_t pixel, uint32_t sample,
                                 uint32_t path) const {
        // normalized_vec_pseudo: sample_idx = random_range(0..NUM_SAMPLES-1)  -> 0..65534
        uint32_t idx = (uint32_t)(((uint64_t)hash4(seed ^ 0xb0c0ffeeU, pixel, sample, path) * 65535ULL) >> 32);
        Vec3 d = sample_generator.normalized_vecs[idx];
        while (dot(d, normal) <= 0.0f) {
            idx = (idx + 1) % 65535;  // normalized_vec_lookup: (idx + 1) % SAMPLE_MAX
            d = sample_generator.normalized_vecs[idx];
        }
        Vec3 hit_point = add(ray.pos, mul(hit.hit_info.t, ray.dir));
        hit_point = add(hit_point, mul(0.00001f, d));
        return Ray{hit_point, d};
    }

    RGB compute_radiance(const Ray& ray, const Hit& hit, int rec, uint32_t spread, uint32_t pixel, uint32_t sample, uint32_t path,
                         Counters& cnt) const {  // mod.rs:132-176
        Vec3 normal = calc_normal(hit);
        RGB radiance = shade(ray, hit, normal, cnt);
        if (rec < 1) return radiance;
        uint32_t num_sub_rays = spread * (uint32_t)rec;
        RGB sum{0, 0, 0};
        for (uint32_t k = 0; k < num_sub_rays; ++k) {
            uint32_t sub_path = path * 31u + k + 1u;
            Ray sub_ray = randomize_reflection_ray(hit, ray, normal, pixel, sample, sub_path);
            Hit sub_hit;
            RGB x{0, 0, 0};
            if (intersect(sub_ray, &sub_hit, cnt, 2)) x = compute_radiance(sub_ray, sub_hit, rec - 1, spread, pixel, sample, sub_path, cnt);
            sum.r = sum.r + x.r;
            sum.g = sum.g + x.g;
            sum.b = sum.b + x.b;
        }
        float inv = 1.0f / (float)num_sub_rays;
        return RGB{radiance.r + sum.r * inv, radiance.g + sum.g * inv, radiance.b + sum.b * inv};
    }

    void trace_pixel(size_t row, size_t i, Counters& cnt) {  // body of mod.rs:88-112
        size_t idx = row * width + i;
        PixelData& pd = film[idx];
        float xi1 = 0.5f, xi2 = 0.5f;
        if (jitter_mode == JITTER_HASHED) {
            xi1 = u01(hash4(seed, (uint32_t)idx, pd.n, 0));
            xi2 = u01(hash4(seed, (uint32_t)idx, pd.n, 1));
        }
        Ray ray = camera.get_ray(idx % width, idx / height, xi1, xi2);  // Q1: idx / height
        Hit hit;
        RGB color{0, 0, 0};
        if (intersect(ray, &hit, cnt, 0)) {
            cnt.primary_hits++;
            primary_ids[idx] = scene.geometries[hit.geometry_index].first_triangle + (uint32_t)(hit.vertex_index / 3);
            color = compute_radiance(ray, hit, recursions, sub_spread, (uint32_t)idx, pd.n, 0u, cnt);
        } else {
            primary_ids[idx] = 0xFFFFFFFFu;
        }
        pd.sum.r += color.r;  // add_sample, film.rs:20-24
        pd.sum.g += color.g;
        pd.sum.b += color.b;
        pd.sum_sq.r += color.r * color.r;
        pd.sum_sq.g += color.g * color.g;
        pd.sum_sq.b += color.b * color.b;
        pd.n += 1;
    }

    // `spp` passes over rows [first_row, first_row + n_rows) (wrapping modulo height)
    void trace_rows(size_t first_row, size_t n_rows, size_t spp, int threads) {
        Counters total;
#ifdef _OPENMP
        if (threads < 1) threads = omp_get_max_threads();
#else
        threads = 1;
#endif
        for (size_t s = 0; s < spp; ++s) {
#pragma omp parallel num_threads(threads)
            {
                Counters local;
#pragma omp for schedule(dynamic, 1)
                for (long r = 0; r < (long)n_rows; ++r) {
                    size_t row = (first_row + (size_t)r) % height;
                    for (size_t i = 0; i < width; ++i) trace_pixel(row, i, local);
                }
#pragma omp critical
                total.add(local);
            }
        }
        counters.add(total);
    }

    uint32_t trace_frame_additive(int threads) {  // mod.rs:80-117 (50 rows per call)
        // rows may wrap; process them one by one in order so duplicates (height < 50) stay sequential
        if (height >= 50) {
            trace_rows(current_row, 50, 1, threads);
        } else {
            for (int k = 0; k < 50; ++k) trace_rows((current_row + k) % height, 1, 1, threads);
        }
        current_row = (current_row + 50) % height;
        return (uint32_t)(50 * width);
    }

    void get_tonemapped_pixels(uint32_t* out) const {  // mod.rs:120-128
        for (size_t i = 0; i < film.size(); ++i) {
            const PixelData& pd = film[i];
            float inv = 1.0f / (float)pd.n;  // film.rs:43-48
            RGB mean{pd.sum.r * inv, pd.sum.g * inv, pd.sum.b * inv};
            out[i] = tonemap_pack(mean);
        }
    }
    void film_clear() {  // film.rs:37-41
        for (PixelData& p : film) p = PixelData();
    }
    // Film::get_estimated_variances, film.rs:50-67 (dead code in the reference, restated for completeness of the film API).
    // `num_samples * (num_samples - 1)` is u32 arithmetic: the release build the reference ships (ci.yaml:35-39) wraps, so
    // n = 0 gives 0 * 0xffffffff = 0 (a debug build would panic); n <= 1 therefore yields x/0 - y/0 = NaN in every channel.
    void get_estimated_variances(float* out) const {
        for (size_t i = 0; i < film.size(); ++i) {
            const PixelData& pd = film[i];
            const float n_times_n_minus_1 = (float)(uint32_t)(pd.n * (pd.n - 1u));
            const float n_squared_times_n_minus_1 = (float)pd.n * n_times_n_minus_1;
            const float vr = pd.sum_sq.r / n_times_n_minus_1 - pd.sum.r * pd.sum.r / n_squared_times_n_minus_1;
            const float vg = pd.sum_sq.g / n_times_n_minus_1 - pd.sum.g * pd.sum.g / n_squared_times_n_minus_1;
            const float vb = pd.sum_sq.b / n_times_n_minus_1 - pd.sum.b * pd.sum.b / n_squared_times_n_minus_1;
            out[3 * i] = vr * 50.0f;
            out[3 * i + 1] = vg * 50.0f;
            out[3 * i + 2] = vb * 50.0f;
        }
    }
};

}  // namespace orc

// =========================================================================================
// C ABI of the oracle (ctypes-friendly). Mirrors include/rt_b200.h's scene description but is
// declared independently on purpose.
// =========================================================================================
extern "C" {

struct orc_material {
    int32_t kind;  // 0 colour, 1 texture
    float rgb[3];
    uint32_t texture_id;
};
struct orc_light {
    float pos[3];
    float color[3];
};
struct orc_texture {
    uint32_t width, height;
    const float* rgb;  // width*height*3, already divided by 256 (texture.rs:40-46)
};
struct orc_scene_desc {
    uint32_t num_triangles;
    const float* vertices;      // num_triangles * 9 : v0 v1 v2 (xyz each), geometry order then triangle order
    const uint32_t* tri_geom;   // num_triangles : geometry index of each triangle (non-decreasing)
    uint32_t num_geometries;
    const orc_material* materials;  // num_geometries
    uint32_t num_lights;
    const orc_light* lights;
    uint32_t num_textures;
    const orc_texture* textures;
    float camera_orientation[16];  // vecmath (row-vector) matrix of cameras[0] (ColladaMatrix::to_vecmath_matrix)
    float camera_fov_deg;
};

void* orc_create(const orc_scene_desc* d, uint32_t width, uint32_t height, uint32_t triangles_per_leaf) {
    using namespace orc;
    RayTracer* rt = new RayTracer();
    rt->width = width;
    rt->height = height;
    rt->scene.geometries.resize(d->num_geometries);
    for (uint32_t g = 0; g < d->num_geometries; ++g) {
        Material m;
        m.kind = d->materials[g].kind;
        m.color = RGB{d->materials[g].rgb[0], d->materials[g].rgb[1], d->materials[g].rgb[2]};
        m.texture_id = d->materials[g].texture_id;
        rt->scene.geometries[g].material = m;
        rt->scene.geometries[g].first_triangle = 0xFFFFFFFFu;
    }
    for (uint32_t t = 0; t < d->num_triangles; ++t) {
        Geometry& g = rt->scene.geometries[d->tri_geom[t]];
        if (g.first_triangle == 0xFFFFFFFFu) g.first_triangle = t;
        for (int k = 0; k < 3; ++k) g.vertices.push_back(v3(d->vertices[9 * t + 3 * k], d->vertices[9 * t + 3 * k + 1], d->vertices[9 * t + 3 * k + 2]));
    }
    for (uint32_t l = 0; l < d->num_lights; ++l)
        rt->scene.lights.push_back(Light{v3(d->lights[l].pos[0], d->lights[l].pos[1], d->lights[l].pos[2]),
                                         RGB{d->lights[l].color[0], d->lights[l].color[1], d->lights[l].color[2]}});
    for (uint32_t k = 0; k < d->num_textures; ++k) {
        Texture tx;
        tx.width = d->textures[k].width;
        tx.height = d->textures[k].height;
        tx.data.resize(tx.width * tx.height);
        for (size_t i = 0; i < tx.data.size(); ++i) tx.data[i] = RGB{d->textures[k].rgb[3 * i], d->textures[k].rgb[3 * i + 1], d->textures[k].rgb[3 * i + 2]};
        rt->scene.textures.push_back(std::move(tx));
    }
    Matrix orient;
    memcpy(orient.e, d->camera_orientation, sizeof(orient.e));
    rt->camera.init(width, height, orient, d->camera_fov_deg);
    rt->octree.build(rt->scene, triangles_per_leaf);  // lib.rs:29-33
    rt->film.assign((size_t)width * height, PixelData());
    rt->primary_ids.assign((size_t)width * height, 0xFFFFFFFFu);
    rt->sample_generator.init(0);
    return rt;
}
void orc_destroy(void* h) { delete (orc::RayTracer*)h; }

// recursions (reference default 2), sub_spread (1), jitter_mode (0 fixed 0.5 / 1 hashed), seed, intersector (0 octree / 1 brute force)
void orc_configure(void* h, int recursions, uint32_t sub_spread, int jitter_mode, uint32_t seed, int intersector) {
    orc::RayTracer* rt = (orc::RayTracer*)h;
    rt->recursions = recursions;
    rt->sub_spread = sub_spread;
    rt->jitter_mode = jitter_mode;
    if (seed != rt->seed) rt->sample_generator.init(seed);
    rt->seed = seed;
    rt->intersector = intersector;
}
void orc_camera_move_rel(void* h, float x, float y, float z) {  // camera.rs:73-78
    orc::Camera& c = ((orc::RayTracer*)h)->camera;
    c.pos.x += x;
    c.pos.y += y;
    c.pos.z += z;
    c.update_matrices();
}
void orc_camera_add_x_angle(void* h, float r) {  // camera.rs:63-66
    orc::Camera& c = ((orc::RayTracer*)h)->camera;
    c.x_angle += r;
    c.update_matrices();
}
void orc_camera_add_y_angle(void* h, float r) {  // camera.rs:68-71
    orc::Camera& c = ((orc::RayTracer*)h)->camera;
    c.y_angle += r;
    c.update_matrices();
}
// out: rotation[16], orientation[16], max_x, max_y
void orc_camera_get(void* h, float* out34) {
    orc::Camera& c = ((orc::RayTracer*)h)->camera;
    memcpy(out34, c.rotation.e, 64);
    memcpy(out34 + 16, c.orientation.e, 64);
    out34[32] = c.max_x;
    out34[33] = c.max_y;
}
void orc_camera_get_ray(void* h, uint32_t u, uint32_t v, float xi1, float xi2, float* out6) {
    orc::Ray r = ((orc::RayTracer*)h)->camera.get_ray(u, v, xi1, xi2);
    out6[0] = r.pos.x; out6[1] = r.pos.y; out6[2] = r.pos.z;
    out6[3] = r.dir.x; out6[4] = r.dir.y; out6[5] = r.dir.z;
}
void orc_film_clear(void* h) { ((orc::RayTracer*)h)->film_clear(); }
uint32_t orc_trace_frame_additive(void* h, int threads) { return ((orc::RayTracer*)h)->trace_frame_additive(threads); }
void orc_trace_rows(void* h, uint32_t first_row, uint32_t n_rows, uint32_t spp, int threads) {
    ((orc::RayTracer*)h)->trace_rows(first_row, n_rows, spp, threads);
}
void orc_get_tonemapped_pixels(void* h, uint32_t* out) { ((orc::RayTracer*)h)->get_tonemapped_pixels(out); }
// out: width*height*3 floats (film.rs:50-67)
void orc_get_estimated_variances(void* h, float* out) { ((orc::RayTracer*)h)->get_estimated_variances(out); }
// test hook: overwrite the film (layout of orc_get_film: sum rgb, sum_sq rgb, n) so that edge cases of the film functions
// (n = 0, n = 1, wrapping n * (n - 1)) can be pinned and a film produced elsewhere can be run through the oracle's readers
void orc_set_film(void* h, const float* in, const uint32_t* n) {
    orc::RayTracer* rt = (orc::RayTracer*)h;
    for (size_t i = 0; i < rt->film.size(); ++i) {
        orc::PixelData& p = rt->film[i];
        const float* o = in + 7 * i;
        p.sum = orc::RGB{o[0], o[1], o[2]};
        p.sum_sq = orc::RGB{o[3], o[4], o[5]};
        p.n = n ? n[i] : (uint32_t)o[6];
    }
}
void orc_get_primary_ids(void* h, uint32_t* out) {
    orc::RayTracer* rt = (orc::RayTracer*)h;
    memcpy(out, rt->primary_ids.data(), rt->primary_ids.size() * 4);
}
// out: width*height*7 floats: sum rgb, sum_sq rgb, n
void orc_get_film(void* h, float* out) {
    orc::RayTracer* rt = (orc::RayTracer*)h;
    for (size_t i = 0; i < rt->film.size(); ++i) {
        const orc::PixelData& p = rt->film[i];
        float* o = out + 7 * i;
        o[0] = p.sum.r; o[1] = p.sum.g; o[2] = p.sum.b;
        o[3] = p.sum_sq.r; o[4] = p.sum_sq.g; o[5] = p.sum_sq.b;
        o[6] = (float)p.n;
    }
}
// out[20]: rays[3], cube_tests[3], tri_tests[3], inner_nodes[3], leaves[3], leaf_rejects[3], primary_hits, shadow_blocked
void orc_get_counters(void* h, uint64_t* out, int reset) {
    orc::RayTracer* rt = (orc::RayTracer*)h;
    const orc::Counters& c = rt->counters;
    for (int k = 0; k < 3; ++k) {
        out[k] = c.rays[k];
        out[3 + k] = c.cube_tests[k];
        out[6 + k] = c.tri_tests[k];
        out[9 + k] = c.inner_nodes[k];
        out[12 + k] = c.leaves[k];
        out[15 + k] = c.leaf_rejects[k];
    }
    out[18] = c.primary_hits;
    out[19] = c.shadow_blocked;
    if (reset) rt->counters = orc::Counters();
}
// out[6]: nodes, inner, leaves, empty leaves, triangle refs, depth
void orc_octree_stats(void* h, uint64_t* out) {
    orc::RayTracer* rt = (orc::RayTracer*)h;
    uint64_t inner = 0, leaves = 0, empty = 0, refs = 0;
    for (const orc::OctNode& n : rt->octree.nodes) {
        if (n.is_leaf) {
            leaves++;
            if (n.triangle_indices.empty()) empty++;
            refs += n.triangle_indices.size();
        } else
            inner++;
    }
    out[0] = rt->octree.nodes.size();
    out[1] = inner;
    out[2] = leaves;
    out[3] = empty;
    out[4] = refs;
    out[5] = rt->octree.max_level;
}
// Flattened export for comparison with the product's builder.
//   cubes: nodes*6 floats; first_child: nodes (int32, -1 for a leaf); leaf_offset: nodes+1 prefix of
//   triangle references (0 width for inner nodes); leaf_tris: global triangle ids
uint64_t orc_octree_export(void* h, float* cubes, int32_t* first_child, uint32_t* leaf_offset, uint32_t* leaf_tris) {
    orc::RayTracer* rt = (orc::RayTracer*)h;
    uint64_t off = 0;
    for (size_t i = 0; i < rt->octree.nodes.size(); ++i) {
        const orc::OctNode& n = rt->octree.nodes[i];
        const orc::Cube& c = rt->octree.cubes[i];
        if (cubes) {
            float* o = cubes + 6 * i;
            o[0] = c.min.x; o[1] = c.min.y; o[2] = c.min.z; o[3] = c.max.x; o[4] = c.max.y; o[5] = c.max.z;
        }
        if (first_child) first_child[i] = n.is_leaf ? -1 : (int32_t)n.children[0];
        if (leaf_offset) leaf_offset[i] = (uint32_t)off;
        if (n.is_leaf) {
            for (const orc::TriangleIndex& ti : n.triangle_indices) {
                if (leaf_tris) leaf_tris[off] = rt->scene.geometries[ti.geom_idx].first_triangle + (uint32_t)(ti.tri_idx / 3);
                off++;
            }
        }
    }
    if (leaf_offset) leaf_offset[rt->octree.nodes.size()] = (uint32_t)off;
    return off;
}

// ---- unit-level hooks used by the golden-vector tests -----------------------------------
int orc_intersect_cube_inverse_ray(const float* inv_ray6, const float* cube6, float* t) {
    orc::Ray r{orc::v3(inv_ray6[0], inv_ray6[1], inv_ray6[2]), orc::v3(inv_ray6[3], inv_ray6[4], inv_ray6[5])};
    orc::Cube c{orc::v3(cube6[0], cube6[1], cube6[2]), orc::v3(cube6[3], cube6[4], cube6[5])};
    return orc::intersect_cube_inverse_ray(r, c, t) ? 1 : 0;
}
int orc_moller_trumbore(const float* ray6, const float* tri9, float* tuv) {
    orc::Ray r{orc::v3(ray6[0], ray6[1], ray6[2]), orc::v3(ray6[3], ray6[4], ray6[5])};
    orc::HitInfo hi;
    bool ok = orc::moller_trumbore(r, orc::v3(tri9[0], tri9[1], tri9[2]), orc::v3(tri9[3], tri9[4], tri9[5]), orc::v3(tri9[6], tri9[7], tri9[8]), &hi);
    if (ok) {
        tuv[0] = hi.t; tuv[1] = hi.u; tuv[2] = hi.v;
    }
    return ok ? 1 : 0;
}
int orc_triangle_cube_intersection(const float* cube6, const float* tri9) {
    orc::Cube c{orc::v3(cube6[0], cube6[1], cube6[2]), orc::v3(cube6[3], cube6[4], cube6[5])};
    orc::Vec3 tri[3] = {orc::v3(tri9[0], tri9[1], tri9[2]), orc::v3(tri9[3], tri9[4], tri9[5]), orc::v3(tri9[6], tri9[7], tri9[8])};
    return orc::triangle_cube_intersection(c, tri) ? 1 : 0;
}
void orc_collada_matrix_to_vecmath(const float* in16, float* out16) {
    orc::Matrix m = orc::collada_to_vecmath(in16);
    memcpy(out16, m.e, 64);
}
void orc_matrix_mul(const float* a16, const float* b16, float* out16) {
    orc::Matrix a, b;
    memcpy(a.e, a16, 64);
    memcpy(b.e, b16, 64);
    orc::Matrix r = orc::mat_mul(a, b);
    memcpy(out16, r.e, 64);
}
void orc_matrix_mul_vec4(const float* m16, const float* v4, float* out4) {
    orc::Matrix m;
    memcpy(m.e, m16, 64);
    orc::Vec4 r = orc::mat_mul_vec4(m, orc::Vec4{v4[0], v4[1], v4[2], v4[3]});
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
uint32_t orc_tonemap_pack(const float* mean_rgb) { return orc::tonemap_pack(orc::RGB{mean_rgb[0], mean_rgb[1], mean_rgb[2]}); }
uint32_t orc_hash4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return orc::hash4(a, b, c, d); }
void orc_sample_table(void* h, float* out /*65536*3*/) {
    orc::RayTracer* rt = (orc::RayTracer*)h;
    memcpy(out, rt->sample_generator.normalized_vecs.data(), 65536 * 12);
}
int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
