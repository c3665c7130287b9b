// host_math.h — f32 vector/matrix helpers for the host side of rt_b200.
//
// The operation order mirrors raytracer_lib/src/vecmath.rs (row-vector convention, 4-term sums evaluated
// left to right) because the bit pattern of every vertex, light position and camera matrix that reaches
// the device depends on it. Host translation units are compiled with
// -ffp-contract=off so no multiply-add is ever fused.
#pragma once
#include <array>
#include <cmath>

namespace rtb {

struct f3 {
    float x = 0.f, y = 0.f, z = 0.f;
};
inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 operator*(float s, f3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // vecmath.rs:74-76
inline f3 cross3(f3 a, f3 b) {                                                // vecmath.rs:79-85
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline f3 unit3(f3 v) {  // Vec3::normalized, vecmath.rs:23-26
    const float len = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return {v.x / len, v.y / len, v.z / len};
}

// 4x4 matrix, row-major storage, used with row vectors: v' = v * M, translation in m[12..14].
using mat4 = std::array<float, 16>;

inline mat4 identity4() { return {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }

inline mat4 product4(const mat4& l, const mat4& r) {  // vecmath.rs:237-313
    mat4 o{};
    for (int row = 0; row < 4; ++row) {
        const float* lr = &l[4 * row];
        for (int col = 0; col < 4; ++col)
            o[4 * row + col] = lr[0] * r[col] + lr[1] * r[4 + col] + lr[2] * r[8 + col] + lr[3] * r[12 + col];
    }
    return o;
}
inline mat4 transposed4(const mat4& m) {
    mat4 o{};
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) o[4 * c + r] = m[4 * r + c];
    return o;
}
// (x, y, z, w) * M  -> first three components                                       vecmath.rs:200-211
inline f3 row_times4(const mat4& m, float x, float y, float z, float w) {
    return {x * m[0] + y * m[4] + z * m[8] + w * m[12], x * m[1] + y * m[5] + z * m[9] + w * m[13],
            x * m[2] + y * m[6] + z * m[10] + w * m[14]};
}
inline mat4 rot_x4(float r) {  // vecmath.rs:116-123
    mat4 m = identity4();
    m[5] = std::cos(r);
    m[6] = -std::sin(r);
    m[9] = std::sin(r);
    m[10] = std::cos(r);
    return m;
}
inline mat4 rot_y4(float r) {  // vecmath.rs:124-131
    mat4 m = identity4();
    m[0] = std::cos(r);
    m[2] = std::sin(r);
    m[8] = -std::sin(r);
    m[10] = std::cos(r);
    return m;
}
inline mat4 translate4(f3 t) {  // vecmath.rs:133-139
    mat4 m = identity4();
    m[12] = t.x;
    m[13] = t.y;
    m[14] = t.z;
    return m;
}

// Collada node matrix (column-major meaning, Z up, right handed) -> row-vector, Y up, left handed:
// reflect_z * transpose(M) * swap_yz, two f32 products in this order (collada_types.rs:76-90).
inline mat4 collada_node_matrix(const float* sixteen) {
    mat4 c{};
    for (int i = 0; i < 16; ++i) c[i] = sixteen[i];
    const mat4 swap_yz = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
    const mat4 reflect_z = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1};
    return product4(product4(reflect_z, transposed4(c)), swap_yz);
}

}  // namespace rtb
