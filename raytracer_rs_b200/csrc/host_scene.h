// host_scene.h — the flattened scene as it crosses the C ABI (structure of arrays), plus the host camera.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "host_math.h"

namespace rtb {

struct HostTexture {
    uint32_t width = 0, height = 0;
    std::vector<float> rgb;  // byte / 256 (scene/texture.rs:40-46)
};

// Flattened `Scene` (scene/mod.rs:24-29): triangle soup in geometry order, one material per geometry.
struct HostScene {
    std::vector<float> vertices;     // 9 floats per triangle
    std::vector<uint32_t> tri_geom;  // geometry index per triangle
    std::vector<rt_material> materials;
    std::vector<rt_light> lights;
    std::vector<HostTexture> textures;
    std::vector<std::string> geometry_ids;
    mat4 camera_orientation = identity4();
    float camera_fov_deg = 0.f;
    bool has_camera = false;

    std::vector<rt_texture> texture_views;  // filled by make_desc
    void make_desc(rt_scene_desc* d);
    static HostScene from_desc(const rt_scene_desc& d);
    uint32_t num_triangles() const { return (uint32_t)tri_geom.size(); }
};

// scene/loaders/colladaloader.rs: Collada::parse + to_scene_flatten. Returns false and fills *err with the
// Display text of the corresponding SceneLoadError on failure.
bool load_collada_str(const std::string& doc, const char* data_dir, HostScene* out, std::string* err);
bool load_collada_file(const std::string& path, HostScene* out, std::string* err);
bool decode_png_rgb8(const std::string& path, uint32_t* width, uint32_t* height, std::vector<uint8_t>* rgb, std::string* err);

// scene/camera.rs `Camera`
struct HostCamera {
    float x_angle = 0.f, y_angle = 0.f;
    f3 pos;
    uint32_t width = 0, height = 0;
    mat4 base_orientation = identity4(), base_rotation = identity4();
    mat4 orientation = identity4(), rotation = identity4();
    float max_x = 0.f, max_y = 0.f;

    void init(uint32_t w, uint32_t h, const mat4& orientation_matrix, float fov_deg);  // camera.rs:22-61
    void update_matrices();                                                            // camera.rs:92-98
    void move_rel(float x, float y, float z) {                                         // camera.rs:73-78
        pos.x += x;
        pos.y += y;
        pos.z += z;
        update_matrices();
    }
    void add_x_angle(float r) {  // camera.rs:63-66
        x_angle += r;
        update_matrices();
    }
    void add_y_angle(float r) {  // camera.rs:68-71
        y_angle += r;
        update_matrices();
    }
    f3 ray_origin() const { return row_times4(orientation, 0.f, 0.f, 0.f, 1.f); }  // camera.rs:88
};

}  // namespace rtb
