// collada_loader.cpp — COLLADA 1.4.1 (Blender export subset) -> flattened HostScene.
//
// Behavioural restatement of raytracer_lib/src/scene/loaders/colladaloader.rs (Collada::parse :59-135,
// to_scene_flatten :137-273, to_cameras :276-319, to_lights :321-349, to_effects :351-467, to_images :469-486,
// to_materials :488-505, to_visual_scenes :507-548, convert_geometry :561-601) and of the host camera
// (scene/camera.rs). Written from scratch: a small DOM parser instead of the `parseval` combinators, the same
// library order, lookup rules, error categories and — above all — the same f32 arithmetic order for every
// number that reaches the renderer.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>

#include "host_scene.h"

namespace rtb {

// ------------------------------------------------------------------------------------------------------
// tiny XML DOM
// ------------------------------------------------------------------------------------------------------
namespace {

struct XmlNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<XmlNode>> kids;
    std::string text;

    const std::string* attr(const char* key) const {
        for (auto& kv : attrs)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const XmlNode* child(const char* nm) const {
        for (auto& k : kids)
            if (k->name == nm) return k.get();
        return nullptr;
    }
    const XmlNode* child_with_attr(const char* key, const std::string& val) const {
        for (auto& k : kids) {
            const std::string* a = k->attr(key);
            if (a && *a == val) return k.get();
        }
        return nullptr;
    }
};

struct XmlError {
    std::string what;
};

class XmlReader {
  public:
    explicit XmlReader(const std::string& s) : s_(s) {}

    void skip_ws() {
        while (i_ < s_.size() && std::isspace((unsigned char)s_[i_])) ++i_;
    }
    bool starts(const char* lit) const { return s_.compare(i_, std::strlen(lit), lit) == 0; }
    bool at_end() const { return i_ >= s_.size(); }
    size_t pos() const { return i_; }

    // <?xml ... ?>
    void prolog() {
        skip_ws();
        if (!starts("<?xml")) throw XmlError{"expected <?xml ...?> declaration"};
        size_t e = s_.find("?>", i_);
        if (e == std::string::npos) throw XmlError{"unterminated xml declaration"};
        i_ = e + 2;
    }
    void skip_misc() {
        for (;;) {
            skip_ws();
            if (starts("<!--")) {
                size_t e = s_.find("-->", i_);
                if (e == std::string::npos) throw XmlError{"unterminated comment"};
                i_ = e + 3;
            } else if (starts("<?")) {
                size_t e = s_.find("?>", i_);
                if (e == std::string::npos) throw XmlError{"unterminated processing instruction"};
                i_ = e + 2;
            } else if (starts("<!DOCTYPE")) {
                size_t e = s_.find('>', i_);
                if (e == std::string::npos) throw XmlError{"unterminated doctype"};
                i_ = e + 1;
            } else
                return;
        }
    }
    // parses "<name attr=...>" and reports whether it was self closing
    std::unique_ptr<XmlNode> open_tag(bool* self_closed) {
        skip_misc();
        if (at_end() || s_[i_] != '<' || starts("</")) throw XmlError{"expected an opening element at byte " + std::to_string(i_)};
        ++i_;
        auto node = std::make_unique<XmlNode>();
        node->name = ident();
        for (;;) {
            skip_ws();
            if (at_end()) throw XmlError{"unterminated element <" + node->name + ">"};
            if (s_[i_] == '/') {
                if (i_ + 1 >= s_.size() || s_[i_ + 1] != '>') throw XmlError{"malformed element <" + node->name + ">"};
                i_ += 2;
                *self_closed = true;
                return node;
            }
            if (s_[i_] == '>') {
                ++i_;
                *self_closed = false;
                return node;
            }
            std::string key = ident();
            skip_ws();
            if (at_end() || s_[i_] != '=') throw XmlError{"attribute without value in <" + node->name + ">"};
            ++i_;
            skip_ws();
            if (at_end() || (s_[i_] != '"' && s_[i_] != '\'')) throw XmlError{"unquoted attribute in <" + node->name + ">"};
            char q = s_[i_++];
            size_t e = s_.find(q, i_);
            if (e == std::string::npos) throw XmlError{"unterminated attribute in <" + node->name + ">"};
            node->attrs.emplace_back(key, unescape(s_.substr(i_, e - i_)));
            i_ = e + 1;
        }
    }
    void close_tag(const std::string& name) {
        skip_misc();
        if (!starts("</")) throw XmlError{"expected </" + name + ">"};
        i_ += 2;
        std::string got = ident();
        skip_ws();
        if (got != name || at_end() || s_[i_] != '>') throw XmlError{"expected </" + name + ">, found </" + got + ">"};
        ++i_;
    }
    // element content up to and including the matching close tag
    void content(XmlNode* node) {
        for (;;) {
            size_t lt = s_.find('<', i_);
            if (lt == std::string::npos) throw XmlError{"missing </" + node->name + ">"};
            node->text += s_.substr(i_, lt - i_);
            i_ = lt;
            if (starts("<!--") || starts("<?")) {
                skip_misc();
                continue;
            }
            if (starts("<![CDATA[")) {
                size_t e = s_.find("]]>", i_);
                if (e == std::string::npos) throw XmlError{"unterminated CDATA"};
                node->text += s_.substr(i_ + 9, e - i_ - 9);
                i_ = e + 3;
                continue;
            }
            if (starts("</")) {
                close_tag(node->name);
                node->text = trim(unescape(node->text));
                return;
            }
            node->kids.push_back(element());
        }
    }
    std::unique_ptr<XmlNode> element() {
        bool self_closed = false;
        auto node = open_tag(&self_closed);
        if (!self_closed) content(node.get());
        return node;
    }
    std::string rest() {
        skip_ws();
        return s_.substr(i_);
    }

  private:
    std::string ident() {
        size_t b = i_;
        while (i_ < s_.size() && !std::isspace((unsigned char)s_[i_]) && s_[i_] != '>' && s_[i_] != '/' && s_[i_] != '=') ++i_;
        if (b == i_) throw XmlError{"expected a name at byte " + std::to_string(b)};
        return s_.substr(b, i_ - b);
    }
    static std::string trim(const std::string& t) {
        size_t b = 0, e = t.size();
        while (b < e && std::isspace((unsigned char)t[b])) ++b;
        while (e > b && std::isspace((unsigned char)t[e - 1])) --e;
        return t.substr(b, e - b);
    }
    static std::string unescape(const std::string& t) {
        if (t.find('&') == std::string::npos) return t;
        std::string o;
        for (size_t k = 0; k < t.size(); ++k) {
            if (t[k] != '&') {
                o += t[k];
                continue;
            }
            static const struct { const char* ent; char ch; } ents[] = {{"&amp;", '&'}, {"&lt;", '<'}, {"&gt;", '>'}, {"&quot;", '"'}, {"&apos;", '\''}};
            bool hit = false;
            for (auto& en : ents)
                if (t.compare(k, std::strlen(en.ent), en.ent) == 0) {
                    o += en.ch;
                    k += std::strlen(en.ent) - 1;
                    hit = true;
                    break;
                }
            if (!hit) o += '&';
        }
        return o;
    }
    const std::string& s_;
    size_t i_ = 0;
};

// ------------------------------------------------------------------------------------------------------
// value helpers
// ------------------------------------------------------------------------------------------------------
struct LoadError {
    std::string what;
};

const XmlNode& need_child(const XmlNode& n, const char* name) {
    const XmlNode* c = n.child(name);
    if (!c) throw LoadError{std::string("ElementError error; no child named '") + name + "' in <" + n.name + ">"};
    return *c;
}
const XmlNode& need_child_attr(const XmlNode& n, const char* key, const std::string& val) {
    const XmlNode* c = n.child_with_attr(key, val);
    if (!c) throw LoadError{std::string("ElementError error; no child with ") + key + "='" + val + "' in <" + n.name + ">"};
    return *c;
}
const std::string& need_attr(const XmlNode& n, const char* key) {
    const std::string* a = n.attr(key);
    if (!a) throw LoadError{std::string("ElementError error; no attribute '") + key + "' on <" + n.name + ">"};
    return *a;
}
const std::string& need_data(const XmlNode& n) {
    if (!n.kids.empty()) throw LoadError{"ElementError error; <" + n.name + "> holds elements, not data"};
    return n.text;
}

// whitespace separated decimals -> f32, correctly rounded (strtof); stands in for parseval's array_f32()
std::vector<float> parse_f32_array(const std::string& s) {
    std::vector<float> out;
    const char* p = s.c_str();
    for (;;) {
        while (*p && std::isspace((unsigned char)*p)) ++p;
        if (!*p) break;
        char* e = nullptr;
        float v = std::strtof(p, &e);
        if (e == p) throw LoadError{"ParseError error; not a number near '" + std::string(p).substr(0, 16) + "'"};
        out.push_back(v);
        p = e;
    }
    if (out.empty()) throw LoadError{"ParseError error; empty number array"};
    return out;
}
std::vector<uint32_t> parse_u32_array(const std::string& s) {
    std::vector<uint32_t> out;
    const char* p = s.c_str();
    for (;;) {
        while (*p && std::isspace((unsigned char)*p)) ++p;
        if (!*p) break;
        char* e = nullptr;
        unsigned long v = std::strtoul(p, &e, 10);
        if (e == p) throw LoadError{"ParseError error; not an index near '" + std::string(p).substr(0, 16) + "'"};
        out.push_back((uint32_t)v);
        p = e;
    }
    return out;
}

struct DaeCamera {
    std::string id;
    float fov;
};
struct DaeLight {
    std::string id;
    float color[3];
};
struct DaeEffect {
    std::string id;
    bool textured = false;
    float diffuse[3] = {0, 0, 0};
    std::string image_id;
};
struct DaeImage {
    std::string id, filename;
};
struct DaeMaterial {
    std::string id, effect;
};
struct DaeGeometry {
    std::string id, material_id;
    std::vector<float> positions;
    std::vector<uint32_t> position_indices;
};
struct DaeNode {
    std::string target;
    mat4 matrix;
};

std::string strip_hash(const std::string& url) { return url.empty() ? url : url.substr(1); }

}  // namespace

// ------------------------------------------------------------------------------------------------------
// Collada::parse + to_scene_flatten
// ------------------------------------------------------------------------------------------------------
bool load_collada_str(const std::string& doc, const char* data_dir, HostScene* out, std::string* err) {
    try {
        // --- document skeleton: fixed order of top-level libraries (colladaloader.rs:60-116) ---
        XmlReader rd(doc);
        try {
            rd.prolog();
        } catch (XmlError& e) {
            throw LoadError{"XmlDefinition error; " + e.what};
        }
        std::unique_ptr<XmlNode> root;
        bool self_closed = false;
        try {
            root = rd.open_tag(&self_closed);
        } catch (XmlError& e) {
            throw LoadError{"ColladaElement error; " + e.what};
        }
        if (root->name != "COLLADA") throw LoadError{"Not a collada doc"};
        static const struct { const char* element; const char* error; } order[] = {
            {"asset", "AssetParsing"},
            {"library_cameras", "LibraryCamerasParsing"},
            {"library_lights", "LibraryLightsParsing"},
            {"library_effects", "LibraryEffectsParsing"},
            {"library_images", "LibraryImagesParsing"},
            {"library_materials", "LibraryMaterialsParsing"},
            {"library_geometries", "LibraryGeometriesParsing"},
            {"library_visual_scenes", "LibraryVisualScenesParsing"},
            {"scene", "LibrarySceneParsing"},
        };
        std::unique_ptr<XmlNode> libs[9];
        for (int k = 0; k < 9; ++k) {
            try {
                if (self_closed) throw XmlError{"document ended"};
                libs[k] = rd.element();
                if (libs[k]->name != order[k].element)
                    throw XmlError{std::string("expected <") + order[k].element + ">, found <" + libs[k]->name + ">"};
            } catch (XmlError& e) {
                throw LoadError{std::string(order[k].error) + " error; " + e.what};
            }
        }
        try {
            if (!self_closed) rd.close_tag("COLLADA");
        } catch (XmlError& e) {
            throw LoadError{"ColladaElement error; " + e.what};
        }
        std::string rest = rd.rest();
        if (!rest.empty()) throw LoadError{"RemainingData error; " + rest};

        // --- libraries -> typed records ---
        std::vector<DaeCamera> cameras;  // to_cameras
        for (auto& cam : libs[1]->kids) {
            const XmlNode& persp = need_child(need_child(need_child(*cam, "optics"), "technique_common"), "perspective");
            const XmlNode& xfov = need_child(persp, "xfov");
            const XmlNode& aspect = need_child(persp, "aspect_ratio");
            if (!xfov.kids.empty()) throw LoadError{"CamerasConversion error; cant read fov"};
            if (!aspect.kids.empty()) throw LoadError{"CamerasConversion error; cant read aspect_ratio"};
            (void)parse_f32_array(aspect.text);  // parsed and discarded
            cameras.push_back(DaeCamera{need_attr(*cam, "id"), parse_f32_array(xfov.text)[0]});
        }
        std::vector<DaeLight> lights;  // to_lights
        for (auto& li : libs[2]->kids) {
            const XmlNode& color = need_child(need_child(need_child(*li, "technique_common"), "point"), "color");
            if (!color.kids.empty()) throw LoadError{"LightsConversion error; cant get color"};
            std::vector<float> c = parse_f32_array(color.text);
            if (c.size() < 3) throw LoadError{"LightsConversion error; cant get color"};
            lights.push_back(DaeLight{need_attr(*li, "id"), {c[0], c[1], c[2]}});
        }
        std::vector<DaeEffect> effects;  // to_effects
        for (auto& ef : libs[3]->kids) {
            DaeEffect e;
            e.id = need_attr(*ef, "id");
            const XmlNode& profile = need_child(*ef, "profile_COMMON");
            const XmlNode& lambert = need_child(need_child(profile, "technique"), "lambert");
            {
                const XmlNode& em = need_child(need_child(lambert, "emission"), "color");
                if (!em.kids.empty()) throw LoadError{"EffectsConversion error; Can't get emission color"};
                if (parse_f32_array(em.text).size() < 4) throw LoadError{"EffectsConversion error; Can't get emission color"};
            }
            const XmlNode& diffuse = need_child(lambert, "diffuse");
            if (const XmlNode* col = diffuse.child("color")) {
                if (!col->kids.empty()) throw LoadError{"EffectsConversion error; Cant get diffuse color"};
                std::vector<float> c = parse_f32_array(col->text);
                if (c.size() < 4) throw LoadError{"EffectsConversion error; Cant get diffuse color"};
                e.diffuse[0] = c[0];
                e.diffuse[1] = c[1];
                e.diffuse[2] = c[2];
            } else {
                // texture (sampler sid) -> sampler2D/source (surface sid) -> surface/init_from (image id)
                const XmlNode& tex = need_child(diffuse, "texture");
                (void)need_attr(tex, "texcoord");
                const std::string& sampler = need_attr(tex, "texture");
                const XmlNode& src = need_child(need_child(need_child_attr(profile, "sid", sampler), "sampler2D"), "source");
                if (!src.kids.empty()) throw LoadError{"EffectsConversion error; Cant get sampler"};
                const XmlNode& init = need_child(need_child(need_child_attr(profile, "sid", src.text), "surface"), "init_from");
                if (!init.kids.empty()) throw LoadError{"EffectsConversion error; Cant get surface"};
                e.textured = true;
                e.image_id = init.text;
            }
            {
                const XmlNode& ior = need_child_attr(need_child(lambert, "index_of_refraction"), "sid", "ior");
                if (!ior.kids.empty()) throw LoadError{"EffectsConversion error; Can't get index of refraction"};
                (void)parse_f32_array(ior.text);
            }
            if (const XmlNode* refl = lambert.child("reflectivity")) {
                const XmlNode& sp = need_child_attr(*refl, "sid", "specular");
                if (!sp.kids.empty()) throw LoadError{"EffectsConversion error; Can't get specular"};
                (void)parse_f32_array(sp.text);
            }
            effects.push_back(e);
        }
        std::vector<DaeImage> images;  // to_images
        for (auto& im : libs[4]->kids) images.push_back(DaeImage{need_attr(*im, "id"), need_data(need_child(*im, "init_from"))});
        std::vector<DaeMaterial> materials;  // to_materials
        for (auto& m : libs[5]->kids)
            materials.push_back(DaeMaterial{need_attr(*m, "id"), strip_hash(need_attr(need_child(*m, "instance_effect"), "url"))});
        std::vector<DaeGeometry> geometries;  // to_geometries / convert_geometry
        for (auto& g : libs[6]->kids) {
            DaeGeometry dg;
            dg.id = need_attr(*g, "id");
            const XmlNode& mesh = need_child(*g, "mesh");
            const XmlNode& arr = need_child_attr(need_child_attr(mesh, "id", dg.id + "-positions"), "id", dg.id + "-positions-array");
            dg.positions = parse_f32_array(need_data(arr));
            const XmlNode& tris = need_child(mesh, "triangles");
            dg.material_id = need_attr(tris, "material");
            std::vector<uint32_t> p = parse_u32_array(need_data(need_child(tris, "p")));
            // (position, normal, texcoord) index triples; only the position index is kept (:588-593)
            if (p.size() % 3 != 0) throw LoadError{"GeometryConversion error"};
            for (size_t k = 0; k + 2 < p.size(); k += 3) dg.position_indices.push_back(p[k]);
            geometries.push_back(std::move(dg));
        }
        std::vector<DaeNode> nodes;  // to_visual_scenes: every <node> of every <visual_scene>, in order
        for (auto& vs : libs[7]->kids) {
            for (auto& nd : vs->kids) {
                const XmlNode* inst = nd->child("instance_light");
                if (!inst) inst = nd->child("instance_geometry");
                if (!inst) inst = nd->child("instance_camera");
                if (!inst) throw LoadError{"VisualSceneConversion error; unsupported node type"};
                std::string target = strip_hash(need_attr(*inst, "url"));
                const XmlNode& mx = need_child(*nd, "matrix");
                if (!mx.kids.empty()) continue;
                std::vector<float> m = parse_f32_array(mx.text);
                if (m.size() < 16) throw LoadError{"VisualSceneConversion error; cant create array"};
                nodes.push_back(DaeNode{target, collada_node_matrix(m.data())});
            }
        }
        if (libs[7]->kids.empty() && libs[7]->text.size()) throw LoadError{"VisualSceneConversion error; No scene element(s)"};

        // --- to_scene_flatten (:137-273) ---
        HostScene sc;
        for (const DaeImage& im : images) {
            std::string path = data_dir && *data_dir ? std::string(data_dir) + "/" + im.filename : im.filename;
            uint32_t w = 0, h = 0;
            std::vector<uint8_t> rgb8;
            std::string perr;
            if (!decode_png_rgb8(path, &w, &h, &rgb8, &perr)) throw LoadError{perr};
            HostTexture t;
            t.width = w;
            t.height = h;
            t.rgb.resize(rgb8.size());
            for (size_t k = 0; k < rgb8.size(); ++k) t.rgb[k] = (float)rgb8[k] / 256.0f;  // texture.rs:42-44
            sc.textures.push_back(std::move(t));
        }
        for (const DaeNode& node : nodes) {
            for (const DaeCamera& cam : cameras) {
                if (cam.id != node.target) continue;
                if (!sc.has_camera) {  // only cameras[0] is ever used (lib.rs:39)
                    sc.camera_orientation = node.matrix;
                    sc.camera_fov_deg = cam.fov;
                    sc.has_camera = true;
                }
                break;
            }
            for (const DaeLight& li : lights) {
                if (li.id != node.target) continue;
                f3 p = row_times4(node.matrix, 0.f, 0.f, 0.f, 1.f);
                rt_light L;
                L.pos[0] = p.x;
                L.pos[1] = p.y;
                L.pos[2] = p.z;
                std::memcpy(L.color, li.color, sizeof(L.color));
                sc.lights.push_back(L);
                break;
            }
            for (const DaeGeometry& g : geometries) {
                if (g.id != node.target) continue;
                const uint32_t geom_index = (uint32_t)sc.materials.size();
                const size_t ntri = g.position_indices.size() / 3;
                for (size_t t = 0; t < ntri; ++t) {
                    for (int c = 0; c < 3; ++c) {
                        size_t vi = 3 * (size_t)g.position_indices[3 * t + c];
                        if (vi + 2 >= g.positions.size()) throw LoadError{"GeometryConversion error"};
                        f3 q = row_times4(node.matrix, g.positions[vi], g.positions[vi + 1], g.positions[vi + 2], 1.f);
                        sc.vertices.push_back(q.x);
                        sc.vertices.push_back(q.y);
                        sc.vertices.push_back(q.z);
                    }
                    sc.tri_geom.push_back(geom_index);
                }
                rt_material mat;  // Material::default(): Diffuse::Color(RGB::default()) = (1000, 0, 1000)   color.rs:37-41
                mat.kind = RT_DIFFUSE_COLOR;
                mat.rgb[0] = 1000.f;
                mat.rgb[1] = 0.f;
                mat.rgb[2] = 1000.f;
                mat.texture_id = 0;
                const DaeMaterial* dm = nullptr;
                for (const DaeMaterial& m : materials)
                    if (m.id == g.material_id) {
                        dm = &m;
                        break;
                    }
                if (dm) {
                    for (const DaeEffect& e : effects) {
                        if (e.id != dm->effect) continue;
                        if (!e.textured) {
                            std::memcpy(mat.rgb, e.diffuse, sizeof(mat.rgb));
                        } else {
                            int pos = -1;
                            for (size_t k = 0; k < images.size(); ++k)
                                if (images[k].id == e.image_id) {
                                    pos = (int)k;
                                    break;
                                }
                            if (pos < 0) throw LoadError{"MaterialsConversion error; can't find texture name"};
                            mat.kind = RT_DIFFUSE_TEXTURE;
                            mat.rgb[0] = mat.rgb[1] = mat.rgb[2] = 0.f;
                            mat.texture_id = (uint32_t)pos;
                        }
                        break;
                    }
                }
                sc.materials.push_back(mat);
                sc.geometry_ids.push_back(g.id);
                break;
            }
        }
        *out = std::move(sc);
        return true;
    } catch (LoadError& e) {
        *err = e.what;
    } catch (XmlError& e) {
        *err = "ParseError error; " + e.what;
    } catch (std::exception& e) {
        *err = std::string("ParseError error; ") + e.what();
    }
    return false;
}

bool load_collada_file(const std::string& path, HostScene* out, std::string* err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) {
        *err = "No such file or directory (os error 2)";  // io::Error Display, as SceneLoadError::Io prints it
        return false;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    std::string dir;  // Path::parent()
    size_t slash = path.find_last_of('/');
    if (slash != std::string::npos) dir = path.substr(0, slash == 0 ? 1 : slash);
    return load_collada_str(ss.str(), dir.c_str(), out, err);
}

// ------------------------------------------------------------------------------------------------------
// HostScene <-> rt_scene_desc
// ------------------------------------------------------------------------------------------------------
void HostScene::make_desc(rt_scene_desc* d) {
    texture_views.clear();
    for (const HostTexture& t : textures) texture_views.push_back(rt_texture{t.width, t.height, t.rgb.data()});
    d->num_triangles = num_triangles();
    d->vertices = vertices.data();
    d->tri_geom = tri_geom.data();
    d->num_geometries = (uint32_t)materials.size();
    d->materials = materials.data();
    d->num_lights = (uint32_t)lights.size();
    d->lights = lights.data();
    d->num_textures = (uint32_t)texture_views.size();
    d->textures = texture_views.data();
    for (int i = 0; i < 16; ++i) d->camera_orientation[i] = camera_orientation[i];
    d->camera_fov_deg = camera_fov_deg;
    // the NORMAL / TEXCOORD inputs are parsed past, not kept, as in the reference (colladaloader.rs:587-593)
    d->normals = nullptr;
    d->uvs = nullptr;
}

HostScene HostScene::from_desc(const rt_scene_desc& d) {
    HostScene s;
    s.vertices.assign(d.vertices, d.vertices + (size_t)d.num_triangles * 9);
    s.tri_geom.assign(d.tri_geom, d.tri_geom + d.num_triangles);
    s.materials.assign(d.materials, d.materials + d.num_geometries);
    s.lights.assign(d.lights, d.lights + d.num_lights);
    for (uint32_t k = 0; k < d.num_textures; ++k) {
        HostTexture t;
        t.width = d.textures[k].width;
        t.height = d.textures[k].height;
        t.rgb.assign(d.textures[k].rgb, d.textures[k].rgb + (size_t)t.width * t.height * 3);
        s.textures.push_back(std::move(t));
    }
    for (int i = 0; i < 16; ++i) s.camera_orientation[i] = d.camera_orientation[i];
    s.camera_fov_deg = d.camera_fov_deg;
    s.has_camera = true;
    return s;
}

// ------------------------------------------------------------------------------------------------------
// camera (scene/camera.rs)
// ------------------------------------------------------------------------------------------------------
void HostCamera::init(uint32_t w, uint32_t h, const mat4& orientation_matrix, float fov_deg) {
    mat4 rot = orientation_matrix;  // rotation part only: clear the last column and the translation row
    rot[3] = rot[7] = rot[11] = 0.f;
    rot[12] = rot[13] = rot[14] = 0.f;
    rot[15] = 1.f;
    const float fov = fov_deg * 3.14159265358979323846f / 180.0f;
    const float half_fov = 0.5f * fov;
    max_x = 1.0f * std::tan(half_fov);
    max_y = 1.0f * std::tan(half_fov);
    x_angle = y_angle = 0.f;
    pos = f3{};
    width = w;
    height = h;
    base_orientation = orientation_matrix;
    base_rotation = rot;
    update_matrices();
}

void HostCamera::update_matrices() {
    rotation = product4(product4(rot_x4(x_angle), rot_y4(y_angle)), base_rotation);
    orientation = product4(product4(rotation, translate4(pos)), base_orientation);
}

}  // namespace rtb
