// obj_loader.cpp — host-side OBJ/MTL loader: Object::load_from_file (mororo18/draw
// src/renderer/scene/object.rs:106-454) behind draw_object_load_obj.  Host only, no CUDA.
//
// The reference parses with the `obj` crate (kvark/obj, unpinned, Cargo.toml:13; its source is not in
// the reference tree).  The crate only tokenises text; the rules object.rs relies on are restated
// here: statements v / vt (first two floats) / vn / f / o / g / usemtl / mtllib, `s` and `l` ignored,
// implicit object and group named "default", `o` closes the current group and object, `g` closes the
// current group, a `usemtl` on a group that already has a material closes it and opens a new group of
// the same name, face indices 1-based with negatives relative to the current count.
//
// Semantics of object.rs that are kept (cited inline): normals normalised on load, uv -> (u, v, 0),
// rescale to radius 100, textures[0] = Texture::default(), materials in file order, quads split
// (a,b,c),(c,d,a), one dummy uv per group lacking uvs, material lookup by name with fallback 0,
// generated normals = normalised sum of un-normalised face normals over that group, appended.
// Documented deviations (the reference panics there): a material without Ka/Kd/Ks takes
// the Texture::default() value; faces with < 3 vertices are skipped; faces with > 4 vertices are an
// error (the reference hits todo!()).
// Float arithmetic is binary32 in the reference's order; this file is built with -ffp-contract=off.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/draw_b200.h"

namespace drawb200 {
int loader_fail(int code, const char *msg);
}

namespace {

struct Img {
    uint8_t *pixels = nullptr; // malloc'ed by the image loader callback, freed with free()
    uint32_t w = 0, h = 0, comp = 0;
};
struct Mat {
    std::string name;
    float ka[3] = {0.9f, 0.9f, 0.9f}, kd[3] = {0.4f, 0.4f, 0.4f}, ks[3] = {0.5f, 0.5f, 0.5f}; // scene/mod.rs:243-245
    float alpha = 1.0f;
    int map_ka = -1, map_kd = -1; // index into draw_object::images
};
struct MeshOut {
    std::string name;
    std::vector<uint32_t> tris; // 9 per triangle
    uint32_t material = 0;
};

struct Corner {
    long v;
    long vt, vn; // -1 = absent
};
struct Group {
    std::string name;
    bool has_material = false;
    std::string material;
    std::vector<std::vector<Corner>> polys;
};

std::vector<std::string> split_ws(const std::string &line) {
    std::vector<std::string> out;
    std::istringstream ss(line);
    std::string w;
    while (ss >> w) out.push_back(w);
    return out;
}
std::string trim(const std::string &s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string dir_of(const std::string &path) {
    size_t p = path.find_last_of('/');
    return p == std::string::npos ? std::string() : path.substr(0, p);
}
std::string base_of(const std::string &path) {
    size_t p = path.find_last_of('/');
    return p == std::string::npos ? path : path.substr(p + 1);
}
std::string join(const std::string &dir, const std::string &name) { return dir.empty() ? name : dir + "/" + name; }
float parse_f32(const std::string &s) { return std::strtof(s.c_str(), nullptr); } // correctly rounded, like Rust's parse

float norm3(const float *a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); } // linalg.rs:167-171

} // namespace

struct draw_object {
    std::string name;
    std::vector<float> pos, nrm, uv; // 3 floats each
    std::vector<MeshOut> meshes;
    std::vector<Mat> materials;
    std::vector<Img> images;
    // borrowed views handed out by draw_object_desc_of
    std::vector<draw_mesh> c_meshes;
    std::vector<draw_material> c_materials;
    ~draw_object() {
        for (Img &i : images) std::free(i.pixels);
    }
};

extern "C" {

int draw_object_load_obj(const char *path_c, draw_image_loader loader, void *user, draw_object **out) {
    if (out) *out = nullptr;
    if (!path_c || !out) return drawb200::loader_fail(DRAW_ERR_INVALID_ARGUMENT, "draw_object_load_obj: NULL argument");
    draw_object *o = nullptr; // owned here until handed to the caller: the handlers below release it
    try {
        const std::string path(path_c), dir = dir_of(path);
        std::ifstream in(path);
        if (!in) return drawb200::loader_fail(DRAW_ERR_IO, ("Unable to open file " + path).c_str()); // object.rs:127-128

        std::vector<float> position, texture, normal; // 3 / 2 / 3 floats
        std::vector<std::pair<std::string, std::vector<Group>>> objects;
        std::vector<std::string> mtllibs;
        std::string obj_name = "default";
        std::vector<Group> obj_groups;
        Group group;
        bool have_group = false;

        auto fix = [](long idx, size_t count) -> long { return idx > 0 ? idx - 1 : (long)count + idx; };
        std::string line;
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            const std::vector<std::string> w = split_ws(line);
            if (w.empty()) continue;
            const std::string &key = w[0];
            if (key == "v" && w.size() >= 4) {
                for (int i = 1; i <= 3; i++) position.push_back(parse_f32(w[i]));
            } else if (key == "vt" && w.size() >= 3) {
                texture.push_back(parse_f32(w[1]));
                texture.push_back(parse_f32(w[2]));
            } else if (key == "vn" && w.size() >= 4) {
                for (int i = 1; i <= 3; i++) normal.push_back(parse_f32(w[i]));
            } else if (key == "f") {
                std::vector<Corner> poly;
                for (size_t i = 1; i < w.size(); i++) {
                    Corner c{0, -1, -1};
                    const std::string &t = w[i];
                    const size_t s1 = t.find('/');
                    const size_t s2 = s1 == std::string::npos ? std::string::npos : t.find('/', s1 + 1);
                    c.v = fix(std::strtol(t.substr(0, s1).c_str(), nullptr, 10), position.size() / 3);
                    if (s1 != std::string::npos) {
                        const std::string vt = t.substr(s1 + 1, s2 == std::string::npos ? std::string::npos : s2 - s1 - 1);
                        if (!vt.empty()) c.vt = fix(std::strtol(vt.c_str(), nullptr, 10), texture.size() / 2);
                        if (s2 != std::string::npos) {
                            const std::string vn = t.substr(s2 + 1);
                            if (!vn.empty()) c.vn = fix(std::strtol(vn.c_str(), nullptr, 10), normal.size() / 3);
                        }
                    }
                    poly.push_back(c);
                }
                if (!have_group) {
                    group = Group();
                    group.name = "default";
                    have_group = true;
                }
                group.polys.push_back(std::move(poly));
            } else if (key == "o") {
                if (have_group) {
                    obj_groups.push_back(group);
                    objects.push_back({obj_name, obj_groups});
                    have_group = false;
                }
                obj_name = line.size() > 2 ? trim(line.substr(1)) : "default";
                obj_groups.clear();
            } else if (key == "g") {
                if (have_group) {
                    obj_groups.push_back(group);
                    have_group = false;
                }
                if (line.size() > 2) {
                    group = Group();
                    group.name = trim(line.substr(2));
                    have_group = true;
                }
            } else if (key == "mtllib") {
                for (size_t i = 1; i < w.size(); i++) mtllibs.push_back(w[i]);
            } else if (key == "usemtl") {
                if (!have_group) {
                    group = Group();
                    group.name = "default";
                    have_group = true;
                }
                if (group.has_material) {
                    obj_groups.push_back(group);
                    group.polys.clear();
                }
                group.has_material = w.size() > 1;
                group.material = w.size() > 1 ? w[1] : std::string();
            }
            // s, l, comments, unknown statements: ignored
        }
        if (have_group) obj_groups.push_back(group);
        objects.push_back({obj_name, obj_groups});

        o = new draw_object();
        o->name = base_of(path); // object.rs:444
        const size_t n_pos = position.size() / 3;
        if (n_pos == 0) {
            delete o;
            o = nullptr;
            return drawb200::loader_fail(DRAW_ERR_IO, "OBJ file has no vertices");
        }
        // normals normalised (object.rs:146-150): three divisions by the norm
        o->nrm.resize(normal.size());
        for (size_t i = 0; i < normal.size() / 3; i++) {
            const float n = norm3(&normal[3 * i]);
            for (int c = 0; c < 3; c++) o->nrm[3 * i + c] = normal[3 * i + c] / n;
        }
        // uv -> (u, v, 0)  (object.rs:151-155)
        for (size_t i = 0; i < texture.size() / 2; i++) {
            o->uv.push_back(texture[2 * i]);
            o->uv.push_back(texture[2 * i + 1]);
            o->uv.push_back(0.0f);
        }
        // rescale (object.rs:159-170): factor = 100 / max |v| ; v = v * factor
        float vmax = -INFINITY;
        for (size_t i = 0; i < n_pos; i++) {
            const float n = norm3(&position[3 * i]);
            if (n > vmax) vmax = n;
        }
        const float factor = 100.0f / vmax;
        o->pos.resize(position.size());
        for (size_t i = 0; i < position.size(); i++) o->pos[i] = position[i] * factor;

        // materials (object.rs:172-221)
        o->materials.push_back(Mat{}); // Texture::default(), name "default"
        o->materials[0].name = "default";
        auto load_img = [&](const std::string &file) -> int {
            if (!loader) return -1;
            Img img;
            const std::string full = join(dir, file);
            if (loader(full.c_str(), user, &img.pixels, &img.w, &img.h, &img.comp) != 0 || !img.pixels ||
                (img.comp != 3 && img.comp != 4))
                throw std::runtime_error("image loader failed for " + full);
            o->images.push_back(img);
            return (int)o->images.size() - 1;
        };
        for (const std::string &lib : mtllibs) {
            std::ifstream mf(join(dir, lib));
            if (!mf) {
                delete o;
                o = nullptr;
                return drawb200::loader_fail(DRAW_ERR_IO, ("Unable to open file " + lib).c_str()); // object.rs:136-137
            }
            int cur = -1;
            while (std::getline(mf, line)) {
                const std::vector<std::string> w = split_ws(line);
                if (w.empty() || w[0][0] == '#') continue;
                if (w[0] == "newmtl") {
                    o->materials.push_back(Mat{});
                    cur = (int)o->materials.size() - 1;
                    o->materials[cur].name = w.size() > 1 ? w[1] : std::string();
                } else if (cur < 0) {
                    continue;
                } else if ((w[0] == "Ka" || w[0] == "Kd" || w[0] == "Ks") && w.size() >= 4) {
                    float *dst = w[0] == "Ka" ? o->materials[cur].ka : (w[0] == "Kd" ? o->materials[cur].kd : o->materials[cur].ks);
                    for (int c = 0; c < 3; c++) dst[c] = parse_f32(w[1 + c]);
                } else if (w[0] == "d" && w.size() >= 2) {
                    o->materials[cur].alpha = parse_f32(w[1]);
                } else if (w[0] == "map_Ka" && w.size() >= 2) {
                    o->materials[cur].map_ka = load_img(w[1]);
                } else if (w[0] == "map_Kd" && w.size() >= 2) {
                    // the same file referenced twice is decoded once
                    o->materials[cur].map_kd = load_img(w[1]);
                }
            }
        }

        // groups -> meshes (object.rs:230-442)
        for (const auto &ob : objects) {
            for (const Group &g : ob.second) {
                if (g.polys.empty()) continue; // :235
                struct T3 { long v[3], t[3], n[3]; bool has_t, has_n; };
                std::vector<T3> tris;
                bool missing_tex = false, missing_nrm = false;
                for (const std::vector<Corner> &face : g.polys) {
                    if (face.size() < 3) continue;
                    if (face.size() > 4) {
                        delete o;
                        o = nullptr;
                        return drawb200::loader_fail(DRAW_ERR_IO, "faces with more than 4 vertices are not supported (object.rs:361-363)");
                    }
                    bool f_mt = false, f_mn = false;
                    for (const Corner &c : face) {
                        f_mt |= c.vt < 0;
                        f_mn |= c.vn < 0;
                        if (c.v < 0 || (size_t)c.v >= n_pos) {
                            delete o;
                            o = nullptr;
                            return drawb200::loader_fail(DRAW_ERR_IO, "face references a vertex that does not exist");
                        }
                    }
                    missing_tex |= f_mt;
                    missing_nrm |= f_mn;
                    auto push = [&](int a, int b, int c) {
                        T3 t{};
                        const int ix[3] = {a, b, c};
                        for (int k = 0; k < 3; k++) {
                            t.v[k] = face[ix[k]].v;
                            t.t[k] = face[ix[k]].vt;
                            t.n[k] = face[ix[k]].vn;
                        }
                        t.has_t = !f_mt;
                        t.has_n = !f_mn;
                        tris.push_back(t);
                    };
                    push(0, 1, 2);                        // :306-331
                    if (face.size() == 4) push(2, 3, 0);  // :333-360
                }
                if (missing_tex) { // :368-384
                    const long dummy = (long)(o->uv.size() / 3);
                    o->uv.insert(o->uv.end(), {0.0f, 0.0f, 0.0f});
                    for (T3 &t : tris)
                        if (!t.has_t) t.t[0] = t.t[1] = t.t[2] = dummy;
                }
                uint32_t material = 0; // :366, 387-392
                const std::string want = g.has_material ? g.material : std::string("default");
                for (size_t i = 0; i < o->materials.size(); i++)
                    if (o->materials[i].name == want) {
                        material = (uint32_t)i;
                        break;
                    }
                if (missing_nrm) { // :394-429
                    const size_t n_before = o->nrm.size() / 3;
                    std::vector<float> gen(3 * n_pos, 0.0f);
                    for (const T3 &t : tris) {
                        const float *a = &o->pos[3 * t.v[0]], *b = &o->pos[3 * t.v[1]], *c = &o->pos[3 * t.v[2]];
                        const float p[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, q[3] = {c[0] - b[0], c[1] - b[1], c[2] - b[2]};
                        const float nx = (p[1] * q[2]) - (p[2] * q[1]), ny = (p[2] * q[0]) - (p[0] * q[2]),
                                    nz = (p[0] * q[1]) - (p[1] * q[0]); // mesh.rs:15-28, linalg.rs:186-200
                        for (int k = 0; k < 3; k++) {
                            float *dst = &gen[3 * t.v[k]];
                            dst[0] = dst[0] + nx;
                            dst[1] = dst[1] + ny;
                            dst[2] = dst[2] + nz;
                        }
                    }
                    for (size_t i = 0; i < n_pos; i++) {
                        const float n = norm3(&gen[3 * i]);
                        for (int c = 0; c < 3; c++) gen[3 * i + c] = gen[3 * i + c] / n; // 0/0 -> NaN for unused vertices, as in the reference
                    }
                    for (T3 &t : tris)
                        if (!t.has_n)
                            for (int k = 0; k < 3; k++) t.n[k] = t.v[k] + (long)n_before;
                    o->nrm.insert(o->nrm.end(), gen.begin(), gen.end());
                }
                MeshOut m;
                m.name = g.name;
                m.material = material;
                for (const T3 &t : tris) {
                    for (int k = 0; k < 3; k++) m.tris.push_back((uint32_t)t.v[k]);
                    for (int k = 0; k < 3; k++) m.tris.push_back((uint32_t)t.t[k]);
                    for (int k = 0; k < 3; k++) m.tris.push_back((uint32_t)t.n[k]);
                }
                o->meshes.push_back(std::move(m));
            }
        }
        *out = o;
        return DRAW_OK;
    } catch (const std::exception &e) {
        delete o; // and the images decoded so far
        return drawb200::loader_fail(DRAW_ERR_IO, e.what());
    } catch (...) {
        delete o;
        o = nullptr;
        return drawb200::loader_fail(DRAW_ERR_INTERNAL, "draw_object_load_obj: internal error");
    }
}

void draw_object_free(draw_object *obj) { delete obj; }

int draw_object_desc_of(const draw_object *obj_c, draw_object_desc *out) {
    if (!obj_c || !out) return drawb200::loader_fail(DRAW_ERR_INVALID_ARGUMENT, "draw_object_desc_of: NULL argument");
    draw_object *obj = const_cast<draw_object *>(obj_c);
    obj->c_meshes.clear();
    obj->c_materials.clear();
    for (const MeshOut &m : obj->meshes)
        obj->c_meshes.push_back(draw_mesh{m.name.c_str(), m.tris.data(), m.tris.size() / 9, m.material});
    for (const Mat &m : obj->materials) {
        draw_material d{};
        d.name = m.name.c_str();
        for (int c = 0; c < 3; c++) { d.ka[c] = m.ka[c]; d.kd[c] = m.kd[c]; d.ks[c] = m.ks[c]; }
        d.alpha = m.alpha;
        auto map = [&](int idx) {
            draw_texture_map t{nullptr, 0, 0, 0};
            if (idx >= 0) {
                const Img &i = obj->images[idx];
                t = draw_texture_map{i.pixels, i.w, i.h, i.comp};
            }
            return t;
        };
        d.map_ka = map(m.map_ka);
        d.map_kd = map(m.map_kd);
        obj->c_materials.push_back(d);
    }
    out->name = obj->name.c_str();
    out->positions = obj->pos.data(); out->n_positions = obj->pos.size() / 3;
    out->normals = obj->nrm.data();   out->n_normals = obj->nrm.size() / 3;
    out->uvs = obj->uv.data();        out->n_uvs = obj->uv.size() / 3;
    out->meshes = obj->c_meshes.data();       out->n_meshes = obj->c_meshes.size();
    out->materials = obj->c_materials.data(); out->n_materials = obj->c_materials.size();
    return DRAW_OK;
}

} // extern "C"
