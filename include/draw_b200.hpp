// draw_b200.hpp — header-only C++ mirror of the reference's Rust host API over the C ABI (draw_b200.h).
//
// The reference is compiled code (Rust); its toolchain is absent from this image, so the host side above
// the C ABI is given here in C++ with the reference's own type and method names, argument meaning and
// error behaviour: where the Rust code panics (assert!/unwrap/expect) these methods throw draw::Error
// carrying draw_last_error().  Nothing here computes: every method is one call into libdraw_b200.so.
//
//   reference (file:line)                              here
//   Scene::new(w, h)            scene/mod.rs:760       draw::Scene(w, h)
//   scene.add_obj(Object)       scene/mod.rs:788       Scene::add_obj(const Object &) -> ObjectInfo
//   scene.camera = Camera::new  scene/mod.rs:297,752   Scene::set_camera(pos, dir) / camera()
//   camera.move_up ... backward scene/mod.rs:381-405   Scene::camera_move_up(d) ... camera_move_backward(d)
//   scene.move_camera_direction scene/mod.rs:803       Scene::move_camera_direction(dx, dy)
//   scene.render(&mut canvas)   scene/mod.rs:901       Scene::render(Canvas &)
//   Canvas::new(w, h)           canvas.rs:366          draw::Canvas(w, h)
//   init_depth / apply_offset / resize / clear         same names          canvas.rs:382-433
//   enable/disable_depth_update canvas.rs:395-401      same names
//   as_bytes_slice / size_bytes / pixel_bytes          same names          canvas.rs:966-982
//   get_pixel_depth(x, y)       canvas.rs:413          Canvas::depth_frame() (whole buffer) / get_pixel_depth
//   Object::new(...)            object.rs:34           draw::Object aggregate (same fields)
//   Object::load_from_file      object.rs:106          Object::load_from_file(path)
//   TextureMap::load_from_file  scene/mod.rs:174       TextureMap::load_from_file(path)   (PNG)
//   Application::export_frame_as(Png) app/mod.rs:316   Canvas::export_png(path)
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "draw_b200.h"

namespace draw {

struct Error : std::runtime_error {
    int status;
    Error(int status_, const char *msg) : std::runtime_error(msg), status(status_) {}
};
inline void check(int rc) {
    if (rc != DRAW_OK) throw Error(rc, draw_last_error());
}

using Vec3 = std::array<float, 3>;

// TextureMap (scene/mod.rs:102-110): bytes as stb_image returns them, row 0 = top; empty = TextureMap::default().
struct TextureMap {
    std::vector<uint8_t> img;
    uint32_t width = 0, height = 0, components = 0;
    static TextureMap load_from_file(const std::string &path) { // scene/mod.rs:174-202
        uint8_t *px = nullptr;
        TextureMap m;
        check(draw_image_load(path.c_str(), &px, &m.width, &m.height, &m.components));
        m.img.assign(px, px + (size_t)m.width * m.height * m.components);
        draw_image_free(px);
        return m;
    }
};

// Texture (scene/mod.rs:206-216); defaults of Texture::default() (:243-245).
struct Texture {
    std::string name = "default";
    Vec3 ka{0.9f, 0.9f, 0.9f}, kd{0.4f, 0.4f, 0.4f}, ks{0.5f, 0.5f, 0.5f};
    float alpha = 1.0f;
    TextureMap map_ka, map_kd;
};

// IndexedMesh (mesh.rs:31-35): one (vertex, texture, normal) index triple per triangle, flattened to
// v0 v1 v2 t0 t1 t2 n0 n1 n2.
struct IndexedMesh {
    std::string name;
    std::vector<uint32_t> triangles; // 9 per triangle
    uint32_t texture_idx = 0;
};

struct ObjectInfo { // scene/mod.rs:256
    uint32_t id;
};

// Object (object.rs:18-31) with the arguments of Object::new (object.rs:34-41).
struct Object {
    std::string name;
    std::vector<Vec3> vertices, normals_vertices, texture_vertices;
    std::vector<IndexedMesh> meshes;
    std::vector<Texture> textures;

    // object.rs:106 through the library's loader; PNG textures are decoded by the library
    static Object load_from_file(const std::string &path) {
        draw_object *h = nullptr;
        check(draw_object_load_obj(path.c_str(), draw_image_loader_builtin, nullptr, &h));
        draw_object_desc d{};
        const int rc = draw_object_desc_of(h, &d);
        if (rc != DRAW_OK) {
            draw_object_free(h);
            check(rc);
        }
        Object o;
        o.name = d.name ? d.name : "";
        auto vec3s = [](const float *p, size_t n) {
            std::vector<Vec3> v(n);
            for (size_t i = 0; i < n; i++) v[i] = {p[3 * i], p[3 * i + 1], p[3 * i + 2]};
            return v;
        };
        o.vertices = vec3s(d.positions, d.n_positions);
        o.normals_vertices = vec3s(d.normals, d.n_normals);
        o.texture_vertices = vec3s(d.uvs, d.n_uvs);
        for (size_t i = 0; i < d.n_meshes; i++)
            o.meshes.push_back({d.meshes[i].name ? d.meshes[i].name : "",
                                std::vector<uint32_t>(d.meshes[i].triangles, d.meshes[i].triangles + 9 * d.meshes[i].n_triangles),
                                d.meshes[i].material_idx});
        auto map = [](const draw_texture_map &m) {
            TextureMap t;
            if (m.pixels) {
                t.width = m.width; t.height = m.height; t.components = m.components;
                t.img.assign(m.pixels, m.pixels + (size_t)m.width * m.height * m.components);
            }
            return t;
        };
        for (size_t i = 0; i < d.n_materials; i++) {
            const draw_material &m = d.materials[i];
            Texture t;
            t.name = m.name ? m.name : "";
            t.ka = {m.ka[0], m.ka[1], m.ka[2]}; t.kd = {m.kd[0], m.kd[1], m.kd[2]}; t.ks = {m.ks[0], m.ks[1], m.ks[2]};
            t.alpha = m.alpha;
            t.map_ka = map(m.map_ka); t.map_kd = map(m.map_kd);
            o.textures.push_back(std::move(t));
        }
        draw_object_free(h);
        return o;
    }
};

class Scene;

// VertexSimpleAttributes (canvas.rs:185-191); color is the payload of Color::Custom.
struct VertexSimpleAttributes {
    float screen_coord[2];
    float texture_coord[2];
    uint8_t color[3];
    float alpha;
};
// The arguments of Rectangle::from_coords (canvas.rs:315-330).
struct Rectangle {
    size_t x0, y0, x1, y1;
    static Rectangle from_coords(size_t x0, size_t y0, size_t x1, size_t y1) { return {x0, y0, x1, y1}; }
};
// The texture Canvas::draw_triangle samples: Texture::map_kd (RGBA8) through get_rgba_slice (scene/mod.rs:137-152),
// copied to the device once.
class DeviceTexture {
  public:
    explicit DeviceTexture(const Texture &texture) {
        const TextureMap &m = texture.map_kd;
        const draw_texture_map map{m.img.empty() ? nullptr : m.img.data(), (uint32_t)m.width, (uint32_t)m.height, (uint32_t)m.components};
        check(draw_texture_create(&map, &h_));
    }
    ~DeviceTexture() { draw_texture_destroy(h_); }
    DeviceTexture(const DeviceTexture &) = delete;
    DeviceTexture &operator=(const DeviceTexture &) = delete;
    const draw_texture *handle() const { return h_; }

  private:
    draw_texture *h_ = nullptr;
};

// Canvas (canvas.rs:353-983).
class Canvas {
  public:
    Canvas(size_t width, size_t height) { check(draw_canvas_create(width, height, &h_)); } // canvas.rs:366
    ~Canvas() { draw_canvas_destroy(h_); }
    Canvas(const Canvas &) = delete;
    Canvas &operator=(const Canvas &) = delete;
    Canvas(Canvas &&o) noexcept : h_(std::exchange(o.h_, nullptr)) {}

    void init_depth(float depth) { check(draw_canvas_init_depth(h_, depth)); }               // :403
    void apply_offset(int x, int y) { check(draw_canvas_apply_offset(h_, x, y)); }            // :382
    void resize(size_t width, size_t height) { check(draw_canvas_resize(h_, width, height)); } // :387
    void clear() { check(draw_canvas_clear(h_)); }                                            // :425
    void enable_depth_update() { check(draw_canvas_enable_depth_update(h_)); }                // :399
    void disable_depth_update() { check(draw_canvas_disable_depth_update(h_)); }              // :395
    size_t width() const { size_t w, h; check(draw_canvas_size(h_, &w, &h)); return w; }
    size_t height() const { size_t w, h; check(draw_canvas_size(h_, &w, &h)); return h; }
    static constexpr size_t pixel_bytes() { return 4; }                                       // :966
    size_t size_bytes() const { return width() * height() * pixel_bytes(); }                  // :970

    // as_bytes_slice (:974): B,G,R,pad per pixel, row 0 = top; valid until the next render / resize on this canvas
    std::pair<const uint8_t *, size_t> as_bytes_slice() {
        const uint8_t *p = nullptr;
        size_t n = 0;
        check(draw_canvas_map_host(h_, &p, &n));
        return {p, n};
    }
    const uint8_t *as_ptr() { return as_bytes_slice().first; }                                // :980
    std::vector<float> depth_frame() {                                                        // depth_frame, :413
        std::vector<float> d(width() * height());
        check(draw_canvas_read_depth(h_, d.data(), d.size()));
        return d;
    }
    float get_pixel_depth(size_t x, size_t y) { return depth_frame().at(y * width() + x); }   // :413
    // Canvas::draw_triangle (canvas.rs:435-575); clipping_rect == nullptr is None
    void draw_triangle(const VertexSimpleAttributes &a, const VertexSimpleAttributes &b, const VertexSimpleAttributes &c,
                       const DeviceTexture &texture, const Rectangle *clipping_rect = nullptr) {
        const VertexSimpleAttributes v[3] = {a, b, c};
        draw_triangles(v, 1, texture, clipping_rect);
    }
    // one draw command of Gui::render (src/app/gui.rs:382-485): n_triangles draw_triangle calls in order
    void draw_triangles(const VertexSimpleAttributes *vertices, size_t n_triangles, const DeviceTexture &texture,
                        const Rectangle *clipping_rect = nullptr) {
        std::vector<draw_vertex2d> v(3 * n_triangles);
        for (size_t i = 0; i < v.size(); i++) {
            const VertexSimpleAttributes &a = vertices[i];
            v[i] = {a.screen_coord[0], a.screen_coord[1], a.texture_coord[0], a.texture_coord[1], a.color[0], a.color[1], a.color[2], 0, a.alpha};
        }
        draw_rect r{};
        if (clipping_rect) r = {clipping_rect->x0, clipping_rect->y0, clipping_rect->x1, clipping_rect->y1};
        check(draw_canvas_draw_triangles(h_, v.data(), n_triangles, texture.handle(), clipping_rect ? &r : nullptr));
    }
    // a whole GUI frame: consecutive runs of triangles, each with its own clipping rectangle (src/app/gui.rs:389-485)
    struct Command {
        size_t n_triangles;
        const Rectangle *clipping_rect; // nullptr is None
    };
    void draw_commands(const VertexSimpleAttributes *vertices, const std::vector<Command> &commands, const DeviceTexture &texture) {
        size_t n_triangles = 0;
        std::vector<draw_command2d> table(commands.size());
        for (size_t k = 0; k < commands.size(); k++) {
            table[k].n_triangles = commands[k].n_triangles;
            table[k].has_clip = commands[k].clipping_rect ? 1 : 0;
            if (const Rectangle *r = commands[k].clipping_rect) table[k].clip = {r->x0, r->y0, r->x1, r->y1};
            n_triangles += commands[k].n_triangles;
        }
        std::vector<draw_vertex2d> v(3 * n_triangles);
        for (size_t i = 0; i < v.size(); i++) {
            const VertexSimpleAttributes &a = vertices[i];
            v[i] = {a.screen_coord[0], a.screen_coord[1], a.texture_coord[0], a.texture_coord[1], a.color[0], a.color[1], a.color[2], 0, a.alpha};
        }
        check(draw_canvas_draw_commands(h_, v.data(), n_triangles, table.data(), table.size(), texture.handle()));
    }
    void export_jpeg(const std::string &path) { check(draw_canvas_export_jpeg(h_, path.c_str())); } // app/mod.rs:316, Jpeg
    void export_png(const std::string &path) { check(draw_canvas_export_png(h_, path.c_str())); } // app/mod.rs:316
    void sync() { check(draw_canvas_sync(h_)); }
    draw_canvas *handle() const { return h_; }

  private:
    draw_canvas *h_ = nullptr;
};

// Scene (scene/mod.rs:749-1249).
class Scene {
  public:
    Scene(size_t width, size_t height) { check(draw_scene_create(width, height, &h_)); } // :760
    ~Scene() { draw_scene_destroy(h_); }
    Scene(const Scene &) = delete;
    Scene &operator=(const Scene &) = delete;

    ObjectInfo add_obj(const Object &o) { // :788 (the Object is copied; the reference moves it in)
        std::vector<draw_mesh> meshes;
        for (const IndexedMesh &m : o.meshes) meshes.push_back({m.name.c_str(), m.triangles.data(), m.triangles.size() / 9, m.texture_idx});
        auto map = [](const TextureMap &t) {
            return draw_texture_map{t.img.empty() ? nullptr : t.img.data(), t.width, t.height, t.components};
        };
        std::vector<draw_material> mats;
        for (const Texture &t : o.textures)
            mats.push_back({t.name.c_str(), {t.ka[0], t.ka[1], t.ka[2]}, {t.kd[0], t.kd[1], t.kd[2]}, {t.ks[0], t.ks[1], t.ks[2]},
                            t.alpha, map(t.map_ka), map(t.map_kd)});
        draw_object_desc d{};
        d.name = o.name.c_str();
        d.positions = o.vertices.empty() ? nullptr : o.vertices[0].data(); d.n_positions = o.vertices.size();
        d.normals = o.normals_vertices.empty() ? nullptr : o.normals_vertices[0].data(); d.n_normals = o.normals_vertices.size();
        d.uvs = o.texture_vertices.empty() ? nullptr : o.texture_vertices[0].data(); d.n_uvs = o.texture_vertices.size();
        d.meshes = meshes.data(); d.n_meshes = meshes.size();
        d.materials = mats.data(); d.n_materials = mats.size();
        uint32_t id = 0;
        check(draw_scene_add_object(h_, &d, &id));
        return {id};
    }
    // scene.camera = Camera::new(pos, dir, ratio) (:297; ratio = scene width / height, :768)
    void set_camera(const Vec3 &pos, const Vec3 &dir) { check(draw_scene_set_camera(h_, pos.data(), dir.data())); }
    std::pair<Vec3, Vec3> camera() const {
        Vec3 p{}, d{};
        check(draw_scene_get_camera(h_, p.data(), d.data()));
        return {p, d};
    }
    void camera_set_pos(const Vec3 &pos) { check(draw_scene_set_camera_pos(h_, pos.data())); }       // Camera::set_pos :377
    void camera_move_up(float d) { check(draw_scene_camera_move(h_, DRAW_CAMERA_UP, d)); }             // :381-405
    void camera_move_down(float d) { check(draw_scene_camera_move(h_, DRAW_CAMERA_DOWN, d)); }
    void camera_move_left(float d) { check(draw_scene_camera_move(h_, DRAW_CAMERA_LEFT, d)); }
    void camera_move_right(float d) { check(draw_scene_camera_move(h_, DRAW_CAMERA_RIGHT, d)); }
    void camera_move_foward(float d) { check(draw_scene_camera_move(h_, DRAW_CAMERA_FOWARD, d)); }     // sic, like the reference
    void camera_move_backward(float d) { check(draw_scene_camera_move(h_, DRAW_CAMERA_BACKWARD, d)); }
    void move_camera_direction(int dx, int dy) { check(draw_scene_move_camera_direction(h_, dx, dy)); } // :803
    void render(Canvas &canvas) { check(draw_scene_render(h_, canvas.handle())); }                     // :901
    draw_scene *handle() const { return h_; }

  private:
    draw_scene *h_ = nullptr;
};

} // namespace draw
