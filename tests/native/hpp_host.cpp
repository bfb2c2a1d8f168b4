// Exercises include/draw_b200.hpp (the C++ mirror of the reference's Rust host API).
//   hpp_host cpu <tmpdir>  : host-only entry points + the error behaviour without a device
//   hpp_host gpu           : renders a two-object scene like the reference's Application would and prints
//                            FNV-1a checksums of the BGRA frame and the depth buffer (compared by the test with
//                            the Python binding's, which is parity-tested against the oracle)
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>

#include "draw_b200.hpp"

static uint64_t fnv(const void *p, size_t n) {
    const uint8_t *b = static_cast<const uint8_t *>(p);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}

static draw::Object triangle(float z, draw::Vec3 kd, float alpha) {
    draw::Object o;
    o.name = "tri";
    o.vertices = {{-60.f, -50.f, z}, {60.f, -50.f, z}, {0.f, 60.f, z}};
    o.normals_vertices = {{0.f, 0.f, 1.f}, {0.f, 0.f, 1.f}, {0.f, 0.f, 1.f}};
    o.texture_vertices = {{0.f, 0.f, 0.f}};
    o.meshes.push_back({"m", {0, 1, 2, 0, 0, 0, 0, 1, 2}, 0});
    draw::Texture t;
    t.kd = kd;
    t.alpha = alpha;
    o.textures.push_back(t);
    return o;
}

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    if (mode == "cpu") {
        const std::string dir = argc > 2 ? argv[2] : ".";
        // TextureMap::load_from_file round trip through the library's PNG writer
        std::vector<uint8_t> px(5 * 3 * 4);
        for (size_t i = 0; i < px.size(); i++) px[i] = (uint8_t)(i * 37 + 11);
        const std::string png = dir + "/t.png";
        draw::check(draw_image_write_png(png.c_str(), px.data(), 5, 3, 4));
        const draw::TextureMap m = draw::TextureMap::load_from_file(png);
        if (m.width != 5 || m.height != 3 || m.components != 4 || m.img != px) { std::puts("FAIL texture round trip"); return 1; }
        // Object::load_from_file on a minimal OBJ (object.rs:106): rescaled to radius 100, normals generated
        { std::ofstream f(dir + "/q.obj"); f << "v 0 0 0\nv 2 0 0\nv 0 1 0\nf 1 2 3\n"; }
        const draw::Object o = draw::Object::load_from_file(dir + "/q.obj");
        if (o.vertices.size() != 3 || o.meshes.size() != 1 || o.meshes[0].triangles.size() != 9 || o.vertices[1][0] != 100.0f) {
            std::puts("FAIL load_from_file"); return 1;
        }
        // a missing file is an error, not a crash (the reference panics: object.rs:131 expect)
        try { draw::Object::load_from_file(dir + "/missing.obj"); std::puts("FAIL no error"); return 1; } catch (const draw::Error &) {}
        // without a CUDA device the compute side fails loudly (no CPU fallback)
        int n_dev = 0;
        draw_device_count(&n_dev);
        if (n_dev == 0) {
            try { draw::Scene s(64, 48); std::puts("FAIL scene without device"); return 1; }
            catch (const draw::Error &e) {
                if (e.status != DRAW_ERR_NO_DEVICE) { std::printf("FAIL status %d\n", e.status); return 1; }
            }
        }
        std::puts("OK cpu");
        return 0;
    }
    // the reference's driver sequence (app/mod.rs:58-84,196-202) with its own names
    draw::Scene scene(320, 240);
    draw::Canvas canvas(320, 240);
    canvas.init_depth(100000.0f);
    canvas.apply_offset(0, 0);
    scene.add_obj(triangle(0.f, {1.f, 1.f, 0.f}, 1.0f));
    scene.add_obj(triangle(20.f, {0.f, 0.f, 1.f}, 0.5f));
    scene.set_camera({20.f, 5.f, 120.f}, {-0.2f, 0.f, -1.f});
    scene.camera_move_up(3.0f);
    scene.render(canvas);
    const auto bytes = canvas.as_bytes_slice();
    const std::vector<float> depth = canvas.depth_frame();
    std::printf("frame %016llx depth %016llx bytes %zu\n", (unsigned long long)fnv(bytes.first, bytes.second),
                (unsigned long long)fnv(depth.data(), depth.size() * 4), bytes.second);
    return 0;
}
