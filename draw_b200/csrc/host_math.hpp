// host_math.hpp — host-side per-frame uniforms in the reference's exact float32 operation order.
//
// The renderer's per-frame constants (transformation matrix, six view planes, camera basis) are
// a few hundred flops; they are computed on the host and passed to the kernels by value.  They
// must match the reference bit for bit, so every expression keeps the evaluation order of
// mororo18/draw src/renderer/linalg.rs and scene/mod.rs (cited per function); this file is
// compiled with -ffp-contract=off and never with -ffast-math.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

namespace drawb200 {

struct f3 {
    float x, y, z;
};

inline f3 add(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }   // linalg.rs:233
inline f3 sub(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }   // linalg.rs:247
inline f3 scale(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }    // linalg.rs:213
inline f3 divide(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }   // linalg.rs:223
inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } // linalg.rs:182
inline float length(f3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); } // linalg.rs:167
inline f3 unit(f3 a) { return divide(a, length(a)); }                     // linalg.rs:173
inline f3 cross3(f3 a, f3 b) {                                            // linalg.rs:186
    return {(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
}

struct m4 {
    float v[16]; // row-major
    float &at(int r, int c) { return v[4 * r + c]; }
    float at(int r, int c) const { return v[4 * r + c]; }
};

// linalg.rs:328-344 — each element accumulates from 0.0 over k = 0..3 in order.
inline m4 mul(const m4 &a, const m4 &b) {
    m4 c;
    for (int r = 0; r < 4; r++)
        for (int col = 0; col < 4; col++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s = s + a.at(r, k) * b.at(k, col);
            c.at(r, col) = s;
        }
    return c;
}

// linalg.rs:346-360 with Vec3::as_vec4 (w = 1) and Vec4::vec3_over_w (linalg.rs:89,163).
inline f3 transform_point_over_w(const m4 &m, f3 p) {
    const float in[4] = {p.x, p.y, p.z, 1.0f};
    float out[4];
    for (int r = 0; r < 4; r++) {
        float s = 0.0f;
        for (int k = 0; k < 4; k++) s = s + m.at(r, k) * in[k];
        out[r] = s;
    }
    return divide(f3{out[0], out[1], out[2]}, out[3]);
}

struct plane4 {
    float nx, ny, nz, k;
};

// ViewPlane::new, scene/mod.rs:603-632.
inline plane4 make_plane(f3 p0, f3 p1, f3 p2, f3 inside) {
    const f3 e0 = sub(p1, p0), e1 = sub(p2, p1);
    f3 n = cross3(e0, e1);
    float k = -dot3(n, p0);
    if (dot3(n, inside) + k < 0.0f) {
        n = cross3(e1, e0);
        k = -dot3(n, p0);
    }
    return {n.x, n.y, n.z, k};
}

// Camera, scene/mod.rs:282-294.
struct CameraState {
    f3 position, direction, up;
    float top, bottom, right, left, near_z, far_z;
    f3 u{0, 0, 0}, v{0, 0, 0}, w{0, 0, 0};

    // Camera::new, scene/mod.rs:297-357
    static CameraState make(f3 pos, f3 dir, float ratio) {
        CameraState c;
        const float near_z = -10.0f;
        const float far_z = near_z - 500.0f;
        const float fov_deg = 135.0f;
        const float pi_f = 3.14159265358979323846f;
        const float fov_rad = fov_deg * (pi_f / 180.0f); // f32::to_radians
        c.right = std::fabs(near_z) * std::tan(fov_rad / 2.0f);
        c.left = -c.right;
        c.top = (1.0f / ratio) * c.right;
        c.bottom = -c.top;
        c.position = pos;
        c.direction = unit(dir);
        c.up = {0.0f, 1.0f, 0.0f};
        c.near_z = near_z;
        c.far_z = far_z;
        return c;
    }

    // Camera::update_basis, scene/mod.rs:438-451
    void update_basis() {
        const f3 back = scale(divide(direction, length(direction)), -1.0f);
        const f3 side = cross3(up, back);
        const f3 side_n = divide(side, length(side));
        const f3 upv = cross3(back, side_n);
        u = unit(side_n);
        v = unit(upv);
        w = unit(back);
    }

    // Camera::get_basis_matrix, scene/mod.rs:415-428
    m4 basis() const {
        return {{u.x, v.x, w.x, 0.0f, u.y, v.y, w.y, 0.0f, u.z, v.z, w.z, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f}};
    }

    // Camera::gen_matrix, scene/mod.rs:453-479 : transpose(basis) * translate(-pos)
    m4 view_matrix() {
        const m4 tr = {{1.0f, 0.0f, 0.0f, -position.x, 0.0f, 1.0f, 0.0f, -position.y,
                        0.0f, 0.0f, 1.0f, -position.z, 0.0f, 0.0f, 0.0f, 1.0f}};
        update_basis();
        const m4 b = basis();
        m4 bt;
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) bt.at(r, c) = b.at(c, r);
        return mul(bt, tr);
    }

    // Camera::gen_view_planes, scene/mod.rs:481-593 ; out = near, far, right, left, top, bottom
    void view_planes(plane4 out[6]) {
        update_basis();
        const m4 b = basis();
        auto to_world = [&](float x, float y, float z) {
            return add(transform_point_over_w(b, f3{x, y, z}), position);
        };
        const float n = near_z, f = far_z, r = right, l = left, t = top, bo = bottom;
        const f3 ur = to_world(r, t, n), ul = to_world(l, t, n);
        const f3 lr = to_world(r, bo, n), ll = to_world(l, bo, n);
        const float xc = (l + r) / 2.0f, yc = (bo + t) / 2.0f;
        const f3 far_up = to_world(xc, (f * t) / n, f);
        const f3 far_lo = to_world(xc, (f * bo) / n, f);
        const f3 far_ri = to_world((f * r) / n, yc, f);
        const f3 far_le = to_world((f * l) / n, yc, f);
        const f3 inside = divide(add(ur, far_lo), 2.0f);
        out[0] = make_plane(ur, lr, ll, inside);           // near   :533
        out[1] = make_plane(far_le, far_ri, far_up, inside); // far    :542
        out[2] = make_plane(far_ri, ur, lr, inside);       // right  :554
        out[3] = make_plane(far_le, ll, ul, inside);       // left   :563
        out[4] = make_plane(far_up, ul, ur, inside);       // top    :572
        out[5] = make_plane(far_lo, ll, lr, inside);       // bottom :581
    }
};

// Scene::gen_transformation_matrix, scene/mod.rs:817-899 : ((viewport * orth) * persp) * view
inline m4 transformation_matrix(CameraState &cam, size_t scene_w, size_t scene_h) {
    const float nx = (float)scene_w, ny = (float)scene_h;
    const float n = cam.near_z, f = cam.far_z;
    const float r = cam.right, l = cam.left, t = cam.top, b = cam.bottom;
    const m4 view = cam.view_matrix();
    const m4 persp = {{n, 0.0f, 0.0f, 0.0f, 0.0f, n, 0.0f, 0.0f, 0.0f, 0.0f, (n + f), -(n * f), 0.0f, 0.0f, 1.0f, 0.0f}};
    const m4 orth = {{2.0f / (r - l), 0.0f, 0.0f, -(r + l) / (r - l),
                      0.0f, 2.0f / (t - b), 0.0f, -(t + b) / (t - b),
                      0.0f, 0.0f, 2.0f / (n - f), -(n + f) / (n - f),
                      0.0f, 0.0f, 0.0f, 1.0f}};
    const m4 viewport = {{nx / 2.0f, 0.0f, 0.0f, (nx - 1.0f) / 2.0f,
                          0.0f, ny / 2.0f, 0.0f, (ny - 1.0f) / 2.0f,
                          0.0f, 0.0f, 1.0f, 0.0f,
                          0.0f, 0.0f, 0.0f, 1.0f}};
    return mul(mul(mul(viewport, orth), persp), view);
}

// f32::total_cmp as an order-preserving integer key (used by the painter sort, :1114).
inline int32_t total_order_key(float x) {
    int32_t i;
    std::memcpy(&i, &x, 4);
    return i ^ (int32_t)(((uint32_t)(i >> 31)) >> 1);
}

} // namespace drawb200
