/*
 * oracle.cpp — CPU restatement of the reference software rasterizer.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED by the reference's own tests
 * (it has none); pinned by oracle/np_oracle.py and the code-derived KATs in tests/.
 *
 * Every function cites the reference lines it restates (paths relative to the reference
 * root, mororo18/draw).  The arithmetic contract is IEEE-754 binary32, round-to-nearest,
 * NO fused multiply-add and NO reassociation: build with -ffp-contract=off and without
 * -ffast-math (oracle/Makefile does).  Expressions keep the reference's evaluation order,
 * including the "0.0 + ..." left over from accumulator loops, because (-0.0)+0.0 == +0.0.
 *
 * Rust semantics that differ from C++ and are restated explicitly:
 *   - `f32 as usize` / `f32 as u8` saturate and map NaN to 0   -> sat_usize / sat_u8
 *   - u8 `+` wraps in release builds                            -> uint8_t arithmetic
 *   - slice::sort_by is stable                                  -> std::stable_sort
 *   - f32::total_cmp                                            -> total_key
 */
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace {

constexpr float EPS = 0.0000001f; /* linalg.rs:6 */

/* ------------------------------------------------------------------ linalg.rs */

struct Vec2 {
    float x, y;
};
/* linalg.rs:29-43 : Sub is Add of the negation (bit-identical to a-b in IEEE). */
inline Vec2 operator+(Vec2 a, Vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline Vec2 operator-(Vec2 a, Vec2 b) { return a + Vec2{-b.x, -b.y}; }
inline Vec2 operator*(Vec2 a, float s) { return {a.x * s, a.y * s}; } /* linalg.rs:45-54 */
inline Vec2 operator/(Vec2 a, float s) { return {a.x / s, a.y / s}; } /* linalg.rs:56-65 */

struct Vec3 {
    float x, y, z;
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; } /* :233 */
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; } /* :247 */
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }       /* :213 */
inline Vec3 operator/(Vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }       /* :223 */
/* linalg.rs:167-171, powi(2) == x*x */
inline float norm(Vec3 a) {
    float sum = a.x * a.x + a.y * a.y + a.z * a.z;
    return std::sqrt(sum);
}
inline Vec3 normalized(Vec3 a) { return a / norm(a); } /* linalg.rs:173-175 */
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* :182 */
/* linalg.rs:186-200 */
inline Vec3 cross(Vec3 a, Vec3 b) {
    return {(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
}
inline float dist(Vec3 a, Vec3 b) { return norm(a - b); } /* linalg.rs:177-180 */
/* canvas.rs:13-17 */
inline Vec3 color_multiply(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }

struct Vec4 {
    float a[4];
};
struct Mat4 {
    float a[4][4];
};
/* linalg.rs:328-344 : c starts at zero and accumulates k = 0..3 in order. */
inline Mat4 operator*(const Mat4 &l, const Mat4 &r) {
    Mat4 c;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float acc = 0.0f;
            for (int k = 0; k < 4; k++) acc += l.a[i][k] * r.a[k][j];
            c.a[i][j] = acc;
        }
    return c;
}
/* linalg.rs:346-360 */
inline Vec4 operator*(const Mat4 &m, const Vec4 &v) {
    Vec4 out;
    for (int i = 0; i < 4; i++) {
        float acc = 0.0f;
        for (int j = 0; j < 4; j++) acc += m.a[i][j] * v.a[j];
        out.a[i] = acc;
    }
    return out;
}
/* linalg.rs:275-287 */
inline Mat4 transposed(const Mat4 &m) {
    Mat4 t;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) t.a[i][j] = m.a[j][i];
    return t;
}
inline Vec4 as_vec4(Vec3 v) { return {{v.x, v.y, v.z, 1.0f}}; }              /* linalg.rs:163 */
inline Vec3 vec3_over_w(const Vec4 &v) { return Vec3{v.a[0], v.a[1], v.a[2]} / v.a[3]; } /* :89 */

/* ------------------------------------------------------------ Rust cast semantics */

inline uint64_t sat_usize(float v) {
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 18446744073709551616.0f) return std::numeric_limits<uint64_t>::max();
    return (uint64_t)v;
}
inline uint8_t sat_u8(float v) {
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
/* f32::total_cmp key: order-isomorphic signed integer */
inline int32_t total_key(float v) {
    int32_t i;
    std::memcpy(&i, &v, 4);
    i ^= (int32_t)(((uint32_t)(i >> 31)) >> 1);
    return i;
}

/* ------------------------------------------------------------------ canvas.rs */

/* canvas.rs:51-59 ; memory order b,g,r,padd */
struct Pixel {
    uint8_t b, g, r, padd;
};
inline Pixel pixel_new(uint8_t r, uint8_t g, uint8_t b) { return {b, g, r, 255}; } /* :63-70 */
/* canvas.rs:154-169 */
inline Pixel pixel_mul(Pixel p, float rhs) {
    Pixel o;
    o.r = sat_u8((float)p.r * rhs);
    o.g = sat_u8((float)p.g * rhs);
    o.b = sat_u8((float)p.b * rhs);
    o.padd = 0;
    return o;
}
/* canvas.rs:136-152 (u8 add; wraps in a release build) */
inline Pixel pixel_add(Pixel a, Pixel b) {
    Pixel o;
    o.r = (uint8_t)(a.r + b.r);
    o.g = (uint8_t)(a.g + b.g);
    o.b = (uint8_t)(a.b + b.b);
    o.padd = 0;
    return o;
}

/* canvas.rs:193-203 */
struct VertexAttributes {
    Vec3 normal, light, halfway, texture_coord;
    Vec2 screen_coord;
    float depth;
};
/* canvas.rs:242-291 */
inline VertexAttributes operator+(const VertexAttributes &a, const VertexAttributes &b) {
    return {a.normal + b.normal, a.light + b.light, a.halfway + b.halfway,
            a.texture_coord + b.texture_coord, a.screen_coord + b.screen_coord, a.depth + b.depth};
}
inline VertexAttributes operator-(const VertexAttributes &a, const VertexAttributes &b) {
    return {a.normal - b.normal, a.light - b.light, a.halfway - b.halfway,
            a.texture_coord - b.texture_coord, a.screen_coord - b.screen_coord, a.depth - b.depth};
}
inline VertexAttributes operator*(const VertexAttributes &a, float s) {
    return {a.normal * s, a.light * s, a.halfway * s, a.texture_coord * s, a.screen_coord * s,
            a.depth * s};
}

/* scene/mod.rs:102-110 */
struct TextureMap {
    std::vector<uint8_t> img;
    size_t width = 0, height = 0, components = 0;
    float f_width = 0, f_height = 0;
    /* scene/mod.rs:154-168.  Deviation 6: indices are clamped into the map so
     * an out-of-range uv cannot read outside the image; never triggers for uv in [0,1]. */
    inline const uint8_t *get_rgb_slice(float u, float v) const {
        uint64_t u_idx = sat_usize(std::floor(u * (f_width - 1.0f)));
        uint64_t v_raw = sat_usize(std::floor(v * (f_height - 1.0f)));
        if (u_idx > width - 1) u_idx = width - 1;
        if (v_raw > height - 1) v_raw = height - 1;
        uint64_t v_idx = height - 1 - v_raw;
        return &img[(v_idx * width + u_idx) * components];
    }
};
TextureMap texture_map_new(const uint8_t *data, size_t w, size_t h, size_t comp) {
    TextureMap t;
    t.img.assign(data, data + w * h * comp);
    t.width = w;
    t.height = h;
    t.components = comp;
    t.f_width = (float)w;
    t.f_height = (float)h;
    return t;
}
TextureMap texture_map_default() { /* scene/mod.rs:128-135 */
    const uint8_t white[3] = {255, 255, 255};
    return texture_map_new(white, 1, 1, 3);
}
/* scene/mod.rs:206-216 */
struct Texture {
    Vec3 ka, kd, ks;
    float alpha;
    TextureMap map_ka, map_kd;
};
Texture texture_default() { /* scene/mod.rs:237-252 */
    return {{0.9f, 0.9f, 0.9f}, {0.4f, 0.4f, 0.4f}, {0.5f, 0.5f, 0.5f}, 1.0f,
            texture_map_default(), texture_map_default()};
}

constexpr uint32_t NO_WINNER = 0xFFFFFFFFu;

} // namespace

/* canvas.rs:353-363 */
struct orc_canvas {
    std::vector<Pixel> frame;
    size_t width = 0, height = 0;
    Vec2 offset{0.0f, 0.0f};
    bool depth_update_enabled = false;
    std::vector<float> depth_frame;
    float depth_max = 0.0f;
    /* bookkeeping, not in the reference */
    std::vector<uint32_t> winner;
    uint32_t cur_id = 0;
    orc_stats *stats = nullptr;

    void init_depth(float depth) { /* canvas.rs:403-411 */
        depth_max = depth;
        depth_frame.assign(frame.size(), depth);
    }
    void clear() { /* canvas.rs:425-433 ; azul_bb = Pixel::new(155,186,255) */
        std::fill(frame.begin(), frame.end(), pixel_new(155, 186, 255));
        if (!depth_frame.empty()) init_depth(depth_max);
        std::fill(winner.begin(), winner.end(), NO_WINNER);
    }
    /* canvas.rs:906-930 with get/draw_pixel_coord (932-960) and get/set_pixel_depth (413-423) */
    inline void draw_pixel_coord_with_depth(size_t x, size_t y, Pixel color, float opacity,
                                            float depth) {
        size_t y_inv = height - y - 1;
        size_t off = width * y_inv + x;
        Pixel new_color;
        if (opacity < 1.0f) {
            Pixel bg = frame[off];
            new_color = pixel_add(pixel_mul(bg, 1.0f - opacity), pixel_mul(color, opacity));
        } else {
            new_color = color;
        }
        if (depth < depth_frame[width * y + x]) {
            frame[off] = new_color;
            winner[width * y + x] = cur_id;
            if (stats) stats->written_frags++;
            if (depth_update_enabled) depth_frame[width * y + x] = depth;
        }
    }
    template <bool STATS>
    void draw_triangle_with_attributes(const VertexAttributes &a_attr,
                                       const VertexAttributes &b_attr,
                                       const VertexAttributes &c_attr, const Texture &texture);
};

namespace {

/* canvas.rs:896-904 */
inline Vec2 pos_map_center(Vec2 pos) { return {std::floor(pos.x + 0.5f), std::floor(pos.y + 0.5f)}; }

/* canvas.rs:293-351 ; from_coords normalises, clip falls back to (0,0) when empty */
struct Rect {
    uint64_t x, y, w, h;
    uint64_t x_max() const { return x + w; }
    uint64_t y_max() const { return y + h; }
};
inline Rect rect_from_coords(uint64_t x0, uint64_t y0, uint64_t x1, uint64_t y1) {
    uint64_t x_min = std::min(x0, x1), y_min = std::min(y0, y1);
    uint64_t x_max = std::max(x0, x1), y_max = std::max(y0, y1);
    return {x_min, y_min, x_max - x_min, y_max - y_min};
}
inline Rect rect_clip(const Rect &a, const Rect &b) {
    uint64_t x_min = std::max(a.x, b.x), y_min = std::max(a.y, b.y);
    uint64_t x_max = std::min(a.x + a.w, b.x + b.w), y_max = std::min(a.y + a.h, b.y + b.h);
    if (x_min > x_max) x_min = x_max = 0;
    if (y_min > y_max) y_min = y_max = 0;
    return rect_from_coords(x_min, y_min, x_max, y_max);
}
/* canvas.rs:618-638 */
inline float min3(float x, float y, float z) {
    float ret = std::numeric_limits<float>::infinity();
    const float v[3] = {x, y, z};
    for (float e : v)
        if (e < ret) ret = e;
    return ret;
}
inline float max3(float x, float y, float z) {
    float ret = -std::numeric_limits<float>::infinity();
    const float v[3] = {x, y, z};
    for (float e : v)
        if (e > ret) ret = e;
    return ret;
}

} // namespace

/* canvas.rs:577-750 */
template <bool STATS>
void orc_canvas::draw_triangle_with_attributes(const VertexAttributes &a_attr,
                                               const VertexAttributes &b_attr,
                                               const VertexAttributes &c_attr,
                                               const Texture &texture) {
    const Vec2 a_center = pos_map_center(a_attr.screen_coord - offset); /* :585-587 */
    const Vec2 b_center = pos_map_center(b_attr.screen_coord - offset);
    const Vec2 c_center = pos_map_center(c_attr.screen_coord - offset);

    const float a_depth = a_attr.depth, b_depth = b_attr.depth, c_depth = c_attr.depth;

    /* canvas.rs:597-616 : ((cx*x + cy*y) + k1) - k2 */
    auto f_ab = [&](float x, float y) -> float {
        return (a_center.y - b_center.y) * x + (b_center.x - a_center.x) * y +
               (a_center.x * b_center.y) - (b_center.x * a_center.y);
    };
    auto f_bc = [&](float x, float y) -> float {
        return (b_center.y - c_center.y) * x + (c_center.x - b_center.x) * y +
               (b_center.x * c_center.y) - (c_center.x * b_center.y);
    };
    auto f_ca = [&](float x, float y) -> float {
        return (c_center.y - a_center.y) * x + (a_center.x - c_center.x) * y +
               (c_center.x * a_center.y) - (a_center.x * c_center.y);
    };

    /* canvas.rs:640-658 */
    uint64_t x_min = sat_usize(min3(a_center.x, b_center.x, c_center.x));
    uint64_t y_min = sat_usize(min3(a_center.y, b_center.y, c_center.y));
    uint64_t x_max = sat_usize(max3(a_center.x, b_center.x, c_center.x));
    uint64_t y_max = sat_usize(max3(a_center.y, b_center.y, c_center.y));

    Rect drawable = rect_from_coords(x_min, y_min, x_max, y_max);
    const Rect screen = rect_from_coords(0, 0, width - 1, height - 1);
    drawable = rect_clip(drawable, screen);
    const Rect valid = rect_clip(screen, drawable);
    x_min = valid.x;
    y_min = valid.y;
    x_max = valid.x_max();
    y_max = valid.y_max();

    /* canvas.rs:660-666 */
    const float f_alpha = f_bc(a_center.x, a_center.y);
    const float f_beta = f_ca(b_center.x, b_center.y);
    const float f_gama = f_ab(c_center.x, c_center.y);
    const float f_alpha_outside = f_bc(-1.0f, -1.0f);
    const float f_beta_outside = f_ca(-1.0f, -1.0f);
    const float f_gama_outside = f_ab(-1.0f, -1.0f);

    for (uint64_t y = y_min; y <= y_max; y++) { /* :668 */
        const float y_f32 = (float)y;
        for (uint64_t x = x_min; x <= x_max; x++) {
            const float x_f32 = (float)x;
            if (STATS) stats->bbox_pixels++;

            const float alpha = f_bc(x_f32, y_f32) / f_alpha; /* :673-675 */
            const float beta = f_ca(x_f32, y_f32) / f_beta;
            const float gama = f_ab(x_f32, y_f32) / f_gama;

            if (alpha >= 0.0f && beta >= 0.0f && gama >= 0.0f) {
                if ((alpha > 0.0f || f_alpha * f_alpha_outside > 0.0f) &&
                    (beta > 0.0f || f_beta * f_beta_outside > 0.0f) &&
                    (gama > 0.0f || f_gama * f_gama_outside > 0.0f)) {
                    if (STATS) stats->covered_frags++;
                    const float pixel_depth =
                        (alpha * a_depth) + (beta * b_depth) + (gama * c_depth); /* :682 */

                    const Vec3 uv = (a_attr.texture_coord * alpha) +
                                    (b_attr.texture_coord * beta) +
                                    (c_attr.texture_coord * gama); /* :685-687 */

                    const uint8_t *d = texture.map_kd.get_rgb_slice(uv.x, uv.y); /* :689-695 */
                    const uint8_t *a = texture.map_ka.get_rgb_slice(uv.x, uv.y);
                    /* Pixel::normalized_as_vec3, canvas.rs:81-87 */
                    const Vec3 diffuse_color{(float)d[0] / 255.0f, (float)d[1] / 255.0f,
                                             (float)d[2] / 255.0f};
                    const Vec3 ambient_color{(float)a[0] / 255.0f, (float)a[1] / 255.0f,
                                             (float)a[2] / 255.0f};

                    const Vec3 pixel_normal =
                        (a_attr.normal * alpha) + (b_attr.normal * beta) + (c_attr.normal * gama);
                    const Vec3 pixel_light =
                        (a_attr.light * alpha) + (b_attr.light * beta) + (c_attr.light * gama);
                    const Vec3 pixel_halfway = (a_attr.halfway * alpha) +
                                               (b_attr.halfway * beta) +
                                               (c_attr.halfway * gama); /* :713-722 */

                    const Vec3 c_l = texture.ks;                                /* :732 */
                    const Vec3 c_r = color_multiply(diffuse_color, texture.kd); /* :733 */
                    const Vec3 c_a = color_multiply(ambient_color, texture.ka); /* :734 */

                    /* 0.0_f32.max(x): NaN -> 0 (maxNum) */
                    const float ln = dot(pixel_light, pixel_normal);
                    const float ln_pos = (ln > 0.0f) ? ln : 0.0f;
                    const float hn = dot(pixel_halfway, pixel_normal);
                    const float hn_pow = hn * hn; /* powi(2), :724,739 */
                    const Vec3 color_normalized =
                        color_multiply(c_r, c_a + c_l * (1.0f - ln_pos)) + c_l * hn_pow; /* :736 */

                    /* Pixel::from_normalized_vec3, canvas.rs:89-92 */
                    const Vec3 scaled = color_normalized * 255.0f;
                    const Pixel color = pixel_new(sat_u8(scaled.x), sat_u8(scaled.y), sat_u8(scaled.z));

                    draw_pixel_coord_with_depth(x, y, color, texture.alpha, pixel_depth); /* :745 */
                }
            }
        }
    }
}

namespace {

/* scene/mod.rs:13-16 */
struct Triangle {
    Vec3 vertices[3];
    VertexAttributes vertices_attr[3];
};

/* scene/mod.rs:596-747 */
struct ViewPlane {
    Vec3 normal;
    float k;
    inline float func(Vec3 p) const { return dot(normal, p) + k; } /* :634 */
    /* :641-660 */
    bool at_least_partially_visible(const Triangle &tri) const {
        float f_a = func(tri.vertices[0]), f_b = func(tri.vertices[1]), f_c = func(tri.vertices[2]);
        if (f_a > 0.0f && f_b > 0.0f && f_c > 0.0f) return true;
        if (f_a <= 0.0f && f_b <= 0.0f && f_c <= 0.0f) return false;
        return true;
    }
    /* :662-746 */
    size_t clip(const Triangle &tri, Triangle *ret) const {
        Vec3 a_vertex = tri.vertices[0], b_vertex = tri.vertices[1], c_vertex = tri.vertices[2];
        VertexAttributes a_attr = tri.vertices_attr[0], b_attr = tri.vertices_attr[1],
                         c_attr = tri.vertices_attr[2];
        float f_a = func(a_vertex), f_b = func(b_vertex), f_c = func(c_vertex);

        if (f_a > 0.0f && f_b > 0.0f && f_c > 0.0f) {
            ret[0] = tri;
            return 1;
        } else if (f_a <= 0.0f && f_b <= 0.0f && f_c <= 0.0f) {
            return 0;
        }
        if (f_a * f_c >= 0.0f) { /* :691-700 */
            std::swap(f_b, f_c);
            std::swap(b_vertex, c_vertex);
            std::swap(b_attr, c_attr);
            std::swap(f_a, f_b);
            std::swap(a_vertex, b_vertex);
            std::swap(a_attr, b_attr);
        } else if (f_b * f_c >= 0.0f) { /* :701-711 */
            std::swap(f_a, f_c);
            std::swap(a_vertex, c_vertex);
            std::swap(a_attr, c_attr);
            std::swap(f_a, f_b);
            std::swap(a_vertex, b_vertex);
            std::swap(a_attr, b_attr);
        }
        /* :715-720 */
        const float t_a = func(a_vertex) / dot(normal, a_vertex - c_vertex) - EPS;
        const Vec3 new_vertex_a = a_vertex + (c_vertex - a_vertex) * t_a;
        const VertexAttributes new_a_attr = a_attr + (c_attr - a_attr) * t_a;
        const float t_b = func(b_vertex) / dot(normal, b_vertex - c_vertex) - EPS;
        const Vec3 new_vertex_b = b_vertex + (c_vertex - b_vertex) * t_b;
        const VertexAttributes new_b_attr = b_attr + (c_attr - b_attr) * t_b;

        if (f_c <= 0.0f) { /* :723-736 */
            ret[0] = Triangle{{a_vertex, new_vertex_a, new_vertex_b}, {a_attr, new_a_attr, new_b_attr}};
            ret[1] = Triangle{{a_vertex, b_vertex, new_vertex_b}, {a_attr, b_attr, new_b_attr}};
            return 2;
        } else { /* :737-745 */
            ret[0] = Triangle{{c_vertex, new_vertex_a, new_vertex_b}, {c_attr, new_a_attr, new_b_attr}};
            return 1;
        }
    }
};
/* scene/mod.rs:603-632 */
ViewPlane view_plane_new(Vec3 a_point, Vec3 b_point, Vec3 c_point, Vec3 visible_point) {
    Vec3 p_vec = b_point - a_point;
    Vec3 q_vec = c_point - b_point;
    Vec3 normal = cross(p_vec, q_vec);
    float k = -dot(normal, a_point);
    float test_value = dot(normal, visible_point) + k;
    if (test_value < 0.0f) {
        normal = cross(q_vec, p_vec);
        k = -dot(normal, a_point);
    }
    return {normal, k};
}

struct Planes {
    ViewPlane depth[2];   /* near, far */
    ViewPlane lateral[4]; /* right, left, top, bottom */
};

/* scene/mod.rs:43-90 ; pools ping-pong, with two depth planes the result is back in `ret` */
size_t clip_against_planes(const Triangle &tri, const Planes &planes, Triangle *ret) {
    size_t pool_size = 0;
    if (planes.lateral[0].at_least_partially_visible(tri) &&
        planes.lateral[1].at_least_partially_visible(tri) &&
        planes.lateral[2].at_least_partially_visible(tri) &&
        planes.lateral[3].at_least_partially_visible(tri)) {
        ret[0] = tri;
        pool_size = 1;
    }
    Triangle new_pool[12];
    std::memset((void *)new_pool, 0, sizeof(new_pool)); /* :69-70 zeroed() */
    Triangle *pool_ref = ret;
    Triangle *new_ref = new_pool;
    for (const ViewPlane &plane : planes.depth) {
        size_t new_size = 0;
        for (size_t i = 0; i < pool_size; i++) new_size += plane.clip(pool_ref[i], new_ref + new_size);
        std::swap(pool_ref, new_ref);
        std::swap(pool_size, new_size);
    }
    return pool_size;
}

/* scene/mod.rs:282-294 */
struct Camera {
    Vec3 position, direction, up_direction;
    float top, bottom, right, left;
    float min_view_dist, max_view_dist;
    Vec3 u{0, 0, 0}, v{0, 0, 0}, w{0, 0, 0};
};
/* scene/mod.rs:297-357 */
Camera camera_new(Vec3 pos, Vec3 dir, float ratio) {
    Camera c;
    const float near = -10.0f;
    const float far = near - 500.0f;
    const float fov_x = 135.0f;
    /* f32::to_radians: self * (PI_f32 / 180.0_f32) */
    const float fov_x_rad = fov_x * (3.14159265358979323846f / 180.0f);
    c.right = std::fabs(near) * std::tan(fov_x_rad / 2.0f); /* :322, libm tanf */
    c.left = -c.right;
    c.top = (1.0f / ratio) * c.right; /* ratio.recip() */
    c.bottom = -c.top;
    c.position = pos;
    c.direction = normalized(dir);
    c.up_direction = {0.0f, 1.0f, 0.0f};
    c.min_view_dist = near;
    c.max_view_dist = far;
    return c;
}
/* scene/mod.rs:438-451 */
void update_basis(Camera &c) {
    Vec3 g = c.direction;
    Vec3 w = (g / norm(g)) * (-1.0f);
    Vec3 t_x_w = cross(c.up_direction, w);
    Vec3 u = t_x_w / norm(t_x_w);
    Vec3 v = cross(w, u);
    c.u = normalized(u);
    c.v = normalized(v);
    c.w = normalized(w);
}
/* scene/mod.rs:415-428 */
Mat4 basis_matrix(const Camera &c) {
    return {{{c.u.x, c.v.x, c.w.x, 0.0f},
             {c.u.y, c.v.y, c.w.y, 0.0f},
             {c.u.z, c.v.z, c.w.z, 0.0f},
             {0.0f, 0.0f, 0.0f, 1.0f}}};
}
/* scene/mod.rs:453-479 */
Mat4 gen_matrix(Camera &c) {
    Vec3 pos = c.position;
    Mat4 matrix_pos = {{{1.0f, 0.0f, 0.0f, -pos.x},
                        {0.0f, 1.0f, 0.0f, -pos.y},
                        {0.0f, 0.0f, 1.0f, -pos.z},
                        {0.0f, 0.0f, 0.0f, 1.0f}}};
    update_basis(c);
    return transposed(basis_matrix(c)) * matrix_pos;
}
/* scene/mod.rs:481-593 */
Planes gen_view_planes(Camera &c) {
    update_basis(c);
    const Mat4 mb = basis_matrix(c);
    const Vec3 cam = c.position;
    const float n = c.min_view_dist, f = c.max_view_dist;
    const float r = c.right, l = c.left, t = c.top, b = c.bottom;

    auto world = [&](Vec3 p) { return vec3_over_w(mb * as_vec4(p)) + cam; };
    const Vec3 ur_near = world({r, t, n});
    const Vec3 ul_near = world({l, t, n});
    const Vec3 lr_near = world({r, b, n});
    const Vec3 ll_near = world({l, b, n});

    const float x_center = (l + r) / 2.0f;
    const float y_center = (b + t) / 2.0f;
    const Vec3 upper_far = world({x_center, (f * t) / n, f});
    const Vec3 lower_far = world({x_center, (f * b) / n, f});
    const Vec3 right_far = world({(f * r) / n, y_center, f});
    const Vec3 left_far = world({(f * l) / n, y_center, f});

    const Vec3 visible = (ur_near + lower_far) / 2.0f;

    Planes p;
    p.depth[0] = view_plane_new(ur_near, lr_near, ll_near, visible);        /* near  :533 */
    p.depth[1] = view_plane_new(left_far, right_far, upper_far, visible);   /* far   :542 */
    p.lateral[0] = view_plane_new(right_far, ur_near, lr_near, visible);    /* right :554 */
    p.lateral[1] = view_plane_new(left_far, ll_near, ul_near, visible);     /* left  :563 */
    p.lateral[2] = view_plane_new(upper_far, ul_near, ur_near, visible);    /* top   :572 */
    p.lateral[3] = view_plane_new(lower_far, ll_near, lr_near, visible);    /* bottom:581 */
    return p;
}

/* scene/mod.rs:256-261 */
struct VertexVisual {
    Vec3 light, eye, halfway;
    float depth;
};

/* mesh.rs:31-35 ; (vertex idx, texture idx, normal idx) */
struct IndexedTri {
    uint32_t v[3], t[3], n[3];
};
struct IndexedMesh {
    std::vector<IndexedTri> triangles;
    uint32_t texture_idx;
};
/* object.rs:18-31 */
struct Object {
    std::vector<Vec3> vertices, normals_vertices, texture_vertices;
    std::vector<VertexVisual> vertices_visual_info;
    std::vector<IndexedMesh> opaque_meshes, transparent_meshes;
    std::vector<Texture> textures;
};

} // namespace

/* scene/mod.rs:749-757 */
struct orc_scene {
    size_t width, height;
    Camera camera;
    std::vector<Object> objects;
    Vec3 light_source;
    Texture default_texture = texture_default();
    orc_stats stats{};
    uint32_t tri_counter = 0;

    /* scene/mod.rs:817-899 */
    Mat4 gen_transformation_matrix() {
        const float n_x = (float)width, n_y = (float)height;
        const float n = camera.min_view_dist, f = camera.max_view_dist;
        const float r = camera.right, l = camera.left, t = camera.top, b = camera.bottom;
        const Mat4 matrix_cam = gen_matrix(camera);
        const Mat4 persp = {{{n, 0.0f, 0.0f, 0.0f},
                             {0.0f, n, 0.0f, 0.0f},
                             {0.0f, 0.0f, (n + f), -(n * f)},
                             {0.0f, 0.0f, 1.0f, 0.0f}}};
        const Mat4 orth = {{{2.0f / (r - l), 0.0f, 0.0f, -(r + l) / (r - l)},
                            {0.0f, 2.0f / (t - b), 0.0f, -(t + b) / (t - b)},
                            {0.0f, 0.0f, 2.0f / (n - f), -(n + f) / (n - f)},
                            {0.0f, 0.0f, 0.0f, 1.0f}}};
        const Mat4 viewport = {{{n_x / 2.0f, 0.0f, 0.0f, (n_x - 1.0f) / 2.0f},
                                {0.0f, n_y / 2.0f, 0.0f, (n_y - 1.0f) / 2.0f},
                                {0.0f, 0.0f, 1.0f, 0.0f},
                                {0.0f, 0.0f, 0.0f, 1.0f}}};
        return viewport * orth * persp * matrix_cam; /* :896, left-associative */
    }

    template <bool STATS, bool CULL>
    void draw_mesh_triangle(orc_canvas &canvas, Object &obj, const IndexedTri &it,
                            const Texture &tex, const Mat4 &matrix_transf, const Planes &planes,
                            Vec3 camera_pos);
    template <bool STATS>
    void render(orc_canvas &canvas);
};

/* scene/mod.rs:933-1085 (CULL=true, opaque) and 1117-1246 (CULL=false, transparent) */
template <bool STATS, bool CULL>
void orc_scene::draw_mesh_triangle(orc_canvas &canvas, Object &obj, const IndexedTri &it,
                                   const Texture &tex, const Mat4 &matrix_transf,
                                   const Planes &planes, Vec3 camera_pos) {
    if (STATS) stats.input_tris++;
    /* bookkeeping: draw id = 4 * (index of the input triangle in draw order) + clip output k */
    const uint32_t id_base = tri_counter * 4u;
    tri_counter++;
    Triangle original_tri;
    for (int i = 0; i < 3; i++) {
        const VertexVisual &vi = obj.vertices_visual_info[it.v[i]];
        original_tri.vertices[i] = obj.vertices[it.v[i]];
        VertexAttributes &va = original_tri.vertices_attr[i];
        va.screen_coord = {0.0f, 0.0f};
        va.depth = vi.depth;
        va.normal = obj.normals_vertices[it.n[i]];
        va.light = vi.light;
        va.halfway = vi.halfway;
        va.texture_coord = obj.texture_vertices[it.t[i]];
    }
    if (CULL) {
        /* Triangle::calc_normal :30-41, get_center :92-99, cull :1016-1027 */
        const Vec3 a = original_tri.vertices[0], b = original_tri.vertices[1],
                   c = original_tri.vertices[2];
        const Vec3 tri_normal = cross(b - a, c - b);
        Vec3 sum{0.0f, 0.0f, 0.0f};
        sum = sum + a;
        sum = sum + b;
        sum = sum + c;
        const Vec3 center = sum / 3.0f;
        const Vec3 tri_eye = camera_pos - center;
        if (dot(tri_eye, tri_normal) <= 0.0f) {
            if (STATS) stats.culled_tris++;
            return;
        }
    }
    Triangle clipped[12];
    std::memset((void *)clipped, 0, sizeof(clipped)); /* :1031 zeroed() */
    const size_t count = clip_against_planes(original_tri, planes, clipped);
    for (size_t k = 0; k < count; k++) {
        Triangle &ct = clipped[k];
        for (int i = 0; i < 3; i++) { /* :1047-1063 */
            const Vec4 p = matrix_transf * as_vec4(ct.vertices[i]);
            ct.vertices_attr[i].screen_coord = Vec2{p.a[0], p.a[1]} / p.a[3];
        }
        if (STATS) stats.emitted_tris++;
        canvas.cur_id = id_base + (uint32_t)k;
        canvas.draw_triangle_with_attributes<STATS>(ct.vertices_attr[0], ct.vertices_attr[1],
                                                    ct.vertices_attr[2], tex);
    }
}

/* scene/mod.rs:901-1249 */
template <bool STATS>
void orc_scene::render(orc_canvas &canvas) {
    canvas.stats = STATS ? &stats : nullptr;
    canvas.clear(); /* :902 */
    tri_counter = 0;
    if (STATS) stats = orc_stats{};

    const Mat4 matrix_transf = gen_transformation_matrix(); /* :904 */
    const Vec3 camera_pos = camera.position;
    const Planes planes = gen_view_planes(camera); /* :908 */

    for (Object &obj : objects) {
        /* :917-926 */
        for (size_t i = 0; i < obj.vertices.size(); i++) {
            const Vec3 vertex = obj.vertices[i];
            VertexVisual &vi = obj.vertices_visual_info[i];
            const Vec3 eye_dir = vertex - camera_pos;
            vi.light = normalized(vertex - light_source);
            vi.eye = normalized(eye_dir);
            vi.depth = norm(eye_dir);
            vi.halfway = normalized(vi.light + vi.eye);
        }

        canvas.depth_update_enabled = true; /* :928 */
        for (const IndexedMesh &mesh : obj.opaque_meshes) {
            /* :1072-1075 : out-of-range texture index falls back to Texture::default() */
            const Texture &tex =
                mesh.texture_idx < obj.textures.size() ? obj.textures[mesh.texture_idx] : default_texture;
            for (const IndexedTri &it : mesh.triangles)
                draw_mesh_triangle<STATS, true>(canvas, obj, it, tex, matrix_transf, planes, camera_pos);
        }

        canvas.depth_update_enabled = false; /* :1088 */
        for (IndexedMesh &mesh : obj.transparent_meshes) {
            const Texture &tex =
                mesh.texture_idx < obj.textures.size() ? obj.textures[mesh.texture_idx] : default_texture;
            /* :1100-1115 : stable sort, far -> near, by total_cmp of centroid distance;
             * the sorted order persists in the mesh across frames */
            std::stable_sort(mesh.triangles.begin(), mesh.triangles.end(),
                             [&](const IndexedTri &a, const IndexedTri &b) {
                                 const Vec3 a_center = (obj.vertices[a.v[0]] + obj.vertices[a.v[1]] +
                                                        obj.vertices[a.v[2]]) / 3.0f;
                                 const Vec3 b_center = (obj.vertices[b.v[0]] + obj.vertices[b.v[1]] +
                                                        obj.vertices[b.v[2]]) / 3.0f;
                                 const float a_depth = dist(a_center, camera_pos);
                                 const float b_depth = dist(b_center, camera_pos);
                                 return total_key(a_depth) > total_key(b_depth);
                             });
            for (const IndexedTri &it : mesh.triangles)
                draw_mesh_triangle<STATS, false>(canvas, obj, it, tex, matrix_transf, planes, camera_pos);
        }
    }
    canvas.stats = nullptr;
}

/* ------------------------------------------------------------------ C interface */

extern "C" {

orc_scene *orc_scene_new(size_t width, size_t height) { /* scene/mod.rs:760-786 */
    orc_scene *s = new orc_scene();
    s->width = width;
    s->height = height;
    const Vec3 camera_pos{0.0f, 0.0f, 150.0f};
    const Vec3 camera_dir = camera_pos * -1.0f;
    s->light_source = {0.0f, 300.0f, 300.0f};
    const float ratio = (float)width / (float)height;
    s->camera = camera_new(camera_pos, camera_dir, ratio);
    return s;
}
void orc_scene_free(orc_scene *s) { delete s; }

int orc_scene_add_object(orc_scene *s, const float *positions, size_t n_pos, const float *normals,
                         size_t n_nrm, const float *uvs, size_t n_uv, const orc_mesh *meshes,
                         size_t n_meshes, const orc_material *materials, size_t n_materials) {
    Object obj;
    auto fill = [](std::vector<Vec3> &dst, const float *src, size_t n) {
        dst.resize(n);
        for (size_t i = 0; i < n; i++) dst[i] = {src[3 * i], src[3 * i + 1], src[3 * i + 2]};
    };
    fill(obj.vertices, positions, n_pos);
    fill(obj.normals_vertices, normals, n_nrm);
    fill(obj.texture_vertices, uvs, n_uv);
    for (size_t i = 0; i < n_materials; i++) {
        const orc_material &m = materials[i];
        Texture t;
        t.ka = {m.ka[0], m.ka[1], m.ka[2]};
        t.kd = {m.kd[0], m.kd[1], m.kd[2]};
        t.ks = {m.ks[0], m.ks[1], m.ks[2]};
        t.alpha = m.alpha;
        t.map_ka = m.map_ka ? texture_map_new(m.map_ka, m.map_ka_w, m.map_ka_h, m.map_ka_comp)
                            : texture_map_default();
        t.map_kd = m.map_kd ? texture_map_new(m.map_kd, m.map_kd_w, m.map_kd_h, m.map_kd_comp)
                            : texture_map_default();
        obj.textures.push_back(std::move(t));
    }
    for (size_t i = 0; i < n_meshes; i++) {
        IndexedMesh mesh;
        mesh.texture_idx = meshes[i].texture_idx;
        mesh.triangles.resize(meshes[i].n_tris);
        for (size_t k = 0; k < meshes[i].n_tris; k++) {
            const uint32_t *p = meshes[i].tris + 9 * k;
            IndexedTri &it = mesh.triangles[k];
            for (int j = 0; j < 3; j++) {
                it.v[j] = p[j];
                it.t[j] = p[3 + j];
                it.n[j] = p[6 + j];
                if (it.v[j] >= n_pos || it.t[j] >= n_uv || it.n[j] >= n_nrm) return -1;
            }
        }
        /* Object::new, object.rs:45-53 (textures[texture_idx] must exist there) */
        if (mesh.texture_idx >= obj.textures.size()) return -2;
        if (obj.textures[mesh.texture_idx].alpha < 1.0f)
            obj.transparent_meshes.push_back(std::move(mesh));
        else
            obj.opaque_meshes.push_back(std::move(mesh));
    }
    obj.vertices_visual_info.assign(n_pos, VertexVisual{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, 0.0f});
    s->objects.push_back(std::move(obj));
    return (int)s->objects.size() - 1;
}

void orc_scene_set_camera(orc_scene *s, const float pos[3], const float dir[3]) {
    const float ratio = (float)s->width / (float)s->height;
    s->camera = camera_new({pos[0], pos[1], pos[2]}, {dir[0], dir[1], dir[2]}, ratio);
}
void orc_scene_set_light(orc_scene *s, const float pos[3]) {
    s->light_source = {pos[0], pos[1], pos[2]};
}
void orc_scene_camera_move(orc_scene *s, int which, float dist_) {
    Camera &c = s->camera;
    switch (which) { /* scene/mod.rs:381-405 */
    case 0: c.position = c.position + c.up_direction * dist_; break;
    case 1: c.position = c.position + c.up_direction * (-dist_); break;
    case 2: c.position = c.position + c.u * (-dist_); break;
    case 3: c.position = c.position + c.u * dist_; break;
    case 4: c.position = c.position + normalized(cross(c.up_direction, c.u)) * dist_; break;
    case 5: c.position = c.position + normalized(cross(c.u, c.up_direction)) * dist_; break;
    default: break;
    }
}
void orc_scene_move_camera_direction(orc_scene *s, int dx, int dy) { /* scene/mod.rs:803-813 */
    Camera &c = s->camera;
    const float fx = (float)dx / (float)s->width;
    const float fy = (float)dy / (float)s->height;
    c.direction = normalized(c.direction + c.u * fx + c.v * fy); /* :430-432 */
    update_basis(c);
}
void orc_scene_get_camera(orc_scene *s, float pos[3], float dir[3]) {
    pos[0] = s->camera.position.x; pos[1] = s->camera.position.y; pos[2] = s->camera.position.z;
    dir[0] = s->camera.direction.x; dir[1] = s->camera.direction.y; dir[2] = s->camera.direction.z;
}
void orc_scene_uniforms(orc_scene *s, float m[16], float planes[24]) {
    const Mat4 mt = s->gen_transformation_matrix();
    std::memcpy(m, mt.a, sizeof(float) * 16);
    const Planes p = gen_view_planes(s->camera);
    const ViewPlane *all[6] = {&p.depth[0], &p.depth[1], &p.lateral[0], &p.lateral[1], &p.lateral[2], &p.lateral[3]};
    for (int i = 0; i < 6; i++) {
        planes[4 * i + 0] = all[i]->normal.x;
        planes[4 * i + 1] = all[i]->normal.y;
        planes[4 * i + 2] = all[i]->normal.z;
        planes[4 * i + 3] = all[i]->k;
    }
}
int orc_scene_vertex_visual(orc_scene *s, size_t obj, float *out, size_t n_vertices) {
    if (obj >= s->objects.size()) return -1;
    const Object &o = s->objects[obj];
    if (n_vertices != o.vertices_visual_info.size()) return -2;
    for (size_t i = 0; i < n_vertices; i++) {
        const VertexVisual &v = o.vertices_visual_info[i];
        float *d = out + 10 * i;
        d[0] = v.light.x; d[1] = v.light.y; d[2] = v.light.z;
        d[3] = v.eye.x; d[4] = v.eye.y; d[5] = v.eye.z;
        d[6] = v.halfway.x; d[7] = v.halfway.y; d[8] = v.halfway.z;
        d[9] = v.depth;
    }
    return 0;
}

orc_canvas *orc_canvas_new(size_t width, size_t height) { /* canvas.rs:366-380 */
    orc_canvas *c = new orc_canvas();
    c->width = width;
    c->height = height;
    c->frame.assign(width * height, pixel_new(0, 0, 0));
    c->winner.assign(width * height, NO_WINNER);
    return c;
}
void orc_canvas_free(orc_canvas *c) { delete c; }
void orc_canvas_init_depth(orc_canvas *c, float depth) { c->init_depth(depth); }
void orc_canvas_apply_offset(orc_canvas *c, int x, int y) { /* canvas.rs:382-385 */
    c->offset.x = (float)x;
    c->offset.y = (float)y;
}
void orc_canvas_resize(orc_canvas *c, size_t width, size_t height) { /* canvas.rs:387-393 */
    c->width = width;
    c->height = height;
    c->frame.resize(width * height, pixel_new(0, 0, 0));
    c->winner.assign(width * height, NO_WINNER);
    c->init_depth(c->depth_max);
}
void orc_canvas_clear(orc_canvas *c) { c->clear(); }
const uint8_t *orc_canvas_bytes(orc_canvas *c, size_t *len) {
    if (len) *len = c->frame.size() * sizeof(Pixel);
    return reinterpret_cast<const uint8_t *>(c->frame.data());
}
const float *orc_canvas_depth(orc_canvas *c, size_t *len) {
    if (len) *len = c->depth_frame.size();
    return c->depth_frame.data();
}
const uint32_t *orc_canvas_winner(orc_canvas *c, size_t *len) {
    if (len) *len = c->winner.size();
    return c->winner.data();
}
/* canvas.rs:435-575 */
void orc_canvas_draw_triangle(orc_canvas *cv, const orc_vertex2d v[3], const uint8_t *rgba, uint32_t tw, uint32_t th,
                              const uint64_t clip[4]) {
    orc_canvas &self = *cv;
    const Vec2 a_center = pos_map_center({v[0].x, v[0].y}); /* :449-451: no canvas offset on this path */
    const Vec2 b_center = pos_map_center({v[1].x, v[1].y});
    const Vec2 c_center = pos_map_center({v[2].x, v[2].y});
    const Vec2 a_uv{v[0].u, v[0].v}, b_uv{v[1].u, v[1].v}, c_uv{v[2].u, v[2].v};
    const Pixel color_a = pixel_new(v[0].r, v[0].g, v[0].b); /* Color::Custom(col).as_pixel() :41 */
    const Pixel color_b = pixel_new(v[1].r, v[1].g, v[1].b);
    const Pixel color_c = pixel_new(v[2].r, v[2].g, v[2].b);
    /* :461-480 */
    auto f_ab = [&](float x, float y) {
        return (a_center.y - b_center.y) * x + (b_center.x - a_center.x) * y + (a_center.x * b_center.y) - (b_center.x * a_center.y);
    };
    auto f_bc = [&](float x, float y) {
        return (b_center.y - c_center.y) * x + (c_center.x - b_center.x) * y + (b_center.x * c_center.y) - (c_center.x * b_center.y);
    };
    auto f_ca = [&](float x, float y) {
        return (c_center.y - a_center.y) * x + (a_center.x - c_center.x) * y + (c_center.x * a_center.y) - (a_center.x * c_center.y);
    };
    /* :504-521 */
    uint64_t x_min = sat_usize(min3(a_center.x, b_center.x, c_center.x)), y_min = sat_usize(min3(a_center.y, b_center.y, c_center.y));
    uint64_t x_max = sat_usize(max3(a_center.x, b_center.x, c_center.x)), y_max = sat_usize(max3(a_center.y, b_center.y, c_center.y));
    Rect drawable = rect_from_coords(x_min, y_min, x_max, y_max);
    const Rect screen = rect_from_coords(0, 0, self.width - 1, self.height - 1);
    drawable = rect_clip(drawable, screen);
    const Rect valid = rect_clip(clip ? rect_from_coords(clip[0], clip[1], clip[2], clip[3]) : screen, drawable);
    x_min = valid.x; y_min = valid.y; x_max = valid.x_max(); y_max = valid.y_max();
    /* :523-529 */
    const float f_alpha = f_bc(a_center.x, a_center.y), f_beta = f_ca(b_center.x, b_center.y), f_gama = f_ab(c_center.x, c_center.y);
    const float f_alpha_outside = f_bc(-1.0f, -1.0f), f_beta_outside = f_ca(-1.0f, -1.0f), f_gama_outside = f_ab(-1.0f, -1.0f);
    const float f_width = (float)tw, f_height = (float)th;
    for (uint64_t y = y_min; y <= y_max; y++) { /* :531-573 */
        const float y_f32 = (float)y;
        for (uint64_t x = x_min; x <= x_max; x++) {
            const float x_f32 = (float)x;
            const float alpha = f_bc(x_f32, y_f32) / f_alpha, beta = f_ca(x_f32, y_f32) / f_beta, gama = f_ab(x_f32, y_f32) / f_gama;
            if (!(alpha >= 0.0f && beta >= 0.0f && gama >= 0.0f)) continue;
            if (!((alpha > 0.0f || f_alpha * f_alpha_outside > 0.0f) && (beta > 0.0f || f_beta * f_beta_outside > 0.0f) &&
                  (gama > 0.0f || f_gama * f_gama_outside > 0.0f)))
                continue;
            Pixel color_pixel = pixel_add(pixel_add(pixel_mul(color_a, alpha), pixel_mul(color_b, beta)), pixel_mul(color_c, gama)); /* :546 */
            const float color_alpha = (alpha * v[0].alpha) + (beta * v[1].alpha) + (gama * v[2].alpha);                            /* :548-550 */
            const Vec2 color_uv = (a_uv * alpha) + (b_uv * beta) + (c_uv * gama);                                                      /* :552 */
            /* get_rgba_slice, scene/mod.rs:137-152: u_idx = floor(u * w), v_idx = h - 1 - floor(v * h); indices clamped into
             * the map for memory safety (SURVEY.md 8c deviation 6; the reference would index out of bounds) */
            uint64_t u_idx = sat_usize(std::floor(color_uv.x * f_width)), v_raw = sat_usize(std::floor(color_uv.y * f_height));
            if (u_idx > tw - 1u) u_idx = tw - 1u;
            if (v_raw > th - 1u) v_raw = th - 1u;
            const uint8_t *px = rgba + ((th - 1u - v_raw) * (uint64_t)tw + u_idx) * 4u;
            const float texture_alpha = (float)px[3] / 255.0f;                                                           /* :555 */
            const Pixel color_texture = pixel_new(px[0], px[1], px[2]);
            color_pixel = pixel_add(pixel_mul(color_pixel, texture_alpha), pixel_mul(color_texture, 1.0f - texture_alpha)); /* :562-563 */
            const float final_alpha = color_alpha * texture_alpha;                                                      /* :566 */
            self.cur_id = NO_WINNER - 1u;
            self.draw_pixel_coord_with_depth((size_t)x, (size_t)y, color_pixel, final_alpha, 0.0f);                      /* :568 */
        }
    }
}
void orc_canvas_set_depth_update(orc_canvas *c, int enabled) { c->depth_update_enabled = enabled != 0; }

void orc_scene_render(orc_scene *s, orc_canvas *c, int count_stats) {
    if (count_stats)
        s->render<true>(*c);
    else
        s->render<false>(*c);
}
void orc_scene_stats(orc_scene *s, orc_stats *out) { *out = s->stats; }

} /* extern "C" */
