// k_overlay.cu — Canvas::draw_triangle (mororo18/draw src/renderer/canvas.rs:435-575), the 2-D path the GUI draws
// its command lists with (src/app/gui.rs:382-485): vertex-coloured, RGBA-textured, alpha-blended screen triangles
// written over the frame in SUBMISSION ORDER with a depth of 0.0.
//
// The reference walks every triangle's clipped bounding box pixel by pixel, one triangle after the other.  Here a
// submission — the draw commands of a GUI frame that share a texture, each with its own clipping rectangle — is two
// launches per 32 768 triangles:
//
//   k_overlay_setup  one thread per triangle: snapped vertices, the three literal edge functions, their values at
//                    the opposite vertex and at (-1,-1), the clipped pixel rectangle (Rectangle::from_coords / clip
//                    in u64, empties collapsing to (0,0) as the reference's do) -> one 80-byte record; and the
//                    triangle's bit set in the mask of every 64x64-pixel bin its rectangle touches.
//   k_overlay_draw   one CTA per 32x8-pixel tile, one pixel per thread.  The pixel's colour and depth live in
//                    registers while the CTA walks its bin's mask in bit order = submission order, so blending and
//                    the depth test see exactly the sequence of writes the reference's loop produces; one store at
//                    the end.  Tiles of bins nothing touches leave at the first load.
//
// Arithmetic is the reference's, literally: barycentrics by IEEE division, Pixel * f32 truncating to u8 per channel,
// Pixel + Pixel wrapping (release build), get_rgba_slice's floor(u * w), texture alpha / 255.
#include "device_math.cuh"

namespace drawb200 {

constexpr int OV_TILE_W = 32, OV_TILE_H = 8, OV_THREADS = OV_TILE_W * OV_TILE_H;
constexpr int OV_BIN = OVERLAY_BIN; // bin edge in pixels: 2 x 8 tiles

// VertexSimpleAttributes (canvas.rs:185-191) as draw_vertex2d lays it out
struct Vertex2D {
    float x, y, u, v;
    uint8_t r, g, b, pad;
    float alpha;
};
static_assert(sizeof(Vertex2D) == 24, "draw_vertex2d layout");

struct OverlayRec { // 20 words
    Edge e[3];      // bc (alpha), ca (beta), ab (gama)
    float f[3];     // f_alpha, f_beta, f_gama (:523-525)
    uint32_t flags; // bit i: f_i * f_i(-1,-1) > 0 (:539-541)
    uint32_t x01, y01, pad0, pad1; // x_min | x_max << 16, y_min | y_max << 16 (canvas y)
};
static_assert(sizeof(OverlayRec) == OVERLAY_REC_BYTES, "OverlayRec is five 16-byte words");

struct RectU64 { // Rectangle, canvas.rs:293-351
    unsigned long long x, y, w, h;
};
__device__ __forceinline__ RectU64 rect_from_coords(unsigned long long x0, unsigned long long y0, unsigned long long x1,
                                                    unsigned long long y1) {
    const unsigned long long xm = min(x0, x1), ym = min(y0, y1), xM = max(x0, x1), yM = max(y0, y1);
    return {xm, ym, xM - xm, yM - ym};
}
__device__ __forceinline__ RectU64 rect_clip(const RectU64 &a, const RectU64 &b) {
    unsigned long long xm = max(a.x, b.x), ym = max(a.y, b.y);
    unsigned long long xM = min(a.x + a.w, b.x + b.w), yM = min(a.y + a.h, b.y + b.h);
    if (xm > xM) xm = xM = 0;
    if (ym > yM) ym = yM = 0;
    return rect_from_coords(xm, ym, xM, yM);
}
// the closures at canvas.rs:482-502: strict comparisons from +-inf, so a NaN never wins
__device__ __forceinline__ float min3_ref(float a, float b, float c) {
    float r = __int_as_float(0x7f800000);
    if (a < r) r = a;
    if (b < r) r = b;
    if (c < r) r = c;
    return r;
}
__device__ __forceinline__ float max3_ref(float a, float b, float c) {
    float r = __int_as_float(0xff800000);
    if (a > r) r = a;
    if (b > r) r = b;
    if (c > r) r = c;
    return r;
}
__device__ __forceinline__ unsigned long long sat_usize_nan0(float v) { return v == v ? __float2ull_rz(v) : 0ull; }

__global__ void __launch_bounds__(128) k_overlay_setup(const OverlayParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const Vertex2D *v = static_cast<const Vertex2D *>(P.verts) + 3ull * i;
    float px[3], py[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { // pos_map_center :896-904, no canvas offset on this path (:449-451)
        px[k] = floorf(FADD(v[k].x, 0.5f));
        py[k] = floorf(FADD(v[k].y, 0.5f));
    }
    OverlayRec r;
    r.e[0] = make_edge(px[1], py[1], px[2], py[2]); // f_bc
    r.e[1] = make_edge(px[2], py[2], px[0], py[0]); // f_ca
    r.e[2] = make_edge(px[0], py[0], px[1], py[1]); // f_ab
    r.flags = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        r.f[k] = edge_eval(r.e[k], px[k], py[k]);
        if (FMUL(r.f[k], edge_eval(r.e[k], -1.0f, -1.0f)) > 0.0f) r.flags |= 1u << k;
    }
    // :504-521
    const RectU64 screen = rect_from_coords(0, 0, P.width - 1, P.height - 1);
    RectU64 drawable = rect_from_coords(sat_usize_nan0(min3_ref(px[0], px[1], px[2])), sat_usize_nan0(min3_ref(py[0], py[1], py[2])),
                                        sat_usize_nan0(max3_ref(px[0], px[1], px[2])), sat_usize_nan0(max3_ref(py[0], py[1], py[2])));
    drawable = rect_clip(drawable, screen);
    // the triangle's draw command (commands own consecutive ranges of the submission): its clipping rectangle
    uint32_t lo = 0, hi = P.n_cmds; // largest c with cmd_first[c] <= number of the triangle
    const unsigned long long number = (unsigned long long)P.first + i;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(P.cmd_first + mid) <= number) lo = mid; else hi = mid;
    }
    const OverlayClip cl = P.cmd_clip[lo];
    const RectU64 valid = rect_clip(cl.has ? rect_from_coords(cl.c[0], cl.c[1], cl.c[2], cl.c[3]) : screen, drawable);
    const uint32_t x0 = (uint32_t)valid.x, y0 = (uint32_t)valid.y, x1 = (uint32_t)(valid.x + valid.w), y1 = (uint32_t)(valid.y + valid.h);
    r.x01 = x0 | x1 << 16;
    r.y01 = y0 | y1 << 16;
    r.pad0 = r.pad1 = 0;
    uint4 *dst = reinterpret_cast<uint4 *>(static_cast<OverlayRec *>(P.recs) + i);
    const uint4 *src = reinterpret_cast<const uint4 *>(&r);
#pragma unroll
    for (int k = 0; k < 5; k++) dst[k] = src[k];
    const uint32_t bit = 1u << (i & 31u), word = i >> 5;
    for (uint32_t by = y0 / OV_BIN; by <= y1 / OV_BIN; by++)
        for (uint32_t bx = x0 / OV_BIN; bx <= x1 / OV_BIN; bx++) {
            const uint32_t bin = by * P.bins_x + bx;
            atomicOr(P.masks + (size_t)bin * P.words + word, bit);
            P.bin_any[bin] = 1u;
        }
}

// Pixel * f32 (canvas.rs:154-169) on the three colour bytes of a b | g << 8 | r << 16 word; the pad byte comes out 0
__device__ __forceinline__ uint32_t px_mul(uint32_t p, float f) {
    return sat_u8(FMUL((float)(p & 255u), f)) | sat_u8(FMUL((float)(p >> 8 & 255u), f)) << 8 | sat_u8(FMUL((float)(p >> 16 & 255u), f)) << 16;
}
// Pixel + Pixel (canvas.rs:136-152): per-channel u8 add, wrapping as a release build does; pad 0
__device__ __forceinline__ uint32_t px_add(uint32_t a, uint32_t b) {
    return (((a & 0x00ff00ffu) + (b & 0x00ff00ffu)) & 0x00ff00ffu) | (((a & 0x0000ff00u) + (b & 0x0000ff00u)) & 0x0000ff00u);
}

__global__ void __launch_bounds__(OV_THREADS) k_overlay_draw(const OverlayParams P) {
    const uint32_t tiles_x = (P.width + OV_TILE_W - 1) / OV_TILE_W;
    const uint32_t tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const uint32_t bin = (ty * OV_TILE_H / OV_BIN) * P.bins_x + tx * OV_TILE_W / OV_BIN;
    if (!P.bin_any[bin]) return;
    const uint32_t tile_x0 = tx * OV_TILE_W, tile_y0 = ty * OV_TILE_H;
    const uint32_t tile_x1 = tile_x0 + OV_TILE_W - 1, tile_y1 = tile_y0 + OV_TILE_H - 1;
    const uint32_t x = tile_x0 + (threadIdx.x & 31u), y = tile_y0 + (threadIdx.x >> 5);
    const bool inside = x < P.width && y < P.height;
    const size_t c_at = (size_t)(P.height - 1 - min(y, P.height - 1)) * P.width + x, d_at = (size_t)y * P.width + x;
    uint32_t color = 0;
    float depth = 0.0f;
    if (inside) {
        color = P.color[c_at];
        depth = P.depth[d_at];
    }
    bool dirty = false;
    const float xf = (float)x, yf = (float)y;
    const uint32_t *mask = P.masks + (size_t)bin * P.words;
    for (uint32_t w = 0; w < P.words; w++) {
        uint32_t m = __ldg(mask + w);
        while (m) {
            const uint32_t i = w * 32u + (__ffs(m) - 1);
            m &= m - 1;
            const uint4 *src = reinterpret_cast<const uint4 *>(static_cast<const OverlayRec *>(P.recs) + i);
            const uint4 q4 = __ldg(src + 4);
            const uint32_t rx0 = q4.x & 0xffffu, rx1 = q4.x >> 16, ry0 = q4.y & 0xffffu, ry1 = q4.y >> 16;
            if (rx0 > tile_x1 || rx1 < tile_x0 || ry0 > tile_y1 || ry1 < tile_y0) continue; // whole CTA
            if (!(inside && x >= rx0 && x <= rx1 && y >= ry0 && y <= ry1)) continue;
            const uint4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2), q3 = __ldg(src + 3);
            const Edge e_bc{__uint_as_float(q0.x), __uint_as_float(q0.y), __uint_as_float(q0.z), __uint_as_float(q0.w)};
            const Edge e_ca{__uint_as_float(q1.x), __uint_as_float(q1.y), __uint_as_float(q1.z), __uint_as_float(q1.w)};
            const Edge e_ab{__uint_as_float(q2.x), __uint_as_float(q2.y), __uint_as_float(q2.z), __uint_as_float(q2.w)};
            const float alpha = FDIV(edge_eval(e_bc, xf, yf), __uint_as_float(q3.x)); // :535-537
            const float beta = FDIV(edge_eval(e_ca, xf, yf), __uint_as_float(q3.y));
            const float gama = FDIV(edge_eval(e_ab, xf, yf), __uint_as_float(q3.z));
            if (!(alpha >= 0.0f && beta >= 0.0f && gama >= 0.0f)) continue;
            if (!((alpha > 0.0f || (q3.w & 1u)) && (beta > 0.0f || (q3.w & 2u)) && (gama > 0.0f || (q3.w & 4u)))) continue;
            const Vertex2D *v = static_cast<const Vertex2D *>(P.verts) + 3ull * i;
            const uint2 ca = __ldg(reinterpret_cast<const uint2 *>(&v[0].r)), cb = __ldg(reinterpret_cast<const uint2 *>(&v[1].r)),
                        cc = __ldg(reinterpret_cast<const uint2 *>(&v[2].r));
            // Color::Custom([r, g, b]).as_pixel(): memory r, g, b -> word b | g << 8 | r << 16
            auto as_pixel = [](uint32_t rgb) { return (rgb >> 16 & 255u) | (rgb & 0xff00u) | (rgb & 255u) << 16; };
            uint32_t px = px_add(px_add(px_mul(as_pixel(ca.x), alpha), px_mul(as_pixel(cb.x), beta)), px_mul(as_pixel(cc.x), gama)); // :546
            const float c_alpha =
                FADD(FADD(FMUL(alpha, __uint_as_float(ca.y)), FMUL(beta, __uint_as_float(cb.y))), FMUL(gama, __uint_as_float(cc.y))); // :548-550
            const float2 ta = __ldg(reinterpret_cast<const float2 *>(&v[0].u)), tb = __ldg(reinterpret_cast<const float2 *>(&v[1].u)),
                         tc = __ldg(reinterpret_cast<const float2 *>(&v[2].u));
            const float u = FADD(FADD(FMUL(ta.x, alpha), FMUL(tb.x, beta)), FMUL(tc.x, gama)); // :552
            const float vv = FADD(FADD(FMUL(ta.y, alpha), FMUL(tb.y, beta)), FMUL(tc.y, gama));
            // get_rgba_slice, scene/mod.rs:137-152; indices clamped into the map (the reference would panic)
            const unsigned long long ui = min(sat_usize_nan0(floorf(FMUL(u, (float)P.tex_w))), (unsigned long long)P.tex_w - 1ull);
            const unsigned long long vi = min(sat_usize_nan0(floorf(FMUL(vv, (float)P.tex_h))), (unsigned long long)P.tex_h - 1ull);
            const uchar4 t = __ldg(static_cast<const uchar4 *>(P.texels) + ((size_t)(P.tex_h - 1u - (uint32_t)vi) * P.tex_w + (uint32_t)ui));
            const float t_alpha = FDIV((float)t.w, 255.0f); // :555
            px = px_add(px_mul(px, t_alpha), px_mul((uint32_t)t.z | (uint32_t)t.y << 8 | (uint32_t)t.x << 16, FSUB(1.0f, t_alpha))); // :562-563
            const float opacity = FMUL(c_alpha, t_alpha);   // :566
            // draw_pixel_coord_with_depth(x, y, colour, opacity, 0.0) :906-930
            const uint32_t fresh = opacity < 1.0f ? px_add(px_mul(color, FSUB(1.0f, opacity)), px_mul(px, opacity)) : px;
            if (0.0f < depth) {
                color = fresh;
                dirty = true;
                if (P.depth_update) depth = 0.0f;
            }
        }
    }
    if (dirty) {
        P.color[c_at] = color;
        if (P.depth_update) P.depth[d_at] = depth;
    }
}

// Enqueues one batch (<= 65535 x 32 triangles per call is the caller's business: words sizes the masks).
cudaError_t launch_overlay(const OverlayParams &P, cudaStream_t stream, uint64_t *launches) {
    const size_t n_bins = (size_t)P.bins_x * P.bins_y;
    cudaError_t err = cudaMemsetAsync(P.masks, 0, n_bins * P.words * sizeof(uint32_t), stream);
    if (err != cudaSuccess) return err;
    err = cudaMemsetAsync(P.bin_any, 0, n_bins * sizeof(uint32_t), stream);
    if (err != cudaSuccess) return err;
    k_overlay_setup<<<(P.n + 127) / 128, 128, 0, stream>>>(P);
    const uint32_t tiles = ((P.width + OV_TILE_W - 1) / OV_TILE_W) * ((P.height + OV_TILE_H - 1) / OV_TILE_H);
    k_overlay_draw<<<tiles, OV_THREADS, 0, stream>>>(P);
    *launches += 2;
    return cudaGetLastError();
}

} // namespace drawb200
