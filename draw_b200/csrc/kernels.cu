// kernels.cu — the sm_100a frame pipeline behind Scene::render (mororo18/draw scene/mod.rs:901).
//
//   k_vertex   per vertex   light / halfway / depth (scene/mod.rs:917-926), screen xy of the
//                           unclipped vertex (:1047-1058), view-plane side flags (:634-660)
//   k_setup    per triangle gather, back-face cull (:1016-1027), lateral reject + near/far clip
//                           (:43-90, :662-746), snap + bbox + edge constants (canvas.rs:585-666),
//                           record allocation (block prefix sums), tile counting
//   k_scan     one CTA      exclusive scan of the per-tile counts
//   k_fill     per record   scatter record slots into per-tile lists
//   k_tile     per tile     edge-function raster + depth test (canvas.rs:668-682, 906-930),
//                           deferred Phong + texel fetch (canvas.rs:685-743), ordered transparent
//                           blend, fused clear, single write of colour + depth to HBM
//
// Numerical contract: every float operation below is a single
// round-to-nearest binary32 operation in the reference's evaluation order.  All arithmetic
// goes through __fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn, which ptxas never fuses
// or reassociates; the file is also compiled with -fmad=false.  No tensor cores: nothing
// here is a dense contraction.
//
// Draw-order semantics without ordered lists: the reference draws triangles sequentially with
// a strict `<` depth test, so for opaque triangles the surviving fragment of a pixel is the
// minimum of (depth, draw id) — ties go to the earlier triangle.  k_tile therefore consumes
// its tile list in any order and breaks depth ties by draw id.  Transparent triangles
// (depth test on, depth write off, blend with the current colour) only see the final opaque
// winner W of their pixel if they are drawn after it: a transparent fragment T is blended iff
// id(T) > id(W) and depth(T) < depth(W); they are applied in draw order in a second phase.
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_types.h"

namespace drawb200 {

#define FADD(a, b) __fadd_rn((a), (b))
#define FSUB(a, b) __fsub_rn((a), (b))
#define FMUL(a, b) __fmul_rn((a), (b))
#define FDIV(a, b) __fdiv_rn((a), (b))

struct v3 {
    float x, y, z;
};
__device__ __forceinline__ v3 v_add(v3 a, v3 b) { return {FADD(a.x, b.x), FADD(a.y, b.y), FADD(a.z, b.z)}; }
__device__ __forceinline__ v3 v_sub(v3 a, v3 b) { return {FSUB(a.x, b.x), FSUB(a.y, b.y), FSUB(a.z, b.z)}; }
__device__ __forceinline__ v3 v_div(v3 a, float s) { return {FDIV(a.x, s), FDIV(a.y, s), FDIV(a.z, s)}; }
// linalg.rs:182-184
__device__ __forceinline__ float v_dot(v3 a, v3 b) {
    return FADD(FADD(FMUL(a.x, b.x), FMUL(a.y, b.y)), FMUL(a.z, b.z));
}
// linalg.rs:167-171
__device__ __forceinline__ float v_norm(v3 a) {
    return __fsqrt_rn(FADD(FADD(FMUL(a.x, a.x), FMUL(a.y, a.y)), FMUL(a.z, a.z)));
}
// linalg.rs:186-200
__device__ __forceinline__ v3 v_cross(v3 a, v3 b) {
    return {FSUB(FMUL(a.y, b.z), FMUL(a.z, b.y)), FSUB(FMUL(a.z, b.x), FMUL(a.x, b.z)),
            FSUB(FMUL(a.x, b.y), FMUL(a.y, b.x))};
}
// ViewPlane::func, scene/mod.rs:634-636
__device__ __forceinline__ float plane_eval(const float *pl, v3 p) {
    return FADD(v_dot(v3{pl[0], pl[1], pl[2]}, p), pl[3]);
}
// One row of Matrix4 * Vec4 with w = 1 (linalg.rs:346-360): (((0 + m0*x) + m1*y) + m2*z) + m3*1
__device__ __forceinline__ float mat_row(const float *m, v3 p) {
    return FADD(FADD(FADD(FADD(0.0f, FMUL(m[0], p.x)), FMUL(m[1], p.y)), FMUL(m[2], p.z)), m[3]);
}

// ------------------------------------------------------------------------------------------
// k_vertex
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vertex(const __grid_constant__ FrameUniforms U, const SceneDev S,
                                                const FrameDev W, const uint32_t n_tiles) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // per-frame reset of the binning state (k_setup runs after this kernel on the same stream)
    for (uint32_t t = i; t < n_tiles; t += gridDim.x * blockDim.x) W.tile_count[t] = 0;
    if (i < 3) W.counters[i] = 0;
    if (i >= S.n_vertices) return;

    const v3 p{S.px[i], S.py[i], S.pz[i]};
    const v3 cam{U.cam[0], U.cam[1], U.cam[2]};
    const v3 lsrc{U.light[0], U.light[1], U.light[2]};

    // scene/mod.rs:920-925
    const v3 eye_dir = v_sub(p, cam);
    const v3 lvec = v_sub(p, lsrc);
    const v3 light = v_div(lvec, v_norm(lvec));
    const float eye_len = v_norm(eye_dir);
    const v3 eye = v_div(eye_dir, eye_len);
    const v3 hsum = v_add(light, eye);
    const v3 halfway = v_div(hsum, v_norm(hsum));

    W.v_lx[i] = light.x; W.v_ly[i] = light.y; W.v_lz[i] = light.z;
    W.v_hx[i] = halfway.x; W.v_hy[i] = halfway.y; W.v_hz[i] = halfway.z;
    W.v_depth[i] = eye_len;

    // scene/mod.rs:1047-1058 for an unclipped corner: rows x, y, w of matrix_transf, then x/w, y/w
    const float cx = mat_row(&U.m[0], p);
    const float cy = mat_row(&U.m[4], p);
    const float cw = mat_row(&U.m[12], p);
    W.v_sx[i] = FDIV(cx, cw);
    W.v_sy[i] = FDIV(cy, cw);

    uint32_t flags = 0;
#pragma unroll
    for (int pl = 0; pl < 6; pl++) {
        const float f = plane_eval(U.planes[pl], p);
        flags |= (f > 0.0f ? 1u : 0u) << (2 * pl);
        flags |= (f <= 0.0f ? 1u : 0u) << (2 * pl + 1);
    }
    W.v_flags[i] = flags;
}

// ------------------------------------------------------------------------------------------
// triangle setup helpers
// ------------------------------------------------------------------------------------------

// canvas.rs:597-616 : f(x,y) = (((P.y-Q.y)*x + (Q.x-P.x)*y) + P.x*Q.y) - Q.x*P.y
struct Edge {
    float cx, cy, k1, k2;
};
__device__ __forceinline__ Edge make_edge(float px, float py, float qx, float qy) {
    return {FSUB(py, qy), FSUB(qx, px), FMUL(px, qy), FMUL(qx, py)};
}
__device__ __forceinline__ float edge_eval(const Edge &e, float x, float y) {
    return FSUB(FADD(FADD(FMUL(e.cx, x), FMUL(e.cy, y)), e.k1), e.k2);
}

// Rust `f32 as usize`: saturating, NaN -> 0 (cvt.rzi.u64.f32 has exactly these semantics)
__device__ __forceinline__ unsigned long long sat_usize(float v) { return __float2ull_rz(v); }

__device__ __forceinline__ float min3_ref(float a, float b, float c) { // canvas.rs:618-627
    float r = __int_as_float(0x7f800000);
    if (a < r) r = a;
    if (b < r) r = b;
    if (c < r) r = c;
    return r;
}
__device__ __forceinline__ float max3_ref(float a, float b, float c) { // canvas.rs:629-638
    float r = __int_as_float(0xff800000);
    if (a > r) r = a;
    if (b > r) r = b;
    if (c > r) r = c;
    return r;
}

// Rectangle::clip of two [min,max] ranges (canvas.rs:332-350), one axis.
__device__ __forceinline__ void clip_axis(unsigned long long a0, unsigned long long a1, unsigned long long b0,
                                          unsigned long long b1, unsigned long long &o0, unsigned long long &o1) {
    unsigned long long lo = a0 > b0 ? a0 : b0;
    unsigned long long hi = a1 < b1 ? a1 : b1;
    if (lo > hi) lo = hi = 0;
    o0 = lo < hi ? lo : hi; // from_coords normalisation (canvas.rs:315-330)
    o1 = lo < hi ? hi : lo;
}

// canvas.rs:585-666.  Builds the raster record of one screen triangle; returns false when the
// triangle provably writes nothing: one of f_alpha/f_beta/f_gama is zero or NaN, so every
// barycentric is +-inf or NaN, the interpolated depth is inf/NaN and `depth < stored` fails.
__device__ __forceinline__ bool setup_raster(const FrameUniforms &U, const float sx[3], const float sy[3],
                                             const float dep[3], uint32_t id, RasterRec &r) {
    // Vec2 sub is add of the negation (linalg.rs:37-43), then pos_map_center (canvas.rs:896-904)
    const float ax = floorf(FADD(FADD(sx[0], -U.off_x), 0.5f)), ay = floorf(FADD(FADD(sy[0], -U.off_y), 0.5f));
    const float bx = floorf(FADD(FADD(sx[1], -U.off_x), 0.5f)), by = floorf(FADD(FADD(sy[1], -U.off_y), 0.5f));
    const float cx = floorf(FADD(FADD(sx[2], -U.off_x), 0.5f)), cy = floorf(FADD(FADD(sy[2], -U.off_y), 0.5f));

    const Edge e_bc = make_edge(bx, by, cx, cy), e_ca = make_edge(cx, cy, ax, ay), e_ab = make_edge(ax, ay, bx, by);
    const float f_alpha = edge_eval(e_bc, ax, ay);
    const float f_beta = edge_eval(e_ca, bx, by);
    const float f_gama = edge_eval(e_ab, cx, cy);
    const bool nonzero = (f_alpha < 0.0f || f_alpha > 0.0f) && (f_beta < 0.0f || f_beta > 0.0f) &&
                         (f_gama < 0.0f || f_gama > 0.0f);
    if (!nonzero) return false;

    // canvas.rs:640-658
    unsigned long long x0 = sat_usize(min3_ref(ax, bx, cx)), y0 = sat_usize(min3_ref(ay, by, cy));
    unsigned long long x1 = sat_usize(max3_ref(ax, bx, cx)), y1 = sat_usize(max3_ref(ay, by, cy));
    if (x0 > x1) { unsigned long long t = x0; x0 = x1; x1 = t; }
    if (y0 > y1) { unsigned long long t = y0; y0 = y1; y1 = t; }
    const unsigned long long sw = U.canvas_w - 1, sh = U.canvas_h - 1;
    unsigned long long dx0, dx1, dy0, dy1;
    clip_axis(x0, x1, 0, sw, dx0, dx1); // clip(drawable, screen)
    clip_axis(y0, y1, 0, sh, dy0, dy1);
    clip_axis(0, sw, dx0, dx1, x0, x1); // clip(screen, drawable)
    clip_axis(0, sh, dy0, dy1, y0, y1);

    r.ax = ax; r.ay = ay; r.bx = bx; r.by = by; r.cx = cx; r.cy = cy;
    r.da = dep[0]; r.db = dep[1]; r.dc = dep[2];
    r.id = id;
    r.bbx = (uint32_t)x0 | ((uint32_t)x1 << 16);
    r.bby = (uint32_t)y0 | ((uint32_t)y1 << 16);
    return true;
}

// Adds the triangle to the count of every tile of this launch's stripe its bbox touches.
// Returns false if it touches none.
__device__ __forceinline__ bool count_tiles(const FrameUniforms &U, const FrameDev &W, const RasterRec &r) {
    const uint32_t tx0 = (r.bbx & 0xFFFF) / TILE, tx1 = (r.bbx >> 16) / TILE;
    uint32_t ty0 = (r.bby & 0xFFFF) / TILE, ty1 = (r.bby >> 16) / TILE;
    if (ty0 < U.tile_y_begin) ty0 = U.tile_y_begin;
    if (ty1 + 1 > U.tile_y_end) ty1 = U.tile_y_end - 1;
    if (U.tile_y_end == 0 || ty0 > ty1) return false;
    for (uint32_t ty = ty0; ty <= ty1; ty++)
        for (uint32_t tx = tx0; tx <= tx1; tx++) atomicAdd(&W.tile_count[ty * U.tiles_x + tx], 1u);
    return true;
}
__device__ __forceinline__ bool touches_stripe(const FrameUniforms &U, const RasterRec &r) {
    const uint32_t ty0 = (r.bby & 0xFFFF) / TILE, ty1 = (r.bby >> 16) / TILE;
    return ty1 >= U.tile_y_begin && ty0 < U.tile_y_end;
}

__device__ __forceinline__ void store_raster(RasterRec *dst, const RasterRec &r) {
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    d[0] = make_uint4(__float_as_uint(r.ax), __float_as_uint(r.ay), __float_as_uint(r.bx), __float_as_uint(r.by));
    d[1] = make_uint4(__float_as_uint(r.cx), __float_as_uint(r.cy), __float_as_uint(r.da), __float_as_uint(r.db));
    d[2] = make_uint4(__float_as_uint(r.dc), r.id, r.bbx, r.bby);
}
__device__ __forceinline__ RasterRec load_raster(const RasterRec *src) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    const uint4 a = __ldg(s), b = __ldg(s + 1), c = __ldg(s + 2);
    RasterRec r;
    r.ax = __uint_as_float(a.x); r.ay = __uint_as_float(a.y); r.bx = __uint_as_float(a.z); r.by = __uint_as_float(a.w);
    r.cx = __uint_as_float(b.x); r.cy = __uint_as_float(b.y); r.da = __uint_as_float(b.z); r.db = __uint_as_float(b.w);
    r.dc = __uint_as_float(c.x); r.id = c.y; r.bbx = c.z; r.bby = c.w;
    return r;
}

// ------------------------------------------------------------------------------------------
// clip path (rare): world-space near/far clipping, scene/mod.rs:43-90 and :662-746
// ------------------------------------------------------------------------------------------
constexpr int NATTR = 12; // depth, normal3, light3, halfway3, uv2 (uv.z and screen_coord are dead)
struct ClipVert {
    float p[3];
    float a[NATTR];
};
struct ClipTri {
    ClipVert v[3];
};

__device__ __forceinline__ v3 cv_pos(const ClipVert &v) { return v3{v.p[0], v.p[1], v.p[2]}; }

// a + (c - a) * t, component-wise, positions and attributes (scene/mod.rs:716-720, canvas.rs:242-291)
__device__ __forceinline__ void cv_lerp(const ClipVert &a, const ClipVert &c, float t, ClipVert &o) {
#pragma unroll
    for (int i = 0; i < 3; i++) o.p[i] = FADD(a.p[i], FMUL(FSUB(c.p[i], a.p[i]), t));
#pragma unroll
    for (int i = 0; i < NATTR; i++) o.a[i] = FADD(a.a[i], FMUL(FSUB(c.a[i], a.a[i]), t));
}

// ViewPlane::clip, scene/mod.rs:662-746
__device__ __noinline__ int clip_plane(const float *pl, const ClipTri &tri, ClipTri *out) {
    ClipVert a = tri.v[0], b = tri.v[1], c = tri.v[2];
    float f_a = plane_eval(pl, cv_pos(a)), f_b = plane_eval(pl, cv_pos(b)), f_c = plane_eval(pl, cv_pos(c));
    if (f_a > 0.0f && f_b > 0.0f && f_c > 0.0f) {
        out[0] = tri;
        return 1;
    }
    if (f_a <= 0.0f && f_b <= 0.0f && f_c <= 0.0f) return 0;
    if (FMUL(f_a, f_c) >= 0.0f) { // (a,b,c) <- (c,a,b)  :691-700
        ClipVert t = b; b = c; c = t;
        float ft = f_b; f_b = f_c; f_c = ft;
        t = a; a = b; b = t;
        ft = f_a; f_a = f_b; f_b = ft;
    } else if (FMUL(f_b, f_c) >= 0.0f) { // (a,b,c) <- (b,c,a)  :701-711
        ClipVert t = a; a = c; c = t;
        float ft = f_a; f_a = f_c; f_c = ft;
        t = a; a = b; b = t;
        ft = f_a; f_a = f_b; f_b = ft;
    }
    const v3 n{pl[0], pl[1], pl[2]};
    const float eps = 0.0000001f; // linalg.rs:6
    const float t_a = FSUB(FDIV(plane_eval(pl, cv_pos(a)), v_dot(n, v_sub(cv_pos(a), cv_pos(c)))), eps);
    ClipVert na;
    cv_lerp(a, c, t_a, na);
    const float t_b = FSUB(FDIV(plane_eval(pl, cv_pos(b)), v_dot(n, v_sub(cv_pos(b), cv_pos(c)))), eps);
    ClipVert nb;
    cv_lerp(b, c, t_b, nb);
    if (f_c <= 0.0f) { // :723-736
        out[0].v[0] = a; out[0].v[1] = na; out[0].v[2] = nb;
        out[1].v[0] = a; out[1].v[1] = b;  out[1].v[2] = nb;
        return 2;
    }
    out[0].v[0] = c; out[0].v[1] = na; out[0].v[2] = nb; // :737-745
    return 1;
}

__device__ __forceinline__ void shade_from_clip(const ClipTri &t, uint32_t material, ShadeRec &s) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            s.n[c][k] = t.v[c].a[1 + k];
            s.l[c][k] = t.v[c].a[4 + k];
            s.h[c][k] = t.v[c].a[7 + k];
        }
        s.uv[c][0] = t.v[c].a[10];
        s.uv[c][1] = t.v[c].a[11];
    }
    s.material = material;
    s.pad[0] = s.pad[1] = 0;
}

__device__ __forceinline__ void store_shade(ShadeRec *dst, const ShadeRec &s) {
    const uint4 *src = reinterpret_cast<const uint4 *>(&s);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(ShadeRec) / 16); i++) d[i] = src[i];
}

// Gathers a vertex' clip-space inputs (scene/mod.rs:938-1008).
__device__ __forceinline__ void gather_clip_vert(const SceneDev &S, const FrameDev &W, uint32_t v, uint32_t t,
                                                 uint32_t n, ClipVert &o) {
    o.p[0] = S.px[v]; o.p[1] = S.py[v]; o.p[2] = S.pz[v];
    o.a[0] = W.v_depth[v];
    o.a[1] = S.nx[n]; o.a[2] = S.ny[n]; o.a[3] = S.nz[n];
    o.a[4] = W.v_lx[v]; o.a[5] = W.v_ly[v]; o.a[6] = W.v_lz[v];
    o.a[7] = W.v_hx[v]; o.a[8] = W.v_hy[v]; o.a[9] = W.v_hz[v];
    o.a[10] = S.tu[t]; o.a[11] = S.tv[t];
}

// The full clip path for one triangle that straddles the near or far plane: up to 4 outputs,
// each projected (:1047-1063), set up, and written.  Opaque outputs take slots with a plain
// atomic (this path is rare); transparent outputs go to their ordered slots 4*ordinal + k.
__device__ __noinline__ void clip_and_emit(const FrameUniforms &U, const SceneDev &S, const FrameDev &W,
                                           uint32_t tri, const uint32_t vi[3], uint32_t material, bool transparent,
                                           uint32_t tslot) {
    ClipTri in;
#pragma unroll
    for (int c = 0; c < 3; c++) gather_clip_vert(S, W, vi[c], S.idx[3 + c][tri], S.idx[6 + c][tri], in.v[c]);

    ClipTri near_out[2], out[4];
    const int n_near = clip_plane(U.planes[0], in, near_out);
    int n_out = 0;
    for (int i = 0; i < n_near; i++) n_out += clip_plane(U.planes[1], near_out[i], out + n_out);

    for (int k = 0; k < 4; k++) {
        RasterRec r;
        bool keep = false;
        if (k < n_out) {
            float sx[3], sy[3], dep[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const v3 p = cv_pos(out[k].v[c]);
                const float w = mat_row(&U.m[12], p);
                sx[c] = FDIV(mat_row(&U.m[0], p), w);
                sy[c] = FDIV(mat_row(&U.m[4], p), w);
                dep[c] = out[k].v[c].a[0];
            }
            keep = setup_raster(U, sx, sy, dep, tri * 4u + (uint32_t)k, r);
        }
        if (transparent) {
            RasterRec *dst = W.t_rrec + (size_t)tslot * 4 + k;
            if (keep) {
                ShadeRec s;
                shade_from_clip(out[k], material, s);
                store_raster(dst, r);
                store_shade(W.t_srec + (size_t)tslot * 4 + k, s);
            } else {
                r.id = NO_SLOT;
                r.bbx = r.bby = 0;
                r.ax = r.ay = r.bx = r.by = r.cx = r.cy = r.da = r.db = r.dc = 0.0f;
                store_raster(dst, r);
            }
        } else if (keep && touches_stripe(U, r)) {
            const uint32_t slot = atomicAdd(&W.counters[0], 1u);
            if (slot >= W.rec_cap) {
                atomicOr(&W.counters[2], OVERFLOW_RECORDS);
                continue;
            }
            ShadeRec s;
            shade_from_clip(out[k], material, s);
            store_raster(W.rrec + slot, r);
            store_shade(W.srec + slot, s);
            count_tiles(U, W, r);
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_setup
// ------------------------------------------------------------------------------------------
constexpr int SETUP_THREADS = 256;

__global__ void __launch_bounds__(SETUP_THREADS) k_setup(const __grid_constant__ FrameUniforms U, const SceneDev S,
                                                         const FrameDev W) {
    __shared__ uint32_t warp_tot[SETUP_THREADS / 32];
    __shared__ uint32_t block_base;

    const uint32_t tri = blockIdx.x * SETUP_THREADS + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    bool emit = false;        // this thread has one unclipped opaque record to place
    bool transparent = false;
    uint32_t tslot = 0, material = 0;
    uint32_t vi[3] = {0, 0, 0};
    RasterRec r;
    r.id = NO_SLOT;

    if (tri < S.n_triangles) {
        vi[0] = S.idx[0][tri]; vi[1] = S.idx[1][tri]; vi[2] = S.idx[2][tri];
        const uint32_t mat = S.tri_mat[tri];
        material = mat & 0x7FFFFFFFu;
        transparent = (mat >> 31) != 0;
        if (transparent) tslot = S.tri_tslot[tri];

        bool alive = true;
        if (!transparent) {
            // back-face cull, scene/mod.rs:1016-1027 with calc_normal :30-41 and get_center :92-99
            const v3 a{S.px[vi[0]], S.py[vi[0]], S.pz[vi[0]]};
            const v3 b{S.px[vi[1]], S.py[vi[1]], S.pz[vi[1]]};
            const v3 c{S.px[vi[2]], S.py[vi[2]], S.pz[vi[2]]};
            const v3 nrm = v_cross(v_sub(b, a), v_sub(c, b));
            v3 sum{0.0f, 0.0f, 0.0f};
            sum = v_add(sum, a);
            sum = v_add(sum, b);
            sum = v_add(sum, c);
            const v3 center = v_div(sum, 3.0f);
            const v3 eye = v_sub(v3{U.cam[0], U.cam[1], U.cam[2]}, center);
            if (v_dot(eye, nrm) <= 0.0f) alive = false;
        }
        bool clip = false;
        if (alive) {
            const uint32_t fa = W.v_flags[vi[0]], fb = W.v_flags[vi[1]], fc = W.v_flags[vi[2]];
            const uint32_t all_nonpos = fa & fb & fc & 0xAAAu; // bit 2p+1 : f <= 0 on all three
            const uint32_t all_pos = fa & fb & fc & 0x555u;    // bit 2p   : f > 0 on all three
            // lateral planes 2..5: reject only if completely outside one of them (:59-66, :641-660)
            if (all_nonpos & 0xAA0u) alive = false;
            else if ((all_pos & 0x5u) == 0x5u) clip = false;   // inside near and far: passes through (:677-681)
            else if (all_nonpos & 0x2u) alive = false;          // completely behind the near plane (:682-686)
            else if ((all_pos & 0x1u) && (all_nonpos & 0x8u)) alive = false; // untouched by near, beyond far
            else clip = true;
        }
        if (alive && clip) {
            clip_and_emit(U, S, W, tri, vi, material, transparent, tslot);
            alive = false;
            if (transparent) transparent = false; // its four ordered slots were written by clip_and_emit
        } else if (alive) {
            const float sx[3] = {W.v_sx[vi[0]], W.v_sx[vi[1]], W.v_sx[vi[2]]};
            const float sy[3] = {W.v_sy[vi[0]], W.v_sy[vi[1]], W.v_sy[vi[2]]};
            const float dep[3] = {W.v_depth[vi[0]], W.v_depth[vi[1]], W.v_depth[vi[2]]};
            alive = setup_raster(U, sx, sy, dep, tri * 4u, r);
            if (!alive) r.id = NO_SLOT;
        }
        emit = alive && !transparent && touches_stripe(U, r);
        if (transparent) emit = false;
        if (transparent && !alive) r.id = NO_SLOT;
    }

    // block-wide slot allocation: ballot + popc inside the warp, one atomic per block
    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, emit);
    const uint32_t warp_rank = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int w = 0; w < SETUP_THREADS / 32; w++) {
            const uint32_t c = warp_tot[w];
            warp_tot[w] = total;
            total += c;
        }
        block_base = total ? atomicAdd(&W.counters[0], total) : 0u;
    }
    __syncthreads();

    RasterRec *rdst = nullptr;
    ShadeRec *sdst = nullptr;
    if (emit) {
        const uint32_t slot = block_base + warp_tot[warp] + warp_rank;
        if (slot < W.rec_cap) {
            rdst = W.rrec + slot;
            sdst = W.srec + slot;
            count_tiles(U, W, r);
        } else {
            atomicOr(&W.counters[2], OVERFLOW_RECORDS);
        }
    } else if (transparent) {
        // unclipped transparent triangle: ordered slot 4*ordinal, the other three are empty
        RasterRec empty;
        empty.id = NO_SLOT;
        empty.bbx = empty.bby = 0;
        empty.ax = empty.ay = empty.bx = empty.by = empty.cx = empty.cy = empty.da = empty.db = empty.dc = 0.0f;
        for (int k = 1; k < 4; k++) store_raster(W.t_rrec + (size_t)tslot * 4 + k, empty);
        rdst = W.t_rrec + (size_t)tslot * 4;
        if (r.id == NO_SLOT) {
            store_raster(rdst, empty);
            rdst = nullptr;
        } else {
            sdst = W.t_srec + (size_t)tslot * 4;
        }
    }
    if (rdst) {
        store_raster(rdst, r);
        // attribute gather for the shading record (scene/mod.rs:938-1008)
        ShadeRec s;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const uint32_t v = vi[c], t = S.idx[3 + c][tri], n = S.idx[6 + c][tri];
            s.n[c][0] = S.nx[n]; s.n[c][1] = S.ny[n]; s.n[c][2] = S.nz[n];
            s.l[c][0] = W.v_lx[v]; s.l[c][1] = W.v_ly[v]; s.l[c][2] = W.v_lz[v];
            s.h[c][0] = W.v_hx[v]; s.h[c][1] = W.v_hy[v]; s.h[c][2] = W.v_hz[v];
            s.uv[c][0] = S.tu[t]; s.uv[c][1] = S.tv[t];
        }
        s.material = material;
        s.pad[0] = s.pad[1] = 0;
        store_shade(sdst, s);
    }
}

// ------------------------------------------------------------------------------------------
// k_scan : exclusive scan of tile_count -> tile_offset, cursors reset, total -> counters[1]
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(SCAN_THREADS) k_scan(const FrameDev W, const uint32_t n_tiles) {
    __shared__ uint32_t warp_sum[SCAN_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (n_tiles + SCAN_THREADS - 1) / SCAN_THREADS;
    const uint32_t begin = tid * per;
    const uint32_t end = begin + per < n_tiles ? begin + per : n_tiles;

    uint32_t local = 0;
    for (uint32_t i = begin; i < end; i++) local += W.tile_count[i];

    uint32_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (uint32_t)d) incl += up;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sum[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= (uint32_t)d) wi += up;
        }
        warp_sum[lane] = wi - w; // exclusive
    }
    __syncthreads();
    uint32_t run = warp_sum[warp] + incl - local;
    for (uint32_t i = begin; i < end; i++) {
        const uint32_t c = W.tile_count[i];
        W.tile_offset[i] = run;
        W.tile_count[i] = 0; // becomes the fill cursor
        run += c;
    }
    if (tid == SCAN_THREADS - 1) {
        W.tile_offset[n_tiles] = run;
        W.counters[1] = run;
        if (run > W.refs_cap) atomicOr(&W.counters[2], OVERFLOW_REFS);
    }
}

// ------------------------------------------------------------------------------------------
// k_fill : scatter record slots into the per-tile lists (order inside a list is irrelevant)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill(const __grid_constant__ FrameUniforms U, const FrameDev W) {
    if (W.counters[2] != 0) return; // a buffer overflowed: the host re-renders with larger buffers
    uint32_t n = W.counters[0];
    if (n > W.rec_cap) n = W.rec_cap;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        const uint4 c = __ldg(reinterpret_cast<const uint4 *>(W.rrec + slot) + 2);
        const uint32_t bbx = c.z, bby = c.w;
        const uint32_t tx0 = (bbx & 0xFFFF) / TILE, tx1 = (bbx >> 16) / TILE;
        uint32_t ty0 = (bby & 0xFFFF) / TILE, ty1 = (bby >> 16) / TILE;
        if (ty0 < U.tile_y_begin) ty0 = U.tile_y_begin;
        if (ty1 + 1 > U.tile_y_end) ty1 = U.tile_y_end - 1;
        for (uint32_t ty = ty0; ty <= ty1; ty++)
            for (uint32_t tx = tx0; tx <= tx1; tx++) {
                const uint32_t tile = ty * U.tiles_x + tx;
                const uint32_t pos = atomicAdd(&W.tile_count[tile], 1u);
                W.tile_refs[W.tile_offset[tile] + pos] = slot;
            }
    }
}

// ------------------------------------------------------------------------------------------
// k_tile
// ------------------------------------------------------------------------------------------
constexpr int TILE_THREADS = 256; // 8 warps; warp = 32x16 px region; thread = 4x4 px block
constexpr int CHUNK = 64;         // triangles staged in shared memory per round

// One triangle prepared for the pixel loop.  In fast mode the three edge functions are
// sign-normalised (all coefficients negated when f < 0; exact, round-to-nearest is symmetric) so
// that f > 0 and "alpha >= 0" reads "e >= 0".
struct Staged {
    float ecx[3][CHUNK], ecy[3][CHUNK], ek1[3][CHUNK], ek2[3][CHUNK], f[3][CHUNK];
    float da[CHUNK], db[CHUNK], dc[CHUNK];
    int x0[CHUNK], x1[CHUNK], y0[CHUNK], y1[CHUNK];
    uint32_t flags[CHUNK]; // bit e: tie rule of edge e admits e == 0 ; bit 3: slow (literal) mode
    uint32_t id[CHUNK], slot[CHUNK];
};

__device__ __forceinline__ void stage_triangle(Staged &st, int k, const RasterRec &r, uint32_t slot) {
    // edge order: 0 = bc (alpha, vertex a), 1 = ca (beta, b), 2 = ab (gama, c)   canvas.rs:660-666
    Edge e[3] = {make_edge(r.bx, r.by, r.cx, r.cy), make_edge(r.cx, r.cy, r.ax, r.ay),
                 make_edge(r.ax, r.ay, r.bx, r.by)};
    const float vx[3] = {r.ax, r.bx, r.cx}, vy[3] = {r.ay, r.by, r.cy};
    uint32_t flags = 0;
    bool tame = true;
    float f[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        f[i] = edge_eval(e[i], vx[i], vy[i]);
        const float f_out = edge_eval(e[i], -1.0f, -1.0f);
        if (FMUL(f[i], f_out) > 0.0f) flags |= 1u << i; // canvas.rs:678-680
        const float lim = 1e30f;
        tame = tame && fabsf(e[i].cx) < lim && fabsf(e[i].cy) < lim && fabsf(e[i].k1) < lim &&
               fabsf(e[i].k2) < lim && fabsf(f[i]) < lim;
    }
    if (!tame) flags |= 8u;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const bool neg = tame && f[i] < 0.0f;
        st.ecx[i][k] = neg ? -e[i].cx : e[i].cx;
        st.ecy[i][k] = neg ? -e[i].cy : e[i].cy;
        st.ek1[i][k] = neg ? -e[i].k1 : e[i].k1;
        st.ek2[i][k] = neg ? -e[i].k2 : e[i].k2;
        st.f[i][k] = neg ? -f[i] : f[i];
    }
    st.da[k] = r.da; st.db[k] = r.db; st.dc[k] = r.dc;
    st.x0[k] = (int)(r.bbx & 0xFFFF); st.x1[k] = (int)(r.bbx >> 16);
    st.y0[k] = (int)(r.bby & 0xFFFF); st.y1[k] = (int)(r.bby >> 16);
    st.flags[k] = flags;
    st.id[k] = r.id;
    st.slot[k] = slot;
}

// Inside test + barycentrics of one pixel against staged triangle k (canvas.rs:673-682).
// Returns true and the interpolated depth when the pixel is covered.
__device__ __forceinline__ bool cover_fast(const Staged &st, int k, uint32_t flags, float px0, float py0, float px1,
                                           float py1, float px2, float py2, float &depth) {
    const float e0 = FSUB(FADD(FADD(px0, py0), st.ek1[0][k]), st.ek2[0][k]);
    const float e1 = FSUB(FADD(FADD(px1, py1), st.ek1[1][k]), st.ek2[1][k]);
    const float e2 = FSUB(FADD(FADD(px2, py2), st.ek1[2][k]), st.ek2[2][k]);
    // f > 0 and finite, e integer-valued: e/f >= 0 <=> e >= 0, e/f > 0 <=> e > 0 (no underflow)
    const bool in = (e0 > 0.0f || (e0 == 0.0f && (flags & 1u))) && (e1 > 0.0f || (e1 == 0.0f && (flags & 2u))) &&
                    (e2 > 0.0f || (e2 == 0.0f && (flags & 4u)));
    if (!in) return false;
    const float alpha = FDIV(e0, st.f[0][k]), beta = FDIV(e1, st.f[1][k]), gama = FDIV(e2, st.f[2][k]);
    depth = FADD(FADD(FMUL(alpha, st.da[k]), FMUL(beta, st.db[k])), FMUL(gama, st.dc[k]));
    return true;
}
__device__ __noinline__ bool cover_slow(const Staged &st, int k, uint32_t flags, float x, float y, float &depth) {
    float bary[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float e = FSUB(FADD(FADD(FMUL(st.ecx[i][k], x), FMUL(st.ecy[i][k], y)), st.ek1[i][k]), st.ek2[i][k]);
        bary[i] = FDIV(e, st.f[i][k]);
    }
    if (!(bary[0] >= 0.0f && bary[1] >= 0.0f && bary[2] >= 0.0f)) return false;
    if (!((bary[0] > 0.0f || (flags & 1u)) && (bary[1] > 0.0f || (flags & 2u)) && (bary[2] > 0.0f || (flags & 4u))))
        return false;
    depth = FADD(FADD(FMUL(bary[0], st.da[k]), FMUL(bary[1], st.db[k])), FMUL(bary[2], st.dc[k]));
    return true;
}

// TextureMap::get_rgb_slice (scene/mod.rs:154-168) + Pixel::normalized_as_vec3 (canvas.rs:81-87).
// Indices are clamped into the map (SURVEY.md §8c deviation 6; never triggers for uv in [0,1]).
__device__ __forceinline__ v3 fetch_texel(const uint8_t *__restrict__ texels, uint32_t off, uint32_t w, uint32_t h,
                                          uint32_t comp, float u, float v) {
    unsigned long long ui = sat_usize(floorf(FMUL(u, FSUB((float)w, 1.0f))));
    unsigned long long vr = sat_usize(floorf(FMUL(v, FSUB((float)h, 1.0f))));
    if (ui > w - 1) ui = w - 1;
    if (vr > h - 1) vr = h - 1;
    const uint8_t *p = texels + off + ((size_t)(h - 1 - (uint32_t)vr) * w + (uint32_t)ui) * comp;
    return v3{FDIV((float)__ldg(p), 255.0f), FDIV((float)__ldg(p + 1), 255.0f), FDIV((float)__ldg(p + 2), 255.0f)};
}

// Rust `f32 as u8` (saturating, NaN -> 0)
__device__ __forceinline__ uint32_t sat_u8(float v) {
    const uint32_t u = __float2uint_rz(v);
    return u > 255u ? 255u : u;
}

// canvas.rs:673-743 for one covered pixel: literal barycentrics, interpolation, texel fetches and
// Phong.  Returns r | g << 8 | b << 16; *depth_out gets the interpolated depth.
__device__ __forceinline__ uint32_t shade_pixel(const SceneDev &S, const RasterRec &r, const ShadeRec *__restrict__ sp,
                                                float x, float y, float *depth_out, float *opacity_out) {
    const Edge e_bc = make_edge(r.bx, r.by, r.cx, r.cy), e_ca = make_edge(r.cx, r.cy, r.ax, r.ay),
               e_ab = make_edge(r.ax, r.ay, r.bx, r.by);
    const float alpha = FDIV(edge_eval(e_bc, x, y), edge_eval(e_bc, r.ax, r.ay));
    const float beta = FDIV(edge_eval(e_ca, x, y), edge_eval(e_ca, r.bx, r.by));
    const float gama = FDIV(edge_eval(e_ab, x, y), edge_eval(e_ab, r.cx, r.cy));
    *depth_out = FADD(FADD(FMUL(alpha, r.da), FMUL(beta, r.db)), FMUL(gama, r.dc));

    // ShadeRec as 9 x uint4: n[3][3] l[3][3] h[3][3] uv[3][2] material pad pad
    const uint4 *q = reinterpret_cast<const uint4 *>(sp);
    float w[36];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const uint4 t = __ldg(q + i);
        w[4 * i] = __uint_as_float(t.x); w[4 * i + 1] = __uint_as_float(t.y);
        w[4 * i + 2] = __uint_as_float(t.z); w[4 * i + 3] = __uint_as_float(t.w);
    }
    // X = ((Xa*alpha) + (Xb*beta)) + (Xc*gama)   canvas.rs:685-722
#define INTERP(base, comp, stride) \
    FADD(FADD(FMUL(w[(base) + (comp)], alpha), FMUL(w[(base) + (stride) + (comp)], beta)), \
         FMUL(w[(base) + 2 * (stride) + (comp)], gama))
    const v3 N{INTERP(0, 0, 3), INTERP(0, 1, 3), INTERP(0, 2, 3)};
    const v3 L{INTERP(9, 0, 3), INTERP(9, 1, 3), INTERP(9, 2, 3)};
    const v3 H{INTERP(18, 0, 3), INTERP(18, 1, 3), INTERP(18, 2, 3)};
    const float u = INTERP(27, 0, 2), v = INTERP(27, 1, 2);
#undef INTERP
    const uint32_t material = __float_as_uint(w[33]);
    const MaterialDev *m = S.materials + material;
    const v3 ka{__ldg(&m->ka[0]), __ldg(&m->ka[1]), __ldg(&m->ka[2])};
    const v3 kd{__ldg(&m->kd[0]), __ldg(&m->kd[1]), __ldg(&m->kd[2])};
    const v3 ks{__ldg(&m->ks[0]), __ldg(&m->ks[1]), __ldg(&m->ks[2])};
    *opacity_out = __ldg(&m->alpha);

    const v3 dcol = fetch_texel(S.texels, __ldg(&m->kd_off), __ldg(&m->kd_w), __ldg(&m->kd_h), __ldg(&m->kd_comp), u, v);
    const v3 acol = fetch_texel(S.texels, __ldg(&m->ka_off), __ldg(&m->ka_w), __ldg(&m->ka_h), __ldg(&m->ka_comp), u, v);

    // canvas.rs:732-739
    const v3 c_r{FMUL(dcol.x, kd.x), FMUL(dcol.y, kd.y), FMUL(dcol.z, kd.z)};
    const v3 c_a{FMUL(acol.x, ka.x), FMUL(acol.y, ka.y), FMUL(acol.z, ka.z)};
    const float ln = v_dot(L, N);
    const float s = FSUB(1.0f, ln > 0.0f ? ln : 0.0f); // 0.0_f32.max(x): NaN -> 0
    const float hn = v_dot(H, N);
    const float spec = FMUL(hn, hn); // powi(2)
    const float cr = FADD(FMUL(c_r.x, FADD(c_a.x, FMUL(ks.x, s))), FMUL(ks.x, spec));
    const float cg = FADD(FMUL(c_r.y, FADD(c_a.y, FMUL(ks.y, s))), FMUL(ks.y, spec));
    const float cb = FADD(FMUL(c_r.z, FADD(c_a.z, FMUL(ks.z, s))), FMUL(ks.z, spec));
    // Pixel::from_normalized_vec3, canvas.rs:89-92
    return sat_u8(FMUL(cr, 255.0f)) | (sat_u8(FMUL(cg, 255.0f)) << 8) | (sat_u8(FMUL(cb, 255.0f)) << 16);
}

// Pixel * f32 + Pixel * f32 (canvas.rs:136-169, :916-921): per channel truncate, u8 wrapping add, pad 0.
// colours are r | g << 8 | b << 16.
__device__ __forceinline__ uint32_t blend_rgb(uint32_t bg, uint32_t fg, float opacity) {
    const float k0 = FSUB(1.0f, opacity);
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float b = (float)((bg >> (8 * c)) & 255u), f = (float)((fg >> (8 * c)) & 255u);
        out |= ((sat_u8(FMUL(b, k0)) + sat_u8(FMUL(f, opacity))) & 255u) << (8 * c);
    }
    return out;
}

__global__ void __launch_bounds__(TILE_THREADS) k_tile(const __grid_constant__ FrameUniforms U, const SceneDev S,
                                                       const FrameDev W, uint8_t *__restrict__ color,
                                                       float *__restrict__ depth) {
    __shared__ Staged st;

    const uint32_t tile_x = blockIdx.x % U.tiles_x;
    const uint32_t tile_y = U.tile_y_begin + blockIdx.x / U.tiles_x;
    const uint32_t tile = tile_y * U.tiles_x + tile_x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // thread -> its 4x4 pixel block (canvas coordinates: x right, y = depth-buffer row)
    const int wx0 = (int)tile_x * TILE + (warp & 1) * 32, wy0 = (int)tile_y * TILE + (warp >> 1) * 16;
    const int bx0 = wx0 + (lane & 7) * 4, by0 = wy0 + (lane >> 3) * 4;

    float zb[16];
    uint32_t sl[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        zb[i] = U.depth_max;
        sl[i] = NO_SLOT;
    }

    const bool usable = W.counters[2] == 0;
    const uint32_t list_begin = usable ? W.tile_offset[tile] : 0u;
    const uint32_t list_end = usable ? W.tile_offset[tile + 1] : 0u;

    // ---- phase 1: opaque triangles of this tile, any order, min over (depth, draw id) -------
    for (uint32_t base = list_begin; base < list_end; base += CHUNK) {
        const int n = (int)min((uint32_t)CHUNK, list_end - base);
        __syncthreads();
        if (tid < n) {
            const uint32_t slot = W.tile_refs[base + tid];
            stage_triangle(st, tid, load_raster(W.rrec + slot), slot);
        }
        __syncthreads();
        for (int k = 0; k < n; k++) {
            // warp-level bbox reject (uniform)
            if (st.x1[k] < wx0 || st.x0[k] > wx0 + 31 || st.y1[k] < wy0 || st.y0[k] > wy0 + 15) continue;
            const int lo_x = max(st.x0[k], bx0), hi_x = min(st.x1[k], bx0 + 3);
            const int lo_y = max(st.y0[k], by0), hi_y = min(st.y1[k], by0 + 3);
            if (lo_x > hi_x || lo_y > hi_y) continue;
            const uint32_t flags = st.flags[k];
            const uint32_t id = st.id[k], slot = st.slot[k];
            if (!(flags & 8u)) {
                // Block-level reject.  Each edge value is a monotone function of x and of y (every
                // rounding step is monotone), so its maximum over the clipped block sits at a corner
                // chosen by the coefficient signs; if even that corner fails, every pixel fails.
                bool any = true;
#pragma unroll
                for (int e = 0; e < 3; e++) {
                    const float cx = st.ecx[e][k], cy = st.ecy[e][k];
                    const float xm = (float)(cx >= 0.0f ? hi_x : lo_x), ym = (float)(cy >= 0.0f ? hi_y : lo_y);
                    const float em = FSUB(FADD(FADD(FMUL(cx, xm), FMUL(cy, ym)), st.ek1[e][k]), st.ek2[e][k]);
                    any = any && (em > 0.0f || (em == 0.0f && (flags & (1u << e))));
                }
                if (!any) continue;
                float pxs[3][4], pys[3][4];
#pragma unroll
                for (int e = 0; e < 3; e++) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        pxs[e][i] = FMUL(st.ecx[e][k], (float)(bx0 + i));
                        pys[e][i] = FMUL(st.ecy[e][k], (float)(by0 + i));
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int x = bx0 + i, y = by0 + j;
                        if (x < lo_x || x > hi_x || y < lo_y || y > hi_y) continue;
                        float d;
                        if (!cover_fast(st, k, flags, pxs[0][i], pys[0][j], pxs[1][i], pys[1][j], pxs[2][i], pys[2][j], d))
                            continue;
                        const int p = j * 4 + i;
                        if (d < zb[p]) {
                            zb[p] = d;
                            sl[p] = slot;
                        } else if (d == zb[p] && sl[p] != NO_SLOT) {
                            if (id < __ldg(&W.rrec[sl[p]].id)) { // equal depth: the earlier draw wins
                                zb[p] = d;
                                sl[p] = slot;
                            }
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int x = bx0 + i, y = by0 + j;
                        if (x < lo_x || x > hi_x || y < lo_y || y > hi_y) continue;
                        float d;
                        if (!cover_slow(st, k, flags, (float)x, (float)y, d)) continue;
                        const int p = j * 4 + i;
                        if (d < zb[p]) {
                            zb[p] = d;
                            sl[p] = slot;
                        } else if (d == zb[p] && sl[p] != NO_SLOT) {
                            if (id < __ldg(&W.rrec[sl[p]].id)) {
                                zb[p] = d;
                                sl[p] = slot;
                            }
                        }
                    }
                }
            }
        }
    }

    // ---- deferred shading of the opaque winners, fused clear -----------------------------------
    // colour as r | g << 8 | b << 16 | pad << 24 ; clear = azul_bb (155,186,255), pad 255 (canvas.rs:131)
    uint32_t col[16];
    uint32_t wid[16]; // draw id of the opaque winner (for the transparent phase), NO_SLOT = none
#pragma unroll
    for (int p = 0; p < 16; p++) {
        col[p] = 155u | (186u << 8) | (255u << 16) | (255u << 24);
        wid[p] = NO_SLOT;
        if (sl[p] != NO_SLOT) {
            const RasterRec r = load_raster(W.rrec + sl[p]);
            float d, op;
            const uint32_t rgb = shade_pixel(S, r, W.srec + sl[p], (float)(bx0 + (p & 3)), (float)(by0 + (p >> 2)), &d, &op);
            col[p] = rgb | (255u << 24);
            wid[p] = r.id;
        }
    }

    // ---- phase 2: transparent triangles in draw order (scene/mod.rs:1088-1246) -----------------
    const uint32_t n_tslots = usable ? S.n_transparent * 4u : 0u;
    for (uint32_t base = 0; base < n_tslots; base += CHUNK) {
        const int n = (int)min((uint32_t)CHUNK, n_tslots - base);
        __syncthreads();
        if (tid < n) {
            RasterRec r = load_raster(W.t_rrec + base + tid);
            if (r.id == NO_SLOT) { // empty slot: give it an empty bbox so every thread skips it
                r.bbx = 1u;        // x_min = 1 > x_max = 0
                r.bby = 1u;
            }
            stage_triangle(st, tid, r, base + tid);
        }
        __syncthreads();
        for (int k = 0; k < n; k++) {
            if (st.x1[k] < wx0 || st.x0[k] > wx0 + 31 || st.y1[k] < wy0 || st.y0[k] > wy0 + 15) continue;
            const int lo_x = max(st.x0[k], bx0), hi_x = min(st.x1[k], bx0 + 3);
            const int lo_y = max(st.y0[k], by0), hi_y = min(st.y1[k], by0 + 3);
            if (lo_x > hi_x || lo_y > hi_y) continue;
            const uint32_t flags = st.flags[k] | 8u; // literal evaluation; this phase is not the hot one
            const uint32_t id = st.id[k], slot = st.slot[k];
#pragma unroll
            for (int p = 0; p < 16; p++) {
                const int x = bx0 + (p & 3), y = by0 + (p >> 2);
                if (x < lo_x || x > hi_x || y < lo_y || y > hi_y) continue;
                float d;
                // staged coefficients may be sign-normalised; the quotient e/f is unchanged by that
                if (!cover_slow(st, k, flags, (float)x, (float)y, d)) continue;
                // drawn after the opaque winner (or no winner) and in front of the final depth
                if (!(wid[p] == NO_SLOT || id > wid[p])) continue;
                if (!(d < zb[p])) continue;
                const RasterRec r = load_raster(W.t_rrec + slot);
                float d2, op;
                const uint32_t rgb = shade_pixel(S, r, W.t_srec + slot, (float)x, (float)y, &d2, &op);
                // canvas.rs:916-921: opacity < 1 blends with the stored colour, else replaces it
                col[p] = op < 1.0f ? blend_rgb(col[p], rgb, op) : (rgb | (255u << 24));
            }
        }
    }

    // ---- single write-back: colour rows are y-flipped (canvas.rs:955-956), depth rows are not ----
    const int W_ = (int)U.canvas_w, H_ = (int)U.canvas_h;
    const bool vec_ok = (W_ & 3) == 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int y = by0 + j;
        if (y >= H_ || bx0 >= W_) continue;
        uint32_t px[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t c = col[j * 4 + i]; // r g b pad -> memory order b g r pad
            px[i] = ((c >> 16) & 255u) | (c & 0x0000FF00u) | ((c & 255u) << 16) | (c & 0xFF000000u);
        }
        const size_t crow = (size_t)(H_ - 1 - y) * W_ + bx0, drow = (size_t)y * W_ + bx0;
        if (vec_ok) {
            *reinterpret_cast<uint4 *>(color + crow * 4) = make_uint4(px[0], px[1], px[2], px[3]);
            *reinterpret_cast<float4 *>(depth + drow) = make_float4(zb[j * 4], zb[j * 4 + 1], zb[j * 4 + 2], zb[j * 4 + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (bx0 + i < W_) {
                    reinterpret_cast<uint32_t *>(color)[crow + i] = px[i];
                    depth[drow + i] = zb[j * 4 + i];
                }
            }
        }
    }
}

// Canvas::clear (canvas.rs:425-433) as a standalone operation (draw_canvas_clear).
__global__ void __launch_bounds__(256) k_clear(uint32_t *__restrict__ color, float *__restrict__ depth, size_t n,
                                               float depth_max, int has_depth) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        color[i] = 255u | (186u << 8) | (155u << 16) | (255u << 24); // memory order b g r pad
        if (has_depth) depth[i] = depth_max;
    }
}
__global__ void __launch_bounds__(256) k_fill_u32(uint32_t *__restrict__ dst, size_t n, uint32_t value) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = value;
}

// ------------------------------------------------------------------------------------------
// launchers (called from scene.cpp)
// ------------------------------------------------------------------------------------------
cudaError_t launch_frame(const FrameUniforms &U, const SceneDev &S, const FrameDev &W, uint8_t *color, float *depth,
                         cudaStream_t stream, uint64_t *launches) {
    const uint32_t n_tiles = U.tiles_x * U.tiles_y;
    const uint32_t vblocks = S.n_vertices ? (S.n_vertices + 255) / 256 : 1;
    k_vertex<<<vblocks, 256, 0, stream>>>(U, S, W, n_tiles);
    if (S.n_triangles) {
        k_setup<<<(S.n_triangles + SETUP_THREADS - 1) / SETUP_THREADS, SETUP_THREADS, 0, stream>>>(U, S, W);
        ++*launches;
    }
    k_scan<<<1, SCAN_THREADS, 0, stream>>>(W, n_tiles);
    int fill_blocks = (int)((W.rec_cap + 255) / 256);
    if (fill_blocks > 148 * 8) fill_blocks = 148 * 8;
    if (fill_blocks < 1) fill_blocks = 1;
    k_fill<<<fill_blocks, 256, 0, stream>>>(U, W);
    const uint32_t stripe_tiles = (U.tile_y_end - U.tile_y_begin) * U.tiles_x;
    if (stripe_tiles) {
        k_tile<<<stripe_tiles, TILE_THREADS, 0, stream>>>(U, S, W, color, depth);
        ++*launches;
    }
    *launches += 3;
    return cudaGetLastError();
}

cudaError_t launch_clear(uint8_t *color, float *depth, size_t n_pixels, float depth_max, cudaStream_t stream,
                         uint64_t *launches) {
    k_clear<<<148 * 4, 256, 0, stream>>>(reinterpret_cast<uint32_t *>(color), depth, n_pixels, depth_max, depth != nullptr);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_fill_u32(uint32_t *dst, size_t n, uint32_t value, cudaStream_t stream, uint64_t *launches) {
    k_fill_u32<<<148 * 4, 256, 0, stream>>>(dst, n, value);
    ++*launches;
    return cudaGetLastError();
}

} // namespace drawb200
