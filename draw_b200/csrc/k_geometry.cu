// k_geometry.cu — per-vertex and per-triangle stages of the frame (mororo18/draw scene/mod.rs:901).
//
//   k_vertex   per vertex   light / halfway / depth (scene/mod.rs:917-926), screen xy of the
//                           unclipped vertex (:1047-1058), view-plane side flags (:634-660)
//   k_setup    per triangle gather, back-face cull (:1016-1027), lateral reject + near/far clip
//                           (:43-90, :662-746), snap + bbox + zero-area cull (canvas.rs:585-666),
//                           draw-order-preserving record allocation (warp prefix sums + chained
//                           scan across CTAs), record write; triangles that need clipping are queued
//   k_clip     per queued triangle: the clip path (rare, register-hungry), fills the reserved slots
//
// Arithmetic contract: see device_math.cuh.
#include "device_math.cuh"

namespace drawb200 {

// ------------------------------------------------------------------------------------------
// k_vertex
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vertex(const FrameUniforms *__restrict__ Up, const SceneDev S,
                                                const FrameDev W) {
    const FrameUniforms &U = *Up; // per-frame uniforms, device-resident (one upload per frame; the launches never change)
    const CtaTrace trace_(W, 0u);
    pdl_prologue(U.pdl_early != 0);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // per-frame reset of the binning state (the binning kernels run after this one on the same stream)
    for (uint32_t t = i; t < U.n_lists; t += gridDim.x * blockDim.x) W.list_count[t] = 0;
    for (uint32_t t = i; t < 2 * U.n_coarse; t += gridDim.x * blockDim.x) W.tile_cost[t] = 0;
    if (i < N_COUNTERS) W.counters[i] = 0;
    const uint32_t n_desc = (S.n_triangles + 255) / 256;
    for (uint32_t t = i; t < n_desc; t += gridDim.x * blockDim.x) W.scan_desc[t] = 0ull;
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

// Does the bbox reach any coarse tile row of this launch's stripe?  (Records that do not are dropped.)
__device__ __forceinline__ bool touches_stripe(const FrameUniforms &U, const RasterRec &r) {
    const uint32_t ty0 = (r.bby & 0xFFFF) / TILE_H, ty1 = (r.bby >> 16) / TILE_H;
    return ty1 >= U.tile_y_begin && ty0 < U.tile_y_end;
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

// The full clip path for one triangle that straddles the near or far plane: up to 4 outputs, each
// projected (:1047-1063) and set up.  Transparent outputs are written straight to their ordered
// slots 4*ordinal + k (empty slots are marked).  Opaque outputs that survive are returned
// compacted in out_r / out_s (in emission order) for the caller to place once its slots are known.
__device__ __noinline__ int clip_triangle(const FrameUniforms &U, const SceneDev &S, const FrameDev &W, uint32_t tri,
                                          const uint32_t vi[3], uint32_t material, bool transparent, uint32_t tslot,
                                          RasterRec *out_r, ShadeRec *out_s) {
    ClipTri in;
#pragma unroll
    for (int c = 0; c < 3; c++) gather_clip_vert(S, W, vi[c], S.idx[3 + c][tri], S.idx[6 + c][tri], in.v[c]);

    ClipTri near_out[2], out[4];
    const int n_near = clip_plane(U.planes[0], in, near_out);
    int n_out = 0;
    for (int i = 0; i < n_near; i++) n_out += clip_plane(U.planes[1], near_out[i], out + n_out);

    int n_keep = 0;
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
            out_r[n_keep] = r;
            shade_from_clip(out[k], material, out_s[n_keep]);
            n_keep++;
        }
    }
    return n_keep;
}

// ------------------------------------------------------------------------------------------
// k_setup
// ------------------------------------------------------------------------------------------
constexpr int SETUP_THREADS = 256;

// Descriptor of the chained scan over CTAs: status in the top 2 bits, count in the rest.
#define DESC_AGGREGATE (1ull << 62)
#define DESC_PREFIX (2ull << 62)
#define DESC_VALUE ((1ull << 62) - 1)

// Record slots are handed out in draw order (stable compaction): slot order == draw order, which is
// what lets k_tile break depth ties by comparing slots.  Inside a CTA: ballot/popc-style prefix sums
// over the per-thread output counts (0..4).  Across CTAs: single-pass chained scan with decoupled
// look-back; CTAs take a ticket so that the chain follows launch order and cannot deadlock.
__global__ void __launch_bounds__(SETUP_THREADS) k_setup(const FrameUniforms *__restrict__ Up, const SceneDev S,
                                                         const FrameDev W) {
    const FrameUniforms &U = *Up; // per-frame uniforms, device-resident (one upload per frame; the launches never change)
    __shared__ uint32_t warp_tot[SETUP_THREADS / 32];
    __shared__ uint32_t s_ticket, s_base;
    const CtaTrace trace_(W, 1u);
    pdl_prologue(U.pdl_early != 0);

    if (threadIdx.x == 0) s_ticket = atomicAdd(&W.counters[3], 1u);
    __syncthreads();
    const uint32_t bid = s_ticket;
    const uint32_t tri = bid * SETUP_THREADS + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    uint32_t n_out = 0;       // opaque record slots this thread reserves (0, 1, or 4 for a triangle to clip)
    bool clipped = false;     // the triangle straddles the near or far plane: k_clip fills its four slots
    bool transparent = false; // unclipped transparent triangle: writes its own ordered slots
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
            // Rare and register-hungry: handed to k_clip through a queue.  An opaque triangle reserves the
            // maximum of four consecutive slots here, so that slot order stays draw order.
            n_out = transparent ? 0u : 4u;
            clipped = true;
            alive = false;
            transparent = false; // a clipped transparent triangle wrote its four ordered slots already
        } else if (alive) {
            const float sx[3] = {W.v_sx[vi[0]], W.v_sx[vi[1]], W.v_sx[vi[2]]};
            const float sy[3] = {W.v_sy[vi[0]], W.v_sy[vi[1]], W.v_sy[vi[2]]};
            const float dep[3] = {W.v_depth[vi[0]], W.v_depth[vi[1]], W.v_depth[vi[2]]};
            alive = setup_raster(U, sx, sy, dep, tri * 4u, r);
            if (!alive) r.id = NO_SLOT;
        }
        if (!clipped) n_out = (alive && !transparent && touches_stripe(U, r)) ? 1u : 0u;
    }

    // ---- CTA-wide exclusive prefix of n_out -------------------------------------------------
    uint32_t incl = n_out;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (uint32_t)d) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    // ---- chained scan across CTAs (warp 0) ----------------------------------------------------
    if (warp == 0) {
        uint32_t wt = lane < SETUP_THREADS / 32 ? warp_tot[lane] : 0u;
        uint32_t wi = wt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= (uint32_t)d) wi += up;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, wi, 31);
        if (lane < SETUP_THREADS / 32) warp_tot[lane] = wi - wt; // exclusive offset of each warp
        volatile unsigned long long *desc = reinterpret_cast<volatile unsigned long long *>(W.scan_desc);
        uint32_t base = 0;
        if (bid == 0) {
            if (lane == 0) desc[0] = DESC_PREFIX | total;
        } else {
            if (lane == 0) desc[bid] = DESC_AGGREGATE | total;
            // look back 32 predecessors at a time until one of them has its inclusive prefix
            int look = (int)bid - 1;
            while (true) {
                const int j = look - (int)lane;
                unsigned long long d = j >= 0 ? desc[j] : DESC_PREFIX;
                // wait until every descriptor in the window up to the first PREFIX is published
                const uint32_t is_prefix = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 2);
                const uint32_t not_ready = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 0);
                const int first_prefix = is_prefix ? __ffs(is_prefix) - 1 : 32;
                const uint32_t window = first_prefix >= 31 ? 0xFFFFFFFFu : ((2u << first_prefix) - 1u);
                if (not_ready & window) continue; // spin
                uint32_t v = (lane <= (uint32_t)first_prefix && j >= 0) ? (uint32_t)(d & DESC_VALUE) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                base += v;
                if (first_prefix < 32) break;
                look -= 32;
            }
            if (lane == 0) desc[bid] = DESC_PREFIX | (unsigned long long)(base + total);
        }
        if (lane == 0) {
            s_base = base;
            if (bid == gridDim.x - 1) W.counters[0] = base + total; // total record count of the frame
        }
    }
    __syncthreads();

    const uint32_t slot0 = s_base + warp_tot[warp] + incl - n_out;
    if (n_out && slot0 + n_out > W.rec_cap) {
        atomicOr(&W.counters[2], OVERFLOW_RECORDS); // the host re-renders the frame with larger buffers
        return;
    }
    if (clipped) {
        {
            RasterRec empty;
            empty.id = NO_SLOT;
            empty.bbx = empty.bby = 0;
            empty.ax = empty.ay = empty.bx = empty.by = empty.cx = empty.cy = empty.da = empty.db = empty.dc = 0.0f;
            for (uint32_t k = 0; k < n_out; k++) store_raster(W.rrec + slot0 + k, empty); // k_clip overwrites the used ones
            const uint32_t q = atomicAdd(&W.counters[5], 1u);
            W.clip_queue[q] = make_uint2(tri, n_out ? slot0 : NO_SLOT);
        }
        return;
    }

    RasterRec *rdst = nullptr;
    ShadeRec *sdst = nullptr;
    if (n_out) {
        rdst = W.rrec + slot0;
        sdst = W.srec + slot0;
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
// k_clip : the triangles k_setup queued because they straddle the near or far plane
// ------------------------------------------------------------------------------------------
constexpr int CLIP_THREADS = 128;

__global__ void __launch_bounds__(CLIP_THREADS) k_clip(const FrameUniforms *__restrict__ Up, const SceneDev S,
                                                       const FrameDev W) {
    const FrameUniforms &U = *Up; // per-frame uniforms, device-resident (one upload per frame; the launches never change)
    const CtaTrace trace_(W, 2u);
    pdl_prologue(U.pdl_early != 0);
    const uint32_t n = W.counters[5];
    for (uint32_t q = blockIdx.x * CLIP_THREADS + threadIdx.x; q < n; q += gridDim.x * CLIP_THREADS) {
        const uint2 item = W.clip_queue[q];
        const uint32_t tri = item.x, slot0 = item.y;
        const uint32_t vi[3] = {S.idx[0][tri], S.idx[1][tri], S.idx[2][tri]};
        const uint32_t mat = S.tri_mat[tri];
        const bool transparent = (mat >> 31) != 0;
        RasterRec out_r[4];
        ShadeRec out_s[4];
        const int n_keep = clip_triangle(U, S, W, tri, vi, mat & 0x7FFFFFFFu, transparent, transparent ? S.tri_tslot[tri] : 0u,
                                         out_r, out_s);
        if (slot0 == NO_SLOT) continue; // transparent: clip_triangle wrote its ordered slots
        for (int k = 0; k < n_keep; k++) {
            store_raster(W.rrec + slot0 + k, out_r[k]);
            store_shade(W.srec + slot0 + k, out_s[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
thread_local unsigned g_clip_ctas = 148u * 2u; // scene.cpp: DRAW_B200_CLIP_CTAS
void launch_clip(const FrameUniforms &U, const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream) {
    if (!S.n_triangles) return;
    launch_pdl(k_clip, g_clip_ctas, CLIP_THREADS, stream, dU, S, W);
}

void launch_vertex(const FrameUniforms &U, const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream) {
    const uint32_t by_vertex = (S.n_vertices + 255) / 256, by_list = (U.n_lists + 255) / 256;
    uint32_t blocks = by_vertex > 1 ? by_vertex : 1;
    if (blocks < by_list && blocks < 148 * 4) blocks = by_list < 148 * 4 ? by_list : 148 * 4;
    launch_pdl(k_vertex, blocks, 256, stream, dU, S, W);
}

void launch_setup(const FrameUniforms &U, const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream) {
    if (!S.n_triangles) return;
    launch_pdl(k_setup, (S.n_triangles + SETUP_THREADS - 1) / SETUP_THREADS, SETUP_THREADS, stream, dU, S, W);
}

} // namespace drawb200
