// k_front.cu — everything of a frame that comes before rasterisation, as ONE persistent cooperative kernel
// (mororo18/draw scene/mod.rs:901-1085 up to the call of draw_triangle_with_attributes, plus the binning the
// reference does not have).
//
// The stages are latency-bound and tiny next to a launch (a 20 k-triangle scene is one triangle per thread), so
// they are phases of one grid separated by grid-wide barriers instead of seven launches:
//
//   P0  per vertex    light / halfway / depth (scene/mod.rs:917-926), screen xy of the unclipped vertex
//                     (:1047-1058), view-plane side flags (:634-660); four vertices per thread from the SoA position
//                     streams (128-bit loads), packed 16 / 32-byte records out.  Resets the frame's counters.
//   ---- barrier ----
//   P1  per triangle  coalesced index loads, packed gathers, back-face cull (:1016-1027), lateral reject + near/far
//                     clip (:43-90, :662-746), snap + bbox + zero-area cull (canvas.rs:585-666), draw-order-
//                     preserving slot numbers (block index * SLOT_STRIDE + prefix sum inside the block) with dense
//                     storage reserved by one atomic per block, record + prepared record + shading record
//                     write, and BINNING of the block's records by the CTA that made them, the (record, tile) pairs of
//                     the block spread evenly over its threads: exact can-it-cover test per tile (rect_may_cover),
//                     class by the bbox area inside the tile; medium / small references are appended to the frame-
//                     wide lists k_raster walks, large and transparent ones to (tile, slot) pair lists with a
//                     per-tile count.  A count pass, ONE round of atomics per CTA and class, a write pass.
//   ---- barrier ----
//   P2  per tile      list offsets (per-CTA scan + one atomic per class), cost, number of windows, cost bucket and
//                     rank inside the bucket (one atomic per warp and bucket); the tile's work items go straight to
//                     bucket segment + rank of k_tile's work list, the empty tiles to the list for the clear.
// (The pairs are scattered into the tiles' lists by the next kernel's prologue: k_raster, which does not need them.)
//
// The grid is 148 x c CTAs of 128 threads at <= 64 registers, c = 1 for small scenes — it fits on an SM beside
// three CTAs of another frame's k_tile, so frames in flight still overlap — and launched cooperatively
// (all CTAs co-resident: the barriers cannot deadlock).  Arithmetic contract: device_math.cuh.
#include "device_math.cuh"

namespace drawb200 {

constexpr int FRONT_THREADS = 128; // a block of 128 triangles per CTA step: small scenes still spread over all SMs
constexpr uint32_t FULL = 0xFFFFFFFFu;

// Estimated k_tile work in quarter block-iterations (one iteration = a warp testing an 8x4 block, ~130 instructions):
// a large reference costs every warp of the CTA a pass, a transparent one a pass of the ordered blend.
constexpr uint32_t COST_LARGE = 80u, COST_TRANSPARENT = 40u, COST_TILE_BASE = 16u;

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Grid-wide barrier.  The counter only ever grows (the host tracks its value, FrameUniforms::bar_base), so it
// needs no reset between frames; `target` is the value it has once every CTA has arrived.  The launch is
// cooperative, so all CTAs are resident and the wait ends; the time-out only turns a broken launch into an
// error the host reports instead of a hung GPU.
__device__ __forceinline__ void grid_barrier(uint32_t *counters, uint32_t target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t *bar = counters + CNT_BARRIER;
        __threadfence();
        atomicAdd(bar, 1u);
        const unsigned long long t0 = global_timer_ns();
        while ((int32_t)(ld_acquire_gpu(bar) - target) < 0) {
            if (global_timer_ns() - t0 > 4000000000ull) {
                atomicOr(counters + CNT_OVERFLOW, OVERFLOW_STALL);
                break;
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// P0 : per vertex
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void vertex_visual(const FrameUniforms &U, const v3 p, float4 &a, float4 &l, float4 &h) {
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
    // scene/mod.rs:1047-1058 for an unclipped corner: rows x, y, w of matrix_transf, then x/w, y/w
    const float cx = mat_row(&U.m[0], p);
    const float cy = mat_row(&U.m[4], p);
    const float cw = mat_row(&U.m[12], p);
    uint32_t flags = 0;
#pragma unroll
    for (int pl = 0; pl < 6; pl++) {
        const float f = plane_eval(U.planes[pl], p);
        flags |= (f > 0.0f ? 1u : 0u) << (2 * pl);
        flags |= (f <= 0.0f ? 1u : 0u) << (2 * pl + 1);
    }
    a = make_float4(FDIV(cx, cw), FDIV(cy, cw), eye_len, __uint_as_float(flags));
    l = make_float4(light.x, light.y, light.z, halfway.x);
    h = make_float4(halfway.y, halfway.z, 0.0f, 0.0f);
}

__device__ __forceinline__ void phase_vertex(const FrameUniforms &U, const SceneDev &S, const FrameDev &W) {
    const uint32_t gtid = blockIdx.x * FRONT_THREADS + threadIdx.x, gsize = gridDim.x * FRONT_THREADS;
    // per-frame reset of the binning state (read and written after the first barrier)
    for (uint32_t t = gtid; t < U.n_coarse; t += gsize) {
        W.l_count[t] = 0;
        W.t_count[t] = 0;
        W.ms_weight[t] = 0;
    }
    if (gtid < (uint32_t)CNT_BARRIER) W.counters[gtid] = 0; // everything but the barrier word

    const uint32_t n = S.n_vertices, n4 = n & ~3u;
    if (n <= gsize) {
        // fewer vertices than threads: one each, so that every SM takes part (the phase is a chain of square roots and
        // divisions: its duration is per thread, not per byte)
        if (gtid < n) {
            float4 a, l, h;
            vertex_visual(U, v3{__ldg(S.px + gtid), __ldg(S.py + gtid), __ldg(S.pz + gtid)}, a, l, h);
            W.vA[gtid] = a;
            W.vLH[2 * gtid] = l;
            W.vLH[2 * gtid + 1] = h;
        }
        return;
    }
    for (uint32_t base = gtid * 4u; base < n4; base += gsize * 4u) {
        // four consecutive vertices: one 128-bit load per SoA stream, 64 + 128 contiguous bytes out
        const float4 x = __ldg(reinterpret_cast<const float4 *>(S.px + base));
        const float4 y = __ldg(reinterpret_cast<const float4 *>(S.py + base));
        const float4 z = __ldg(reinterpret_cast<const float4 *>(S.pz + base));
        const float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w}, zs[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float4 a, l, h;
            vertex_visual(U, v3{xs[j], ys[j], zs[j]}, a, l, h);
            W.vA[base + j] = a;
            W.vLH[2 * (base + j)] = l;
            W.vLH[2 * (base + j) + 1] = h;
        }
    }
    if (gtid < n - n4) { // the last n % 4 vertices
        const uint32_t i = n4 + gtid;
        float4 a, l, h;
        vertex_visual(U, v3{S.px[i], S.py[i], S.pz[i]}, a, l, h);
        W.vA[i] = a;
        W.vLH[2 * i] = l;
        W.vLH[2 * i + 1] = h;
    }
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

// Tiles of this launch's rows (sort-first partition, FrameUniforms) that a record's bbox touches: columns
// tx0..tx1, rows ty_first, ty_first + row_step, ... (n_rows of them).
struct TileRange {
    int x0, x1, y0, y1; // bbox in pixels
    int tx0, cols, ty_first, n_rows, row_step;
    __device__ __forceinline__ int count() const { return cols * n_rows; }
};
__device__ __forceinline__ TileRange tile_range(const FrameUniforms &U, uint32_t bbx, uint32_t bby) {
    TileRange t;
    t.x0 = (int)(bbx & 0xFFFF); t.x1 = (int)(bbx >> 16);
    t.y0 = (int)(bby & 0xFFFF); t.y1 = (int)(bby >> 16);
    t.tx0 = t.x0 / TILE_W;
    t.cols = t.x1 / TILE_W - t.tx0 + 1;
    const int step = (int)U.row_step;
    const int lo = max(t.y0 / TILE_H, (int)U.tile_y_begin), hi = min(t.y1 / TILE_H, (int)U.tile_y_end - 1);
    t.row_step = step;
    if (step == 1) { // every row (no software division on the common path)
        t.ty_first = lo;
        t.n_rows = lo <= hi ? hi - lo + 1 : 0;
        return t;
    }
    const int first = lo + ((int)U.row_phase + step - lo % step) % step; // first row >= lo of this launch
    t.ty_first = first;
    t.n_rows = first <= hi ? (hi - first) / step + 1 : 0;
    return t;
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
    const float4 p = __ldg(S.pos4 + v), nn = __ldg(S.nrm4 + n), a = W.vA[v], l = W.vLH[2 * v], h = W.vLH[2 * v + 1];
    const float2 uv = __ldg(S.uv2 + t);
    o.p[0] = p.x; o.p[1] = p.y; o.p[2] = p.z;
    o.a[0] = a.z;
    o.a[1] = nn.x; o.a[2] = nn.y; o.a[3] = nn.z;
    o.a[4] = l.x; o.a[5] = l.y; o.a[6] = l.z;
    o.a[7] = l.w; o.a[8] = h.x; o.a[9] = h.y;
    o.a[10] = uv.x; o.a[11] = uv.y;
}

// The full clip path for one triangle that straddles the near or far plane: up to 4 outputs, each
// projected (:1047-1063), set up and stored — raster, prepared and shading record — at slot_base + emission
// index (opaque: the four slots k_front reserved in draw order; transparent: the ordered slots 4*ordinal + k).
// Returns the bit mask of the emission indices that survive set-up; their bboxes are returned in bbx / bby.
__device__ __noinline__ uint32_t clip_triangle(const FrameUniforms &U, const SceneDev &S, const FrameDev &W, uint32_t tri,
                                               const uint32_t vi[3], uint32_t material, bool transparent, uint32_t loc_base,
                                               uint32_t bbx[4], uint32_t bby[4]) {
    ClipTri in;
#pragma unroll
    for (int c = 0; c < 3; c++) gather_clip_vert(S, W, vi[c], S.idx[3 + c][tri], S.idx[6 + c][tri], in.v[c]);

    ClipTri near_out[2], out[4];
    const int n_near = clip_plane(U.planes[0], in, near_out);
    int n_out = 0;
    for (int i = 0; i < n_near; i++) n_out += clip_plane(U.planes[1], near_out[i], out + n_out);

    uint32_t kept = 0;
    for (int k = 0; k < n_out; k++) {
        RasterRec r;
        float sx[3], sy[3], dep[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const v3 p = cv_pos(out[k].v[c]);
            const float w = mat_row(&U.m[12], p);
            sx[c] = FDIV(mat_row(&U.m[0], p), w);
            sy[c] = FDIV(mat_row(&U.m[4], p), w);
            dep[c] = out[k].v[c].a[0];
        }
        if (!setup_raster(U, sx, sy, dep, tri * 4u + (uint32_t)k, r)) continue;
        const uint32_t slot = loc_base + (uint32_t)k;   // where the record is stored (its slot number is the caller's slot + k)
        if (!transparent && slot >= W.rec_cap) continue; // overflow (flagged by the caller): the host re-renders
        ShadeRec s;
        shade_from_clip(out[k], material, s);
        PrepRec p;
        make_prep(r, p);
        store_raster((transparent ? W.t_rrec : W.rrec) + slot, r);
        store_prep((transparent ? W.t_prep : W.prep) + slot, p);
        store_shade((transparent ? W.t_srec : W.srec) + slot, s);
        kept |= 1u << k;
        bbx[k] = r.bbx;
        bby[k] = r.bby;
    }
    return kept;
}

// ------------------------------------------------------------------------------------------
// binning (part of P1)
// ------------------------------------------------------------------------------------------
// The part of a prepared record binning needs.
struct BinTri {
    float ecx[3], ecy[3], ek1[3], ek2[3];
    uint32_t flags;
};
__device__ __forceinline__ BinTri bin_tri_of(const PrepRec &p) {
    BinTri b;
#pragma unroll
    for (int i = 0; i < 3; i++) { b.ecx[i] = p.ecx[i]; b.ecy[i] = p.ecy[i]; b.ek1[i] = p.ek1[i]; b.ek2[i] = p.ek2[i]; }
    b.flags = p.flags;
    return b;
}
__device__ __forceinline__ BinTri bin_tri_load(const PrepRec *src) {
    const uint4 *q = reinterpret_cast<const uint4 *>(src);
    const uint4 q0 = q[0], q1 = q[1], q2 = q[2], q6 = q[6];
    BinTri b;
    b.ecx[0] = __uint_as_float(q0.x); b.ecx[1] = __uint_as_float(q0.y); b.ecx[2] = __uint_as_float(q0.z);
    b.ecy[0] = __uint_as_float(q0.w); b.ecy[1] = __uint_as_float(q1.x); b.ecy[2] = __uint_as_float(q1.y);
    b.ek1[0] = __uint_as_float(q1.z); b.ek1[1] = __uint_as_float(q1.w); b.ek1[2] = __uint_as_float(q2.x);
    b.ek2[0] = __uint_as_float(q2.y); b.ek2[1] = __uint_as_float(q2.z); b.ek2[2] = __uint_as_float(q2.w);
    b.flags = q6.y;
    return b;
}
__device__ __forceinline__ BinTri bin_tri_shfl(const BinTri &t, int src) {
    BinTri w;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        w.ecx[i] = __shfl_sync(FULL, t.ecx[i], src); w.ecy[i] = __shfl_sync(FULL, t.ecy[i], src);
        w.ek1[i] = __shfl_sync(FULL, t.ek1[i], src); w.ek2[i] = __shfl_sync(FULL, t.ek2[i], src);
    }
    w.flags = __shfl_sync(FULL, t.flags, src);
    return w;
}

// One (record, tile) pair: exact can-it-cover test (rect_may_cover), class by the bbox area inside the tile.
//   cls 0 large, 1 medium, 2 small, 3 transparent, 4 none; entries = list entries the pair takes (a medium reference
//   with many 8x4 blocks is entered 2-4 times, each entry naming a share of the blocks, so that k_raster's
//   warp-per-reference jobs stay short)
struct PairClass {
    uint32_t cls, blocks, entries, tile;
};
__device__ __forceinline__ PairClass classify_pair(const FrameUniforms &U, const BinTri &t, int x0, int x1, int y0, int y1, int tx, int ty,
                                                   bool transparent) {
    PairClass c;
    c.cls = 4u;
    c.blocks = c.entries = 0u;
    c.tile = (uint32_t)ty * U.tiles_x + (uint32_t)tx;
    const int lx = max(x0, tx * TILE_W), hx = min(x1, tx * TILE_W + TILE_W - 1);
    const int ly = max(y0, ty * TILE_H), hy = min(y1, ty * TILE_H + TILE_H - 1);
    if (rect_may_cover(t, (float)lx, (float)hx, (float)ly, (float)hy)) {
        const int area = (hx - lx + 1) * (hy - ly + 1);
        c.cls = transparent ? 3u : (area <= SMALL_AREA ? 2u : (area <= MEDIUM_AREA ? 1u : 0u));
        c.blocks = (uint32_t)((hx - lx) / 8 + 1) * (uint32_t)((hy - ly) / 4 + 1);
        c.entries = c.cls == 1u ? min(4u, (c.blocks + 7u) / 8u) : 1u;
    }
    return c;
}

// The lane's reserved positions in the four frame-wide lists (or, in the counting pass, its entry counts).
struct ListPos {
    uint32_t l, m, s, t;
    __device__ __forceinline__ void count(const PairClass &c) {
        l += c.cls == 0u ? c.entries : 0u;
        m += c.cls == 1u ? c.entries : 0u;
        s += c.cls == 2u ? c.entries : 0u;
        t += c.cls == 3u ? c.entries : 0u;
    }
};

// Writes the list entries of one classified pair at the thread's reserved positions (advanced), and adds the
// pair to its tile's counts.  Entries beyond the capacity are dropped and flagged (the host grows the buffers and
// re-renders the frame).
__device__ __forceinline__ void emit_pair(const FrameDev &W, const PairClass &c, uint32_t slot, int tx, int ty, ListPos &pos) {
    if (c.cls >= 4u) return;
    const uint32_t at = c.cls == 0u ? pos.l : (c.cls == 1u ? pos.m : (c.cls == 2u ? pos.s : pos.t));
    pos.count(c);
    if (at + c.entries > W.refs_cap) {
        atomicOr(W.counters + CNT_OVERFLOW, OVERFLOW_REFS);
        return;
    }
    const uint32_t tile_xy = (uint32_t)tx | (uint32_t)ty << 10;
    if (c.cls == 0u) {
        W.l_pairs[at] = make_uint2(c.tile, slot);
        atomicAdd(&W.l_count[c.tile], 1u);
    } else if (c.cls == 3u) {
        W.t_pairs[at] = make_uint2(c.tile, slot);
        atomicAdd(&W.t_count[c.tile], 1u);
    } else if (c.cls == 2u) {
        W.s_refs[at] = make_uint2(slot, tile_xy);
        atomicAdd(&W.ms_weight[c.tile], 1u);
    } else {
        for (uint32_t part = 0; part < c.entries; part++) W.m_refs[at + part] = make_uint2(slot, tile_xy | part << 21 | (c.entries - 1u) << 23);
        atomicAdd(&W.ms_weight[c.tile], c.blocks);
    }
}

// ------------------------------------------------------------------------------------------
// P1 : per triangle
// ------------------------------------------------------------------------------------------

// A record to bin: bbox, slot (bit 31: transparent), and where its edge functions are: in sh.tri[tri_idx] (the
// record of an unclipped triangle, staged by the thread that made it) or, for the outputs of the clip path
// (tri_idx = NO_TRI), in its prepared record in global memory.  A block of FRONT_THREADS triangles makes at most
// four records each.
struct BinJob {
    uint32_t bbx, bby, slot, tri_idx;
};
// A record whose bbox covers more tiles than this is not binned by the CTA that made it (a block of triangles that fill
// the screen would keep one CTA busy for hundreds of microseconds): it goes to a frame-wide queue (FrameDev::huge_jobs)
// whose (record, tile) pairs are split evenly over the whole grid in a phase of its own.
constexpr uint32_t HUGE_TILES = 32;
constexpr int MAX_JOBS = 4 * FRONT_THREADS;
static_assert(SLOT_STRIDE == 4u * FRONT_THREADS, "a block of FRONT_THREADS triangles numbers at most SLOT_STRIDE records");
constexpr uint32_t NO_TRI = 0xFFFFFFFFu, DIVERTED = 0xFFFFFFFEu; // BinJob::tri_idx: edges in global memory / job moved to the huge queue
constexpr int TRI_WORDS = 13; // BinTri as words: odd stride, conflict-free for neighbouring jobs
struct FrontShared {
    uint32_t warp_tot[FRONT_THREADS / 32];
    uint32_t ticket, base;
    uint32_t sum4[4][FRONT_THREADS / 32];
    uint32_t base4[4];
    BinJob jobs[MAX_JOBS];
    uint32_t pair_prefix[MAX_JOBS + 1]; // exclusive prefix of the jobs' tile counts
    float tri[FRONT_THREADS][TRI_WORDS];
};

// Exclusive prefix sums of four values per thread over the CTA (scratch: sh.sum4); tot[k] = the sums.
__device__ __forceinline__ void block_exclusive4(uint32_t v[4], FrontShared &sh, uint32_t tot[4]) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t up = __shfl_up_sync(FULL, incl[k], d);
            if (lane >= (uint32_t)d) incl[k] += up;
        }
    }
    __syncthreads(); // scratch of an earlier call is no longer read
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < 4; k++) sh.sum4[k][warp] = incl[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t before = 0, sum = 0;
#pragma unroll
        for (int w = 0; w < FRONT_THREADS / 32; w++) {
            const uint32_t t = sh.sum4[k][w];
            before += w < (int)warp ? t : 0u;
            sum += t;
        }
        tot[k] = sum;
        v[k] = before + incl[k] - v[k];
    }
}
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, FrontShared &sh, uint32_t *total) {
    uint32_t a[4] = {v, 0u, 0u, 0u}, tot[4];
    block_exclusive4(a, sh, tot);
    *total = tot[0];
    return a[0];
}

// Bins the block's records (sh.jobs[0 .. n_jobs)).  The unit of work is a (record, tile) PAIR: the pairs of all the
// block's records form one flat index space (prefix sums of the records' tile counts) that the CTA's threads
// stride over, so a record covering a thousand tiles is spread over the whole CTA and one covering a single tile
// costs a single step.  Two passes: the first counts the list entries of every class per thread, then the CTA
// reserves its ranges in the four frame-wide lists with ONE round of atomics (four threads, four counters, in
// flight together — a round trip per pair would be the whole duration of the phase), the second writes them.
// A thread's first pair (usually its only one) is classified once and kept in registers between the passes.
struct PairEval {
    PairClass c;
    uint32_t slot;
    int tx, ty;
};
__device__ __forceinline__ PairEval eval_pair(const FrameUniforms &U, const FrameDev &W, const FrontShared &sh, uint32_t p) {
    uint32_t lo = 0, hi = MAX_JOBS; // largest j with pair_prefix[j] <= p (jobs without tiles are skipped over)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sh.pair_prefix[mid] <= p) lo = mid; else hi = mid;
    }
    const BinJob job = sh.jobs[lo];
    const bool transparent = (job.slot >> 31) != 0;
    PairEval e;
    e.slot = job.slot & 0x7FFFFFFFu;
    const TileRange tr = tile_range(U, job.bbx, job.bby);
    const int i = (int)(p - sh.pair_prefix[lo]);
    e.tx = tr.tx0 + i % tr.cols;
    e.ty = tr.ty_first + (i / tr.cols) * tr.row_step;
    BinTri t;
    if (job.tri_idx != NO_TRI) {
        const float *w = sh.tri[job.tri_idx];
#pragma unroll
        for (int k = 0; k < 3; k++) { t.ecx[k] = w[k]; t.ecy[k] = w[3 + k]; t.ek1[k] = w[6 + k]; t.ek2[k] = w[9 + k]; }
        t.flags = __float_as_uint(w[12]);
    } else {
        // written by this CTA before the barrier; an opaque record of this block is stored at sh.base + its number in the block
        t = bin_tri_load(transparent ? W.t_prep + e.slot : W.prep + (sh.base + (e.slot & (SLOT_STRIDE - 1u))));
    }
    e.c = classify_pair(U, t, tr.x0, tr.x1, tr.y0, tr.y1, e.tx, e.ty, transparent);
    return e;
}

// Pass over the pairs [p_lo, p_hi) of the job window in shared memory (sh.jobs, sh.pair_prefix), CTA-strided.
// EMIT false: counts the list entries into pos; true: writes them at pos.  `first` caches the evaluation of the thread's
// first pair between the two passes (usually its only one).
template <bool EMIT>
__device__ __forceinline__ void bin_pass(const FrameUniforms &U, const FrameDev &W, const FrontShared &sh, uint32_t p_lo, uint32_t p_hi,
                                         ListPos &pos, PairEval &first, bool use_first) {
    const uint32_t tid = threadIdx.x;
    uint32_t p = p_lo + tid;
    if (use_first && p < p_hi) {
        if (!EMIT) {
            first = eval_pair(U, W, sh, p);
            pos.count(first.c);
        } else {
            emit_pair(W, first.c, first.slot, first.tx, first.ty, pos);
        }
        p += FRONT_THREADS;
    }
#pragma unroll 1
    for (; p < p_hi; p += FRONT_THREADS) {
        const PairEval e = eval_pair(U, W, sh, p);
        if (!EMIT) pos.count(e.c);
        else emit_pair(W, e.c, e.slot, e.tx, e.ty, pos);
    }
}

// One reservation per CTA and class: per-thread entry counts in, per-thread first positions out.
__device__ __forceinline__ void reserve_lists(const FrameDev &W, FrontShared &sh, ListPos &pos) {
    const uint32_t tid = threadIdx.x;
    uint32_t v[4] = {pos.l, pos.m, pos.s, pos.t}, tot[4];
    block_exclusive4(v, sh, tot);
    if (tid < 4) {
        const int ctr = tid == 0 ? CNT_L_PAIRS : (tid == 1 ? CNT_MEDIUM : (tid == 2 ? CNT_SMALL : CNT_T_PAIRS));
        sh.base4[tid] = tot[tid] ? atomicAdd(W.counters + ctr, tot[tid]) : 0u;
    }
    __syncthreads();
    pos = ListPos{sh.base4[0] + v[0], sh.base4[1] + v[1], sh.base4[2] + v[2], sh.base4[3] + v[3]};
}

// Bins the block's records (sh.jobs[0 .. n_jobs), tile counts in n_tiles per thread's four jobs).
__device__ __forceinline__ void bin_block(const FrameUniforms &U, const FrameDev &W, FrontShared &sh, const uint32_t n_tiles[4]) {
    const uint32_t tid = threadIdx.x;
    uint32_t n_pairs;
    uint32_t at = block_exclusive(n_tiles[0] + n_tiles[1] + n_tiles[2] + n_tiles[3], sh, &n_pairs);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        sh.pair_prefix[tid * 4u + (uint32_t)k] = at;
        at += n_tiles[k];
    }
    if (tid == FRONT_THREADS - 1) sh.pair_prefix[MAX_JOBS] = at;
    __syncthreads();
    if (n_pairs == 0) return; // block-uniform
    ListPos pos{0u, 0u, 0u, 0u};
    PairEval first;
    first.c.cls = 4u;
    first.c.blocks = first.c.entries = first.c.tile = 0u;
    first.slot = 0u;
    first.tx = first.ty = 0;
    bin_pass<false>(U, W, sh, 0u, n_pairs, pos, first, true);
    reserve_lists(W, sh, pos);
    bin_pass<true>(U, W, sh, 0u, n_pairs, pos, first, true);
}

// Does edge e of the triangle admit any pixel of tile (tx, row band [ly, hy])?  The same arithmetic as rect_may_cover
// for the tile's rectangle clipped to the record's bbox.
__device__ __forceinline__ bool edge_admits_tile(const BinTri &t, int e, int tx, int x0, int x1, float lyf, float hyf) {
    const int lx = max(x0, tx * TILE_W), hx = min(x1, tx * TILE_W + TILE_W - 1);
    const float xm = t.ecx[e] >= 0.0f ? (float)hx : (float)lx, ym = t.ecy[e] >= 0.0f ? hyf : lyf;
    const float em = FSUB(FADD(FADD(FMUL(t.ecx[e], xm), FMUL(t.ecy[e], ym)), t.ek1[e]), t.ek2[e]);
    return em > 0.0f || (em == 0.0f && (t.flags & (1u << e)));
}

// The tiles of one tile row that a record can cover: an interval [a, b] of tile columns (empty: a > b).  Exactly the
// tiles rect_may_cover admits: per edge, the best-corner value is monotone in the tile column (the corner's x grows
// with the column and every rounding step is monotone), so the admitted columns of an edge are a half-line whose end
// a binary search finds with the very evaluation rect_may_cover would make; the three half-lines intersect in an
// interval.  A triangle that fills the screen costs 3 x log2(columns) evaluations per row instead of one per tile.
__device__ __forceinline__ void row_interval(const BinTri &t, const TileRange &tr, int ty, int &a, int &b) {
    a = tr.tx0;
    b = tr.tx0 + tr.cols - 1;
    if (t.flags & TRI_SLOW) return; // rect_may_cover admits everything
    const float lyf = (float)max(tr.y0, ty * TILE_H), hyf = (float)min(tr.y1, ty * TILE_H + TILE_H - 1);
    // the whole row band first: most rows of a big bbox hold nothing of the triangle (one evaluation instead of three searches)
    if (!rect_may_cover(t, (float)tr.x0, (float)tr.x1, lyf, hyf)) {
        b = a - 1;
        return;
    }
#pragma unroll
    for (int e = 0; e < 3; e++) { // unrolled: the edge arrays stay in registers
        if (a > b) break;
        if (t.ecx[e] >= 0.0f) { // admitted columns: [first admitted, b]
            int lo = a, hi = b + 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (edge_admits_tile(t, e, mid, tr.x0, tr.x1, lyf, hyf)) hi = mid; else lo = mid + 1;
            }
            a = lo;
        } else { // admitted columns: [a, last admitted]
            int lo = a - 1, hi = b;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (edge_admits_tile(t, e, mid, tr.x0, tr.x1, lyf, hyf)) lo = mid; else hi = mid - 1;
            }
            b = lo;
        }
    }
}

// Class of a (record, tile) pair known to be admitted (classify_pair without the test).
__device__ __forceinline__ PairClass class_of_admitted(const FrameUniforms &U, int x0, int x1, int y0, int y1, int tx, int ty, bool transparent) {
    PairClass c;
    c.tile = (uint32_t)ty * U.tiles_x + (uint32_t)tx;
    const int lx = max(x0, tx * TILE_W), hx = min(x1, tx * TILE_W + TILE_W - 1);
    const int ly = max(y0, ty * TILE_H), hy = min(y1, ty * TILE_H + TILE_H - 1);
    const int area = (hx - lx + 1) * (hy - ly + 1);
    c.cls = transparent ? 3u : (area <= SMALL_AREA ? 2u : (area <= MEDIUM_AREA ? 1u : 0u));
    c.blocks = (uint32_t)((hx - lx) / 8 + 1) * (uint32_t)((hy - ly) / 4 + 1);
    c.entries = c.cls == 1u ? min(4u, (c.blocks + 7u) / 8u) : 1u;
    return c;
}

// Writes the list entries of one admitted pair at position `at` of its class' list (emit_pair with the position given).
__device__ __forceinline__ void emit_at(const FrameDev &W, const PairClass &c, uint32_t slot, int tx, int ty, uint32_t at) {
    ListPos pos{at, at, at, at};
    emit_pair(W, c, slot, tx, ty, pos);
}

// The frame's huge records (FrameDev::huge_jobs, appended by the triangle phase).  The records are dealt out to the
// warps of the whole grid; a warp holds its record's edge functions in registers.  Per chunk of up to 128 tile rows:
// a lane takes a row (four row groups) and finds the interval of admitted columns (row_interval) — work in
// proportion to the references made, not to the tiles of the bbox (a near-plane-clipped triangle whose bbox is the
// whole screen typically covers a few per cent of it); the admitted (row, column) pairs of a group then form one
// flat space that the 32 lanes stride over (a row that spans the screen is not one lane's job), once to count the
// list entries per class, then — after ONE round of atomics per chunk — to write them at ballot / prefix positions.
__device__ __forceinline__ void phase_huge(const FrameUniforms &U, const FrameDev &W, uint32_t n_jobs) {
    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    const uint32_t gwarp = blockIdx.x * (FRONT_THREADS / 32) + (threadIdx.x >> 5), n_warps = gridDim.x * (FRONT_THREADS / 32);
    constexpr int GROUPS = 4;
    if (gwarp >= n_jobs) return;
    uint4 q = __ldcg(&W.huge_jobs[gwarp]); // the same address in every lane: one broadcast load
#pragma unroll 1
    for (uint32_t j = gwarp; j < n_jobs; j += n_warps) { // warp-uniform
        const bool transparent = (q.z >> 31) != 0;
        const uint32_t slot = q.z & 0x7FFFFFFFu;
        const TileRange tr = tile_range(U, q.x, q.y);
        const BinTri t = bin_tri_load(transparent ? W.t_prep + slot : W.prep + record_index_cg(W, slot));
        if (j + n_warps < n_jobs) q = __ldcg(&W.huge_jobs[j + n_warps]); // the next record's header
#pragma unroll 1
        for (int chunk = 0; chunk < tr.n_rows; chunk += 32 * GROUPS) {
            int a[GROUPS];
            uint32_t len[GROUPS], excl[GROUPS], tot[GROUPS];
#pragma unroll
            for (int g = 0; g < GROUPS; g++) {
                const int i = chunk + g * 32 + (int)lane;
                a[g] = 0;
                len[g] = 0u;
                if (i < tr.n_rows) {
                    int b;
                    row_interval(t, tr, tr.ty_first + i * tr.row_step, a[g], b);
                    len[g] = b >= a[g] ? (uint32_t)(b - a[g] + 1) : 0u;
                }
                uint32_t incl = len[g];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t up = __shfl_up_sync(FULL, incl, d);
                    if (lane >= (uint32_t)d) incl += up;
                }
                excl[g] = incl - len[g];
                tot[g] = __shfl_sync(FULL, incl, 31);
            }
            // the h-th admitted pair of group g: row = the last lane whose exclusive prefix is <= h
            auto locate = [&](int g, uint32_t h, int &tx, int &ty) {
                int lo = 0, hi = 32;
#pragma unroll
                for (int step = 0; step < 5; step++) {
                    const int mid = (lo + hi) >> 1;
                    if (__shfl_sync(FULL, excl[g], mid) <= h) lo = mid; else hi = mid;
                }
                tx = __shfl_sync(FULL, a[g], lo) + (int)(h - __shfl_sync(FULL, excl[g], lo));
                ty = tr.ty_first + (chunk + g * 32 + lo) * tr.row_step;
            };
            // count
            ListPos cnt{0u, 0u, 0u, 0u};
#pragma unroll
            for (int g = 0; g < GROUPS; g++) {
#pragma unroll 1
                for (uint32_t h0 = 0; h0 < tot[g]; h0 += 32) { // warp-uniform
                    const uint32_t h = h0 + lane;
                    int tx, ty;
                    locate(g, min(h, tot[g] - 1u), tx, ty);
                    if (h < tot[g]) cnt.count(class_of_admitted(U, tr.x0, tr.x1, tr.y0, tr.y1, tx, ty, transparent));
                }
            }
            uint32_t sum[4] = {cnt.l, cnt.m, cnt.s, cnt.t};
#pragma unroll
            for (int k = 0; k < 4; k++) sum[k] = __reduce_add_sync(FULL, sum[k]);
            if ((sum[0] | sum[1] | sum[2] | sum[3]) == 0u) continue; // warp-uniform
            uint32_t base = 0;
            if (lane < 4) {
                const uint32_t mine = lane == 0 ? sum[0] : (lane == 1 ? sum[1] : (lane == 2 ? sum[2] : sum[3]));
                const int ctr = lane == 0 ? CNT_L_PAIRS : (lane == 1 ? CNT_MEDIUM : (lane == 2 ? CNT_SMALL : CNT_T_PAIRS));
                if (mine) base = atomicAdd(W.counters + ctr, mine);
            }
            uint32_t run_l = __shfl_sync(FULL, base, 0), run_m = __shfl_sync(FULL, base, 1), run_s = __shfl_sync(FULL, base, 2),
                     run_t = __shfl_sync(FULL, base, 3);
            // write
#pragma unroll
            for (int g = 0; g < GROUPS; g++) {
#pragma unroll 1
                for (uint32_t h0 = 0; h0 < tot[g]; h0 += 32) { // warp-uniform
                    const uint32_t h = h0 + lane;
                    const bool valid = h < tot[g];
                    int tx, ty;
                    locate(g, min(h, tot[g] - 1u), tx, ty);
                    PairClass c = class_of_admitted(U, tr.x0, tr.x1, tr.y0, tr.y1, tx, ty, transparent);
                    if (!valid) c.cls = 4u;
                    const uint32_t ml = __ballot_sync(FULL, c.cls == 0u), ms = __ballot_sync(FULL, c.cls == 2u),
                                   mt = __ballot_sync(FULL, c.cls == 3u);
                    uint32_t e_incl = c.cls == 1u ? c.entries : 0u; // medium: 1-4 entries each
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t up = __shfl_up_sync(FULL, e_incl, d);
                        if (lane >= (uint32_t)d) e_incl += up;
                    }
                    const uint32_t m_tot = __shfl_sync(FULL, e_incl, 31);
                    uint32_t at = 0;
                    if (c.cls == 0u) at = run_l + (uint32_t)__popc(ml & lt_mask);
                    else if (c.cls == 1u) at = run_m + e_incl - c.entries;
                    else if (c.cls == 2u) at = run_s + (uint32_t)__popc(ms & lt_mask);
                    else if (c.cls == 3u) at = run_t + (uint32_t)__popc(mt & lt_mask);
                    if (c.cls < 4u) emit_at(W, c, slot, tx, ty, at);
                    run_l += (uint32_t)__popc(ml);
                    run_m += m_tot;
                    run_s += (uint32_t)__popc(ms);
                    run_t += (uint32_t)__popc(mt);
                }
            }
        }
    }
}

// Record slots are numbered in draw order: slot order == draw order, which is what lets k_tile break depth ties by
// comparing slots.  Inside a block: prefix sums over the per-thread output counts (0, 1, or 4 for a triangle to
// clip).  Across blocks: block b's numbers start at b * SLOT_STRIDE — no block waits for another (a chained scan
// with decoupled look-back, the first design, cost C5 a quarter of the phase in waiting) — and the records are
// stored densely wherever the block's atomicAdd put them (FrameDev::block_loc).
__device__ __forceinline__ void phase_setup(const FrameUniforms &U, const SceneDev &S, const FrameDev &W, FrontShared &sh) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_blocks = (S.n_triangles + FRONT_THREADS - 1) / FRONT_THREADS;
    // A CTA's first block is its own index (no round trip); the ticket counter hands out the blocks beyond the grid.
    bool first_block = true;
    while (true) {
        uint32_t bid = blockIdx.x;
        if (!first_block) {
            __syncthreads(); // sh.* of the previous block are no longer read
            if (threadIdx.x == 0) sh.ticket = gridDim.x + atomicAdd(&W.counters[CNT_TICKET], 1u);
            __syncthreads();
            bid = sh.ticket;
        }
        first_block = false;
        if (bid >= n_blocks) break;
        const uint32_t tri = bid * FRONT_THREADS + threadIdx.x;
        // stamps of the first block's sub-phases (debug statistics: counters[CNT_PHASE_NS + 8 ...])
        const bool stamp = bid == 0 && threadIdx.x == 0;
        uint32_t sub[6] = {0, 0, 0, 0, 0, 0};
        if (stamp) sub[0] = (uint32_t)global_timer_ns();

        uint32_t n_out = 0;       // opaque record slots this thread reserves (0, 1, or 4 for a triangle to clip)
        bool clipped = false;     // the triangle straddles the near or far plane
        bool transparent = false;
        bool alive = false;
        uint32_t tslot = 0, material = 0;
        uint32_t vi[3] = {0, 0, 0};
        RasterRec r;
        r.id = NO_SLOT;
        r.bbx = r.bby = 0;

        if (tri < S.n_triangles) {
            vi[0] = S.idx[0][tri]; vi[1] = S.idx[1][tri]; vi[2] = S.idx[2][tri];
            const uint32_t mat = S.tri_mat[tri];
            material = mat & 0x7FFFFFFFu;
            transparent = (mat >> 31) != 0;
            if (transparent) tslot = S.tri_tslot[tri];
            // both gathers of the three corners are issued together (one round trip, not two: the phase is a chain of latencies)
            const float4 pa = __ldg(S.pos4 + vi[0]), pb = __ldg(S.pos4 + vi[1]), pc = __ldg(S.pos4 + vi[2]);
            const float4 va = W.vA[vi[0]], vb = W.vA[vi[1]], vc = W.vA[vi[2]];

            alive = true;
            if (!transparent) {
                // back-face cull, scene/mod.rs:1016-1027 with calc_normal :30-41 and get_center :92-99
                const v3 a{pa.x, pa.y, pa.z}, b{pb.x, pb.y, pb.z}, c{pc.x, pc.y, pc.z};
                const v3 nrm = v_cross(v_sub(b, a), v_sub(c, b));
                v3 sum{0.0f, 0.0f, 0.0f};
                sum = v_add(sum, a);
                sum = v_add(sum, b);
                sum = v_add(sum, c);
                const v3 center = v_div(sum, 3.0f);
                const v3 eye = v_sub(v3{U.cam[0], U.cam[1], U.cam[2]}, center);
                if (v_dot(eye, nrm) <= 0.0f) alive = false;
            }
            if (alive) {
                const uint32_t fa = __float_as_uint(va.w), fb = __float_as_uint(vb.w), fc = __float_as_uint(vc.w);
                const uint32_t all_nonpos = fa & fb & fc & 0xAAAu; // bit 2p+1 : f <= 0 on all three
                const uint32_t all_pos = fa & fb & fc & 0x555u;    // bit 2p   : f > 0 on all three
                bool clip = false;
                // lateral planes 2..5: reject only if completely outside one of them (:59-66, :641-660)
                if (all_nonpos & 0xAA0u) alive = false;
                else if ((all_pos & 0x5u) == 0x5u) clip = false;   // inside near and far: passes through (:677-681)
                else if (all_nonpos & 0x2u) alive = false;          // completely behind the near plane (:682-686)
                else if ((all_pos & 0x1u) && (all_nonpos & 0x8u)) alive = false; // untouched by near, beyond far
                else clip = true;
                if (alive && clip) {
                    // An opaque triangle reserves the maximum of four consecutive slots, so that slot order stays draw order.
                    n_out = transparent ? 0u : 4u;
                    clipped = true;
                } else if (alive) {
                    const float sx[3] = {va.x, vb.x, vc.x}, sy[3] = {va.y, vb.y, vc.y}, dep[3] = {va.z, vb.z, vc.z};
                    alive = setup_raster(U, sx, sy, dep, tri * 4u, r);
                    if (alive) alive = tile_range(U, r.bbx, r.bby).count() > 0; // outside this launch's tile rows: dropped
                    n_out = (alive && !transparent) ? 1u : 0u;
                }
            }
        }

        // ---- block-wide exclusive prefix of n_out ---------------------------------------------------
        if (stamp) sub[1] = (uint32_t)global_timer_ns();
        uint32_t incl = n_out;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(FULL, incl, d);
            if (lane >= (uint32_t)d) incl += up;
        }
        if (lane == 31) sh.warp_tot[warp] = incl;
        __syncthreads();
        // ---- the block's records: slot numbers and storage ------------------------------------------------
        // Slot numbers order the frame's records by draw order (ties in depth go to the smaller slot): block b owns
        // the numbers [b * SLOT_STRIDE, (b + 1) * SLOT_STRIDE) and hands them out by prefix sum, so no block waits for
        // another one.  Storage is dense: one atomicAdd per block reserves its records' places, FrameDev::block_loc[b]
        // remembers where they start, and a record is found at block_loc[slot / SLOT_STRIDE] + slot % SLOT_STRIDE
        // (record_index, device_math.cuh).
        if (warp == 0) {
            uint32_t wt = lane < FRONT_THREADS / 32 ? sh.warp_tot[lane] : 0u;
            uint32_t wi = wt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(FULL, wi, d);
                if (lane >= (uint32_t)d) wi += up;
            }
            const uint32_t total = __shfl_sync(FULL, wi, 31);
            if (lane < FRONT_THREADS / 32) sh.warp_tot[lane] = wi - wt; // exclusive offset of each warp
            if (lane == 0) {
                const uint32_t base = total ? atomicAdd(&W.counters[CNT_RECORDS], total) : 0u; // the counter ends at the frame's record count
                sh.base = base;
                W.block_loc[bid] = base;
            }
        }
        __syncthreads();

        if (stamp) sub[2] = (uint32_t)global_timer_ns();
        const uint32_t local0 = sh.warp_tot[warp] + incl - n_out;          // < SLOT_STRIDE: at most four records per thread
        const uint32_t slot0 = bid * SLOT_STRIDE + local0, loc0 = sh.base + local0;
        if (n_out && loc0 + n_out > W.rec_cap) {
            atomicOr(&W.counters[CNT_OVERFLOW], OVERFLOW_RECORDS); // the host re-renders the frame with larger buffers
            alive = false;
            clipped = false;
        }

        // ---- unclipped survivors: raster, prepared and shading record ---------------------------------
        const bool single = alive && !clipped;
        const uint32_t slot = transparent ? tslot * 4u : slot0, loc = transparent ? tslot * 4u : loc0; // transparent records sit at their slot
        if (single) {
            PrepRec p;
            make_prep(r, p);
            store_raster((transparent ? W.t_rrec : W.rrec) + loc, r);
            store_prep((transparent ? W.t_prep : W.prep) + loc, p);
            { // the edge functions stay on chip for the binning below
                float *w = sh.tri[threadIdx.x];
#pragma unroll
                for (int k = 0; k < 3; k++) { w[k] = p.ecx[k]; w[3 + k] = p.ecy[k]; w[6 + k] = p.ek1[k]; w[9 + k] = p.ek2[k]; }
                w[12] = __uint_as_float(p.flags);
            }
            // attribute gather for the shading record (scene/mod.rs:938-1008)
            ShadeRec s;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const uint32_t v = vi[c], t = S.idx[3 + c][tri], n = S.idx[6 + c][tri];
                const float4 nn = __ldg(S.nrm4 + n), l = W.vLH[2 * v], h = W.vLH[2 * v + 1];
                const float2 uv = __ldg(S.uv2 + t);
                s.n[c][0] = nn.x; s.n[c][1] = nn.y; s.n[c][2] = nn.z;
                s.l[c][0] = l.x; s.l[c][1] = l.y; s.l[c][2] = l.z;
                s.h[c][0] = l.w; s.h[c][1] = h.x; s.h[c][2] = h.y;
                s.uv[c][0] = uv.x; s.uv[c][1] = uv.y;
            }
            s.material = material;
            s.pad[0] = s.pad[1] = 0;
            store_shade((transparent ? W.t_srec : W.srec) + loc, s);
        }
        // ---- triangles that straddle the near or far plane (rare, register-hungry: out of line) -------
        uint32_t kept = single ? 1u : 0u, bbx[4] = {r.bbx, 0, 0, 0}, bby[4] = {r.bby, 0, 0, 0};
        if (clipped) kept = clip_triangle(U, S, W, tri, vi, material, transparent, loc, bbx, bby);
        if (stamp) sub[3] = (uint32_t)global_timer_ns();

        // ---- binning: the block's records become jobs in shared memory, their (record, tile) pairs are
        // spread over the CTA's threads ----------------------------------------------------------------
        // Fast path — every record of the block is an unclipped triangle inside ONE tile (micro-triangle scenes; most
        // blocks of any scene with small triangles): a thread bins its own record from registers; no job table, no
        // pair search.  One count, one reservation per CTA and class, one write.
        const TileRange own = tile_range(U, bbx[0], bby[0]);
        const bool simple = kept == 0u || (kept == 1u && single && own.count() == 1);
        if (__syncthreads_and(simple)) {
            PairClass c;
            c.cls = 4u;
            c.blocks = c.entries = c.tile = 0u;
            if (kept) {
                BinTri t;
                const float *w = sh.tri[threadIdx.x];
#pragma unroll
                for (int k = 0; k < 3; k++) { t.ecx[k] = w[k]; t.ecy[k] = w[3 + k]; t.ek1[k] = w[6 + k]; t.ek2[k] = w[9 + k]; }
                t.flags = __float_as_uint(w[12]);
                c = classify_pair(U, t, own.x0, own.x1, own.y0, own.y1, own.tx0, own.ty_first, transparent);
            }
            ListPos pos{0u, 0u, 0u, 0u};
            pos.count(c);
            reserve_lists(W, sh, pos);
            emit_pair(W, c, slot, own.tx0, own.ty_first, pos);
            if (stamp) {
                sub[4] = sub[5] = (uint32_t)global_timer_ns();
#pragma unroll
                for (int i = 0; i < 5; i++) W.counters[CNT_PHASE_NS + 8 + i] = sub[i + 1] - sub[i];
            }
            continue;
        }
        uint32_t n_jobs;
        uint32_t jb = block_exclusive((uint32_t)__popc(kept), sh, &n_jobs);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if ((kept >> k) & 1u) {
                const uint32_t slot_t = (slot + (uint32_t)k) | (transparent ? 0x80000000u : 0u);
                const uint32_t nt = (uint32_t)tile_range(U, bbx[k], bby[k]).count();
                if (nt > HUGE_TILES) {
                    // diverted: position in the queue and first tile row of the record from ONE 64-bit counter (jobs << 32 | rows),
                    // so that row bases grow with the queue position
                    const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(W.counters + CNT_HUGE),
                                                             (1ull << 32) | (uint32_t)tile_range(U, bbx[k], bby[k]).n_rows);
                    const uint32_t pos = (uint32_t)(old >> 32);
                    if (pos < W.huge_cap) W.huge_jobs[pos] = make_uint4(bbx[k], bby[k], slot_t, (uint32_t)old);
                    else atomicOr(&W.counters[CNT_OVERFLOW], OVERFLOW_HUGE);
                    sh.jobs[jb++] = BinJob{0u, 0u, slot_t, DIVERTED}; // keeps its place in the block's list, with no tiles
                } else {
                    sh.jobs[jb++] = BinJob{bbx[k], bby[k], slot_t, single ? threadIdx.x : NO_TRI};
                }
            }
        __syncthreads(); // jobs complete; the records written above are visible to the whole CTA
        uint32_t n_tiles[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t j = threadIdx.x * 4u + (uint32_t)k;
            n_tiles[k] = (j < n_jobs && sh.jobs[j].tri_idx != DIVERTED) ? (uint32_t)tile_range(U, sh.jobs[j].bbx, sh.jobs[j].bby).count() : 0u;
        }
        bin_block(U, W, sh, n_tiles);
        if (stamp) {
            sub[4] = sub[5] = (uint32_t)global_timer_ns();
#pragma unroll
            for (int i = 0; i < 5; i++) W.counters[CNT_PHASE_NS + 8 + i] = sub[i + 1] - sub[i]; // set-up, scan, records + clip, binning
        }
    }
}

// ------------------------------------------------------------------------------------------
// P2 / P3 : per tile
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int cost_bucket(uint32_t cost) { // 0 = heaviest
    return cost == 0 ? COST_BUCKETS - 1 : __clz(cost);      // clz in 0..31 (larger cost -> smaller clz)
}
// Number of windows a tile of this cost is cut into: the largest power of two <= cost / target,
// so that the sum over tiles stays <= total cost / target <= TILE_EXTRA_ITEMS.
__device__ __forceinline__ uint32_t tile_splits(uint32_t cost, uint32_t target, uint32_t max_split) {
    if (!TILE_SPLITTABLE || cost < 2u * target) return 1u;
    const uint32_t q = cost / target;
    return min(max_split, 1u << (31 - __clz(q)));
}

__device__ __forceinline__ void phase_alloc(const FrameUniforms &U, const FrameDev &W, FrontShared &sh) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nc = U.n_coarse;
    // estimated total cost of the frame's tiles -> target cost of a work item
    const uint32_t n_l = __ldcg(&W.counters[CNT_L_PAIRS]), n_t = __ldcg(&W.counters[CNT_T_PAIRS]);
    const uint32_t total_cost = COST_LARGE * n_l + COST_TRANSPARENT * n_t + COST_TILE_BASE * rows_mine(U) * U.tiles_x;
    const uint32_t target = max(U.split_min_cost, total_cost / U.split_div + 1u); // split_div <= TILE_EXTRA_ITEMS
    for (uint32_t base = blockIdx.x * FRONT_THREADS; base < nc; base += gridDim.x * FRONT_THREADS) { // block-uniform trip count
        const uint32_t tile = base + tid;
        const bool valid = tile < nc;
        const uint32_t ty = valid ? tile / U.tiles_x : 0u, tx = valid ? tile % U.tiles_x : 0u;
        const bool mine = valid && row_is_mine(U, ty);
        const uint32_t c0 = mine ? __ldcg(&W.l_count[tile]) : 0u, c1 = mine ? __ldcg(&W.t_count[tile]) : 0u, ms = mine ? __ldcg(&W.ms_weight[tile]) : 0u;
        const bool nonempty = (c0 | c1 | ms) != 0u;
        // exclusive prefixes of the two counts over the CTA, one range reservation per class
        uint32_t inc0 = c0, inc1 = c1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u0 = __shfl_up_sync(FULL, inc0, d), u1 = __shfl_up_sync(FULL, inc1, d);
            if (lane >= (uint32_t)d) { inc0 += u0; inc1 += u1; }
        }
        __syncthreads(); // sh.sum4 / base4 of the previous iteration are no longer read
        if (lane == 31) { sh.sum4[0][warp] = inc0; sh.sum4[1][warp] = inc1; }
        __syncthreads();
        if (tid < 2) {
            uint32_t total = 0;
            for (int w = 0; w < FRONT_THREADS / 32; w++) {
                const uint32_t t = sh.sum4[tid][w];
                sh.sum4[tid][w] = total;
                total += t;
            }
            sh.base4[tid] = total ? atomicAdd(&W.counters[tid == 0 ? CNT_L_CURSOR : CNT_T_CURSOR], total) : 0u;
        }
        __syncthreads();
        // cost, windows, bucket; rank inside the bucket with one atomic per (warp, bucket)
        uint32_t cost = 0, splits = 0;
        int bucket = -1;
        if (nonempty) {
            cost = min(COST_TILE_BASE + COST_LARGE * c0 + COST_TRANSPARENT * c1, 0x0FFFFFFFu);
            splits = tile_splits(cost, target, U.split_max);
            bucket = cost_bucket(cost / splits);
        }
        uint32_t before = 0, group_total = 0;
        int leader = -1;
#pragma unroll 1
        for (int j = 0; j < 32; j++) {
            const int bj = __shfl_sync(FULL, bucket, j);
            const uint32_t sj = __shfl_sync(FULL, splits, j);
            if (bj == bucket && bucket >= 0) {
                if (leader < 0) leader = j;
                if (j < (int)lane) before += sj;
                group_total += sj;
            }
        }
        uint32_t rank = 0;
        if (bucket >= 0 && leader == (int)lane) rank = atomicAdd(&W.counters[CNT_BUCKETS + bucket], group_total);
        rank = __shfl_sync(FULL, rank, leader < 0 ? 0 : leader) + before;
        // empty tiles of this launch's rows are listed on their own (k_tile's CTAs write the clear colour and depth)
        const bool empty = mine && !nonempty;
        const uint32_t eb = __ballot_sync(FULL, empty);
        if (eb) {
            uint32_t wbase = 0;
            const int el = __ffs(eb) - 1;
            if ((int)lane == el) wbase = atomicAdd(&W.counters[CNT_EMPTY], (uint32_t)__popc(eb));
            wbase = __shfl_sync(FULL, wbase, el);
            if (empty) W.empty_tiles[wbase + __popc(eb & ((1u << lane) - 1u))] = tx | ty << 10;
        }
        if (valid) {
            W.l_offset[tile] = sh.base4[0] + sh.sum4[0][warp] + inc0 - c0;
            W.t_offset[tile] = sh.base4[1] + sh.sum4[1][warp] + inc1 - c1;
            W.l_count[tile] = 0; // become the fill cursors of the scatter (k_raster's prologue)
            W.t_count[tile] = 0;
        }
        // the tile's work items, straight into its bucket's segment of the work list (device_types.h)
        if (nonempty) {
            uint32_t *seg = W.tile_order + (size_t)bucket * W.bucket_cap + rank;
            for (uint32_t i = 0; i < splits; i++) seg[i] = make_item(tx, ty, splits, i);
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_front
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FRONT_THREADS, 8) k_front(const FrameUniforms *__restrict__ Up, const SceneDev S, const FrameDev W) {
    const FrameUniforms &U = *Up; // per-frame uniforms, device-resident (one upload per frame; the launches never change)
    __shared__ FrontShared sh;
    const CtaTrace trace_(W, 1u);
    const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
    uint32_t stamps[6] = {0, 0, 0, 0, 0, 0};
    if (stamp) stamps[0] = (uint32_t)global_timer_ns();
    phase_vertex(U, S, W);
    if (stamp) stamps[1] = (uint32_t)global_timer_ns();
    grid_barrier(W.counters, U.bar_base + gridDim.x);
    if (stamp) stamps[2] = (uint32_t)global_timer_ns();
    if (S.n_triangles) phase_setup(U, S, W, sh);
    if (stamp) stamps[3] = (uint32_t)global_timer_ns();
    if (threadIdx.x == 0) atomicMax(W.counters + CNT_PHASE_NS + 6, (uint32_t)global_timer_ns()); // when the slowest CTA left the phase
    grid_barrier(W.counters, U.bar_base + 2u * gridDim.x);
    {
        // huge records, if the frame has any: their pairs are binned by the whole grid, then a third barrier.  Without them
        // the CTAs only arrive at that barrier (the counter's value stays what the host expects) and go on.
        const unsigned long long hq = __ldcg(reinterpret_cast<const unsigned long long *>(W.counters + CNT_HUGE));
        const uint32_t n_huge = min((uint32_t)(hq >> 32), W.huge_cap);
        if (n_huge) { // the same value in every CTA: final since the barrier above
            phase_huge(U, W, n_huge);
            if (threadIdx.x == 0) atomicMax(W.counters + CNT_PHASE_NS + 7, (uint32_t)global_timer_ns());
            grid_barrier(W.counters, U.bar_base + 3u * gridDim.x);
        } else {
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(W.counters + CNT_BARRIER, 1u);
        }
    }
    if (stamp) stamps[4] = (uint32_t)global_timer_ns();
    phase_alloc(U, W, sh);
    if (stamp) {
        stamps[5] = (uint32_t)global_timer_ns();
#pragma unroll
        for (int i = 0; i < 6; i++) W.counters[CNT_PHASE_NS + i] = stamps[i];
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
thread_local unsigned g_front_ctas = 148u; // scene.cpp: by scene size (DRAW_B200_FRONT_CPS)
uint32_t tile_grid_items(const FrameUniforms &U) { // k_tile's work-list slots
    const uint32_t tiles = rows_mine(U) * U.tiles_x;
    return tiles ? tiles + (TILE_SPLITTABLE ? TILE_EXTRA_ITEMS : 0) : 0;
}
int front_max_ctas_per_sm() {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_front, FRONT_THREADS, 0) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    return n < 1 ? 1 : n;
}
cudaError_t launch_front(const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g_front_ctas);
    cfg.blockDim = dim3(FRONT_THREADS);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_front, dU, S, W);
}

} // namespace drawb200
