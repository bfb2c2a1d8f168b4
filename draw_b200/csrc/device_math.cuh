// device_math.cuh — strict float32 device arithmetic shared by all kernels.
//
// Numerical contract: every float operation is a single
// round-to-nearest binary32 operation in the reference's evaluation order (mororo18/draw
// src/renderer/linalg.rs).  Everything goes through __fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn/
// __fsqrt_rn, which ptxas never fuses into FMAs or reassociates; the .cu files are also
// compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_types.h"

namespace drawb200 {

// Debug timeline (draw_scene_debug_trace): thread 0 of every CTA of the frame kernels records
// (kernel id | SM << 8 | work set << 24, CTA index, start, end) with the nanosecond global timer, so that
// the overlap of the frames in flight can be looked at without a system profiler.  Off (null pointer)
// it costs one uniform branch per CTA.
struct CtaTrace {
    uint4 *trace;
    uint32_t *count;
    uint32_t cap, word, t0;
    static __device__ __forceinline__ uint32_t now() {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return (uint32_t)t;
    }
    __device__ __forceinline__ CtaTrace(const FrameDev &W, uint32_t kernel_id) : trace(W.trace), count(W.trace_count), cap(W.trace_cap), word(0), t0(0) {
        if (trace && threadIdx.x == 0) {
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            word = kernel_id | smid << 8 | W.trace_tag << 24;
            t0 = now();
        }
    }
    __device__ __forceinline__ ~CtaTrace() {
        if (trace && threadIdx.x == 0) {
            const uint32_t t1 = now(), at = atomicAdd(count, 1u);
            if (at < cap) trace[at] = make_uint4(word, blockIdx.x, t0, t1);
        }
    }
};

// Storage index of the opaque record with this slot number (device_types.h: SLOT_STRIDE).  record_index reads the
// table through the read-only path (kernels after k_front); record_index_cg is for k_front itself, which writes it.
__device__ __forceinline__ uint32_t record_index(const FrameDev &W, uint32_t slot) {
    return __ldg(W.block_loc + (slot >> SLOT_SHIFT)) + (slot & (SLOT_STRIDE - 1u));
}
__device__ __forceinline__ uint32_t record_index_cg(const FrameDev &W, uint32_t slot) {
    return __ldcg(W.block_loc + (slot >> SLOT_SHIFT)) + (slot & (SLOT_STRIDE - 1u));
}

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
// Rust `f32 as u8` (saturating, NaN -> 0)
__device__ __forceinline__ uint32_t sat_u8(float v) {
    const uint32_t u = __float2uint_rz(v);
    return u > 255u ? 255u : u;
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

// The three edge functions of a screen triangle prepared for coverage tests.
//   edge 0 = bc (alpha, opposite vertex a), 1 = ca (beta, b), 2 = ab (gama, c)   canvas.rs:660-666
// "Tame" triangles (all coefficients finite and < 1e30, so no product with a pixel coordinate can
// overflow and every edge value is an exact-or-rounded integer) are sign-normalised: when f < 0
// every coefficient of that edge is negated.  Negation commutes with round-to-nearest, so the
// edge value e' is exactly -e, f' = -f > 0, the quotient e'/f' is bit-identical to e/f, and
//   e/f >= 0  <=>  e' >= 0        e/f > 0  <=>  e' > 0
// (|e'| >= 1 when non-zero, f' < 2^100: the quotient cannot underflow to zero).
// Other triangles keep the literal coefficients and are evaluated with the reference's division
// ("slow" mode, flag bit 3).
struct TriEdges {
    float ecx[3], ecy[3], ek1[3], ek2[3], f[3];
    uint32_t flags; // bit e: f_e * f_e(-1,-1) > 0, the tie rule admits e == 0 (canvas.rs:678-680); bit 3: slow
};
constexpr uint32_t TRI_SLOW = 8u;

__device__ __forceinline__ TriEdges prepare_edges(const RasterRec &r) {
    const Edge e[3] = {make_edge(r.bx, r.by, r.cx, r.cy), make_edge(r.cx, r.cy, r.ax, r.ay),
                       make_edge(r.ax, r.ay, r.bx, r.by)};
    const float vx[3] = {r.ax, r.bx, r.cx}, vy[3] = {r.ay, r.by, r.cy};
    TriEdges t;
    t.flags = 0;
    bool tame = true;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        t.f[i] = edge_eval(e[i], vx[i], vy[i]);
        const float f_out = edge_eval(e[i], -1.0f, -1.0f);
        if (FMUL(t.f[i], f_out) > 0.0f) t.flags |= 1u << i;
        const float lim = 1e30f;
        tame = tame && fabsf(e[i].cx) < lim && fabsf(e[i].cy) < lim && fabsf(e[i].k1) < lim &&
               fabsf(e[i].k2) < lim && fabsf(t.f[i]) < lim;
    }
    if (!tame) t.flags |= TRI_SLOW;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const bool neg = tame && t.f[i] < 0.0f;
        t.ecx[i] = neg ? -e[i].cx : e[i].cx;
        t.ecy[i] = neg ? -e[i].cy : e[i].cy;
        t.ek1[i] = neg ? -e[i].k1 : e[i].k1;
        t.ek2[i] = neg ? -e[i].k2 : e[i].k2;
        t.f[i] = neg ? -t.f[i] : t.f[i];
    }
    return t;
}

// ---- prepared records --------------------------------------------------------------------------
// Early depth reject.  For a covered pixel of a tame triangle with finite non-negative vertex depths,
//   D = e0*da/f0 + e1*db/f1 + e2*dc/f2                       (real arithmetic, all terms >= 0)
// and the reference's float depth d_ref = fl(fl(fl(e0/f0)*da + fl(e1/f1)*db) + fl(e2/f2)*dc) satisfies
// |d_ref - D| <= 4.1 u D (u = 2^-24: one division, one product and two sums of non-negative terms),
// while d_app = fma(e2, g2, fma(e1, g1, e0*g0)) with g_i = fl(d_i * fl(1/f_i)) satisfies
// |d_app - D| <= 5.1 u D.  Hence d_ref >= d_app * (1 - 9.3u) > d_app * (1 - 2^-20), so
//   d_app * (1 - 2^-20) > z   ==>   d_ref > z :  the fragment fails the strict `<` test and is not a tie,
// and its divisions can be skipped.  The bounds need normal (not denormal) products, so the
// flag is only set when every g_i is 0 or >= 1e-30; NaN/inf make the comparison false or d_ref = inf.
constexpr uint32_t TRI_EARLYZ = 16u;
constexpr float EARLYZ_SCALE = 0.99999904632568359375f; // 1 - 2^-20
// exact_div may replace the IEEE division e / f_i for this triangle (tame, 1 <= f_i <= 2^40)
constexpr uint32_t TRI_FASTDIV = 32u;

// Correctly rounded e / f from rf = RN(1/f) in three operations (Markstein's theorem: with a correctly
// rounded reciprocal, q0 = RN(e*rf) is within one ulp of e/f, the remainder r = e - f*q0 is exact in
// an FMA, and RN(q0 + r*rf) is the correctly rounded quotient).  No branches, so the three quotients
// of a pixel — and those of neighbouring pixels — overlap in the pipeline, unlike __fdiv_rn's
// subroutine.  Used only under TRI_FASTDIV (no overflow, underflow or special values in reach);
// tests/test_fastdiv_cpu.py checks the identity on 10^8 near-midpoint quotients, and every GPU
// parity test exercises it.  The FMAs here are deliberate and exact-result-preserving.
__device__ __forceinline__ float exact_div(float e, float f, float rf) {
    const float q0 = __fmul_rn(e, rf);
    const float r = __fmaf_rn(-f, q0, e);
    return __fmaf_rn(r, rf, q0);
}

__device__ __forceinline__ void make_prep(const RasterRec &r, PrepRec &p) {
    const TriEdges t = prepare_edges(r);
    const float dep[3] = {r.da, r.db, r.dc};
    bool earlyz = !(t.flags & TRI_SLOW), fastdiv = earlyz;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        p.ecx[i] = t.ecx[i]; p.ecy[i] = t.ecy[i]; p.ek1[i] = t.ek1[i]; p.ek2[i] = t.ek2[i]; p.f[i] = t.f[i];
        p.rf[i] = __frcp_rn(t.f[i]);
        const float g = __fmul_rn(dep[i], p.rf[i]);
        earlyz = earlyz && dep[i] >= 0.0f && dep[i] < 3.0e38f && (g == 0.0f || g >= 1e-30f);
        fastdiv = fastdiv && t.f[i] >= 1.0f && t.f[i] <= 1099511627776.0f;
    }
    p.da = r.da; p.db = r.db; p.dc = r.dc;
    p.x0 = (float)(r.bbx & 0xFFFF); p.x1 = (float)(r.bbx >> 16);
    p.y0 = (float)(r.bby & 0xFFFF); p.y1 = (float)(r.bby >> 16);
    p.flags = t.flags | (earlyz ? TRI_EARLYZ : 0u) | (fastdiv ? TRI_FASTDIV : 0u);
    p.id = r.id;
    p.slot = 0;
    p.pad[0] = p.pad[1] = p.pad[2] = p.pad[3] = 0;
}
__device__ __forceinline__ void store_prep(PrepRec *dst, const PrepRec &p) {
    const uint4 *src = reinterpret_cast<const uint4 *>(&p);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = 0; i < 7; i++) d[i] = src[i]; // the last quad is padding
}

// One triangle in registers.
struct TriRegs {
    float ecx[3], ecy[3], ek1[3], ek2[3], f[3], rf[3];
    float da, db, dc;
    uint32_t flags;
};
__device__ __forceinline__ TriRegs tri_from_prep(const PrepRec &p) {
    TriRegs t;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        t.ecx[i] = p.ecx[i]; t.ecy[i] = p.ecy[i]; t.ek1[i] = p.ek1[i]; t.ek2[i] = p.ek2[i]; t.f[i] = p.f[i]; t.rf[i] = p.rf[i];
    }
    t.da = p.da; t.db = p.db; t.dc = p.dc;
    t.flags = p.flags;
    return t;
}

// Coverage + depth of one pixel (canvas.rs:673-682).  Tame triangles: sign tests on the
// sign-normalised edge values, divisions only for covered pixels.  Others: the reference's literal
// divide-then-compare.  The quotients e/f are the same either way.
__device__ __forceinline__ bool cover_pixel(const TriRegs &t, uint32_t flags, float x, float y, float &depth) {
    float e[3];
#pragma unroll
    for (int i = 0; i < 3; i++) e[i] = FSUB(FADD(FADD(FMUL(t.ecx[i], x), FMUL(t.ecy[i], y)), t.ek1[i]), t.ek2[i]);
    float bary[3];
    if (!(flags & TRI_SLOW)) {
        const bool in = (e[0] > 0.0f || (e[0] == 0.0f && (flags & 1u))) && (e[1] > 0.0f || (e[1] == 0.0f && (flags & 2u))) &&
                        (e[2] > 0.0f || (e[2] == 0.0f && (flags & 4u)));
        if (!in) return false;
        if (flags & TRI_FASTDIV) {
#pragma unroll
            for (int i = 0; i < 3; i++) bary[i] = exact_div(e[i], t.f[i], t.rf[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 3; i++) bary[i] = FDIV(e[i], t.f[i]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) bary[i] = FDIV(e[i], t.f[i]);
        if (!(bary[0] >= 0.0f && bary[1] >= 0.0f && bary[2] >= 0.0f)) return false;
        if (!((bary[0] > 0.0f || (flags & 1u)) && (bary[1] > 0.0f || (flags & 2u)) && (bary[2] > 0.0f || (flags & 4u))))
            return false;
    }
    depth = FADD(FADD(FMUL(bary[0], t.da), FMUL(bary[1], t.db)), FMUL(bary[2], t.dc)); // canvas.rs:682
    return true;
}

// Order-preserving map float -> uint32 (-0 is folded onto +0: the reference's `<` treats them as
// equal, so the earlier draw must win between them), and its inverse.
__device__ __forceinline__ uint32_t depth_key(float d) {
    const uint32_t b = __float_as_uint(FADD(d, 0.0f));
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float depth_from_key(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
__device__ __forceinline__ unsigned long long make_key(float d, uint32_t slot) {
    return ((unsigned long long)depth_key(d) << 32) | slot;
}

// Can any pixel of the rectangle [lx,hx] x [ly,hy] be covered?  Exact, not heuristic: each edge
// value fl(fl(fl(cx*x + cy*y) + k1) - k2) is a monotone function of x and of y because every
// rounding step is monotone, so its maximum over the rectangle sits at the corner selected by
// the coefficient signs; if that corner fails an edge test, every pixel of the rectangle fails.
template <typename Tri> // TriEdges or any struct with the same ecx / ecy / ek1 / ek2 / flags members
__device__ __forceinline__ bool rect_may_cover(const Tri &t, float lx, float hx, float ly, float hy) {
    if (t.flags & TRI_SLOW) return true;
    bool any = true;
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const float xm = t.ecx[e] >= 0.0f ? hx : lx, ym = t.ecy[e] >= 0.0f ? hy : ly;
        const float em = FSUB(FADD(FADD(FMUL(t.ecx[e], xm), FMUL(t.ecy[e], ym)), t.ek1[e]), t.ek2[e]);
        any = any && (em > 0.0f || (em == 0.0f && (t.flags & (1u << e))));
    }
    return any;
}
__device__ __forceinline__ bool rect_may_cover(const TriEdges &t, int lx, int hx, int ly, int hy) {
    return rect_may_cover(t, (float)lx, (float)hx, (float)ly, (float)hy);
}

} // namespace drawb200
