// shading.cuh — per-pixel Phong + nearest-texel shading (mororo18/draw canvas.rs:673-743).
#pragma once
#include "device_math.cuh"

namespace drawb200 {

// u8 -> (u8 as f32) / 255.0 (Pixel::normalized_as_vec3, canvas.rs:81-87) as a 256-entry table that
// each CTA fills in shared memory with the same IEEE division, so a lookup is bit-identical to the
// reference's divide and six divisions per shaded pixel disappear.
__device__ __forceinline__ void fill_u8_table(float *tab, int tid, int n_threads) {
    for (int i = tid; i < 256; i += n_threads) tab[i] = FDIV((float)i, 255.0f);
}

// TextureMap::get_rgb_slice (scene/mod.rs:154-168).  Maps are stored with 4 bytes per texel on the
// device (3-component images are padded at upload), so a texel is one aligned 32-bit load through
// the read-only path; only bytes 0..2 (r, g, b) are used, as in the reference.  Indices are clamped
// into the map (SURVEY.md §8c deviation 6; never triggers for uv in [0,1]).
__device__ __forceinline__ uint32_t fetch_texel(const uint8_t *__restrict__ texels, uint32_t off, uint32_t w, uint32_t h,
                                                float u, float v) {
    // Rust's `as usize` saturates and maps NaN to 0; so does cvt.rzi.u32, and whatever exceeds 2^32 - 1
    // is clamped to the last texel below either way
    uint32_t ui = __float2uint_rz(floorf(FMUL(u, FSUB((float)w, 1.0f))));
    uint32_t vr = __float2uint_rz(floorf(FMUL(v, FSUB((float)h, 1.0f))));
    if (ui > w - 1) ui = w - 1;
    if (vr > h - 1) vr = h - 1;
    const uint32_t *p = reinterpret_cast<const uint32_t *>(texels + off) + ((size_t)(h - 1 - vr) * w + ui);
    return __ldg(p);
}

// canvas.rs:685-743 for one covered pixel whose barycentrics are known: interpolation, texel fetches
// and Phong.  Returns r | g << 8 | b << 16.
// MATERIALS_CACHED: `materials` is a copy of the table in shared memory (plain loads) instead of the
// global table (read-only path).
// RECORDS_ON_CHIP: the records are copies in shared memory (plain loads) instead of global memory (read-only path).
template <bool MATERIALS_CACHED = false, bool RECORDS_ON_CHIP = false>
__device__ __forceinline__ uint32_t shade_bary(const MaterialDev *materials, const uint8_t *__restrict__ texels,
                                               const float *u8tab, const ShadeRec *sp, float alpha, float beta,
                                               float gama, float *opacity_out) {
    // ShadeRec as 9 x uint4: n[3][3] l[3][3] h[3][3] uv[3][2] material pad pad
    const uint4 *q = reinterpret_cast<const uint4 *>(sp);
    float w[36];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const uint4 t = RECORDS_ON_CHIP ? q[i] : __ldg(q + i);
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
    // MaterialDev as 4 x uint4: ka3 kd1 | kd2 ks3 | alpha ka_off ka_w ka_h | kd_off kd_w kd_h pad
    const uint4 *mq = reinterpret_cast<const uint4 *>(materials + __float_as_uint(w[33]));
    const uint4 m0 = MATERIALS_CACHED ? mq[0] : __ldg(mq), m1 = MATERIALS_CACHED ? mq[1] : __ldg(mq + 1),
                m2 = MATERIALS_CACHED ? mq[2] : __ldg(mq + 2), m3 = MATERIALS_CACHED ? mq[3] : __ldg(mq + 3);
    const v3 ka{__uint_as_float(m0.x), __uint_as_float(m0.y), __uint_as_float(m0.z)};
    const v3 kd{__uint_as_float(m0.w), __uint_as_float(m1.x), __uint_as_float(m1.y)};
    const v3 ks{__uint_as_float(m1.z), __uint_as_float(m1.w), __uint_as_float(m2.x)};
    *opacity_out = __uint_as_float(m2.y);

    // A 1x1 map (TextureMap::default(), scene/mod.rs:128-135: every material without an image) has one texel, whatever
    // u and v are: floor(u * (1 - 1)) is 0, or NaN -> 0 under `as usize`.  When both maps are 1x1 the texture
    // coordinate is not interpolated at all (it feeds nothing else, canvas.rs:685-695).
    uint32_t dt, at;
    if ((m3.y | m3.z | m2.w | m3.w) == 1u) { // widths and heights are >= 1
        dt = __ldg(reinterpret_cast<const uint32_t *>(texels + m3.x));
        at = __ldg(reinterpret_cast<const uint32_t *>(texels + m2.z));
    } else {
        const float u = INTERP(27, 0, 2), v = INTERP(27, 1, 2);
        dt = fetch_texel(texels, m3.x, m3.y, m3.z, u, v); // map_kd   canvas.rs:689-691
        at = fetch_texel(texels, m2.z, m2.w, m3.w, u, v); // map_ka   canvas.rs:693-695
    }
#undef INTERP
    const v3 dcol{u8tab[dt & 255u], u8tab[(dt >> 8) & 255u], u8tab[(dt >> 16) & 255u]};
    const v3 acol{u8tab[at & 255u], u8tab[(at >> 8) & 255u], u8tab[(at >> 16) & 255u]};

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

// canvas.rs:673-682 from the raster record: literal barycentrics (edge functions rebuilt from the snapped
// vertices) and the interpolated depth; then shade_bary.  Used where no prepared record exists (transparent
// triangles).
template <bool MATERIALS_CACHED = false>
__device__ __forceinline__ uint32_t shade_pixel(const MaterialDev *materials, const uint8_t *__restrict__ texels,
                                                const float *u8tab, const RasterRec &r, const ShadeRec *__restrict__ sp,
                                                float x, float y, float *depth_out, float *opacity_out) {
    const Edge e_bc = make_edge(r.bx, r.by, r.cx, r.cy), e_ca = make_edge(r.cx, r.cy, r.ax, r.ay),
               e_ab = make_edge(r.ax, r.ay, r.bx, r.by);
    const float alpha = FDIV(edge_eval(e_bc, x, y), edge_eval(e_bc, r.ax, r.ay));
    const float beta = FDIV(edge_eval(e_ca, x, y), edge_eval(e_ca, r.bx, r.by));
    const float gama = FDIV(edge_eval(e_ab, x, y), edge_eval(e_ab, r.cx, r.cy));
    *depth_out = FADD(FADD(FMUL(alpha, r.da), FMUL(beta, r.db)), FMUL(gama, r.dc));
    return shade_bary<MATERIALS_CACHED>(materials, texels, u8tab, sp, alpha, beta, gama, opacity_out);
}

// The same from the prepared record (opaque triangles): its edge functions are the raster record's, sign-
// normalised together with f (the quotients are bit-identical, device_math.cuh), and under TRI_FASTDIV the
// division is exact_div.  ~55 instructions fewer per pixel than rebuilding the edges.  *id_out = draw id.
template <bool MATERIALS_CACHED = false, bool RECORDS_ON_CHIP = false>
__device__ __forceinline__ uint32_t shade_pixel_prep(const MaterialDev *materials, const uint8_t *__restrict__ texels,
                                                     const float *u8tab, const PrepRec *pp, const ShadeRec *sp, float x, float y,
                                                     float *depth_out, float *opacity_out, uint32_t *id_out) {
    const uint4 *q = reinterpret_cast<const uint4 *>(pp);
    uint4 q0, q1, q2, q3, q4, q6;
    if (RECORDS_ON_CHIP) {
        q0 = q[0]; q1 = q[1]; q2 = q[2]; q3 = q[3]; q4 = q[4]; q6 = q[6];
    } else {
        q0 = __ldg(q); q1 = __ldg(q + 1); q2 = __ldg(q + 2); q3 = __ldg(q + 3); q4 = __ldg(q + 4); q6 = __ldg(q + 6);
    }
    const float ecx[3] = {__uint_as_float(q0.x), __uint_as_float(q0.y), __uint_as_float(q0.z)};
    const float ecy[3] = {__uint_as_float(q0.w), __uint_as_float(q1.x), __uint_as_float(q1.y)};
    const float ek1[3] = {__uint_as_float(q1.z), __uint_as_float(q1.w), __uint_as_float(q2.x)};
    const float ek2[3] = {__uint_as_float(q2.y), __uint_as_float(q2.z), __uint_as_float(q2.w)};
    const float f[3] = {__uint_as_float(q3.x), __uint_as_float(q3.y), __uint_as_float(q3.z)};
    const float rf[3] = {__uint_as_float(q3.w), __uint_as_float(q4.x), __uint_as_float(q4.y)};
    float bary[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float e = FSUB(FADD(FADD(FMUL(ecx[i], x), FMUL(ecy[i], y)), ek1[i]), ek2[i]);
        bary[i] = (q6.y & TRI_FASTDIV) ? exact_div(e, f[i], rf[i]) : FDIV(e, f[i]);
    }
    *depth_out = FADD(FADD(FMUL(bary[0], __uint_as_float(q4.z)), FMUL(bary[1], __uint_as_float(q4.w))), FMUL(bary[2], __uint_as_float(q6.x)));
    *id_out = q6.z;
    return shade_bary<MATERIALS_CACHED, RECORDS_ON_CHIP>(materials, texels, u8tab, sp, bary[0], bary[1], bary[2], opacity_out);
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

} // namespace drawb200
