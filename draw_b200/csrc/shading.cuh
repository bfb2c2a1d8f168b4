// shading.cuh — per-pixel Phong + nearest-texel shading (mororo18/draw canvas.rs:673-743).
#pragma once
#include "device_math.cuh"

namespace drawb200 {


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

} // namespace drawb200
