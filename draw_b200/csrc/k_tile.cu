// k_tile.cu — per-tile raster / depth / shade kernel (mororo18/draw canvas.rs:577-750, 906-960).
//
// One CTA (512 threads) per 64x32-pixel tile.  The tile's depth, winning record and colour stay on
// chip (registers, then shared memory) for the whole kernel; colour and depth go to HBM exactly
// once at the end, with the clear fused in.  No tensor cores: nothing here is a contraction.
//
//   phase A  "large" list: triangles are staged through shared memory 64 at a time; every lane owns
//            a 4x1 pixel block (warp = 16x8 region) and tests it against each triangle, after a
//            warp-level bbox reject and an exact block-level edge reject.  Depth/winner in registers.
//   merge    each lane publishes its 8 pixels as 64-bit keys (depth, slot) in shared memory.
//   phase B  "medium" list (bbox in the tile <= 1024 px): a coarse pass tests every 8x4 block of every
//            triangle's bbox (one thread per block, exact test) and queues the blocks that can be
//            covered; a fine pass takes queued blocks, one pixel per lane.  "small" list (<= 8 px): one
//            triangle per lane.  Covered fragments are committed with a shared-memory atomicMin on the key.
//   phase C  deferred shading, one pixel per lane per step: only the winner of a pixel is shaded
//            (canvas.rs:685-743); the key becomes (exact depth, draw id), colour goes to smem.
//   phase D  transparent triangles in draw order, blended over the shaded colour (rare).
//   phase E  write-back, 128 B per warp store: colour rows y-flipped (canvas.rs:955-956), depth
//            rows not (canvas.rs:413-423).
//
// Draw-order semantics without ordered lists: the reference draws triangles sequentially with a
// strict `<` depth test (canvas.rs:923), so for opaque triangles the surviving fragment of a pixel
// is the minimum of (depth, draw order) — ties go to the earlier triangle.  Record slots are
// allocated in draw order by k_setup, so the key (depth, slot) ordered as an unsigned 64-bit
// integer is exactly that minimum, and lists can be consumed in any order by any lane.
// Transparent triangles (depth test on, depth write off, blend with the current colour,
// scene/mod.rs:1088) are only visible over the final opaque winner W of their pixel if drawn after
// it: fragment T is blended iff id(T) > id(W) and depth(T) < depth(W), in draw order.
#include "shading.cuh"

namespace drawb200 {

#ifndef DRAW_TILE_THREADS
#define DRAW_TILE_THREADS 512
#endif
constexpr int TILE_THREADS = DRAW_TILE_THREADS;
constexpr int CHUNK = 64; // large-list triangles staged per round
constexpr int TILE_PIXELS = TILE_W * TILE_H;
// phase A geometry: a warp owns a REGION x REGION_H rectangle (4 lanes across, 8 down), a lane a
// 4 x BLK_H block of it
constexpr int BLK_H = TILE_PIXELS / TILE_THREADS / 4;
constexpr int PX = 4 * BLK_H;
constexpr int REGION_H = 8 * BLK_H;
static_assert(BLK_H >= 1 && TILE_W / REGION * (TILE_H / REGION_H) * 32 == TILE_THREADS, "tile geometry");

// One triangle prepared for the phase-A pixel loop (25 words; stride 25 is conflict-free for the
// staging writes, and every read in the pixel loop is a broadcast).  The bbox is kept as floats
// (exact: < 65536) so the loop does no int->float conversions (those run on the slow XU pipe).
struct StagedTri {
    float ecx[3], ecy[3], ek1[3], ek2[3], f[3];
    float da, db, dc;
    float g[3];   // early depth reject: depth_i / f_i (approximate), see TRI_EARLYZ
    float thr[3]; // tame triangles: edge e passes  <=>  value > thr[e]  (0, or -0.5 when the tie rule admits 0)
    float x0, x1, y0, y1;
    uint32_t flags, id, slot;
};
static_assert(sizeof(StagedTri) == 31 * 4, "31 words: odd stride keeps the staging writes conflict-free");

// Early depth reject.  For a covered pixel of a tame triangle with finite non-negative vertex depths,
//   D = e0*da/f0 + e1*db/f1 + e2*dc/f2                       (real arithmetic, all terms >= 0)
// and the reference's float depth d_ref = fl(fl(fl(e0/f0)*da + fl(e1/f1)*db) + fl(e2/f2)*dc) satisfies
// |d_ref - D| <= 4.1 u D (u = 2^-24: one division, one product and two sums of non-negative terms),
// while d_app = fma(e2, g2, fma(e1, g1, e0*g0)) with g_i = fl(d_i * fl(1/f_i)) satisfies
// |d_app - D| <= 5.1 u D.  Hence d_ref >= d_app * (1 - 9.3u) > d_app * (1 - 2^-20), so
//   d_app * (1 - 2^-20) > z   ==>   d_ref > z :  the fragment fails the strict `<` test and is not a tie,
// and its three IEEE divisions can be skipped.  The bounds need normal (not denormal) products, so the
// flag is only set when every g_i is 0 or >= 1e-30; NaN/inf make the comparison false or d_ref = inf.
constexpr uint32_t TRI_EARLYZ = 16u;
constexpr float EARLYZ_SCALE = 0.99999904632568359375f; // 1 - 2^-20

__device__ __forceinline__ void stage_triangle(StagedTri *dst, const RasterRec *src, uint32_t slot) {
    RasterRec r = load_raster(src);
    if (r.id == NO_SLOT) { // empty transparent slot: an empty bbox makes every lane skip it
        r.bbx = 1u;        // x_min = 1 > x_max = 0
        r.bby = 1u;
    }
    const TriEdges t = prepare_edges(r);
    StagedTri &s = *dst;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        s.ecx[i] = t.ecx[i]; s.ecy[i] = t.ecy[i]; s.ek1[i] = t.ek1[i]; s.ek2[i] = t.ek2[i]; s.f[i] = t.f[i];
    }
    s.da = r.da; s.db = r.db; s.dc = r.dc;
    const float dep[3] = {r.da, r.db, r.dc};
    bool earlyz = !(t.flags & TRI_SLOW);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        s.thr[i] = (t.flags & (1u << i)) ? -0.5f : 0.0f; // edge values are integers: e >= 0 <=> e > -0.5
        const float g = dep[i] * __frcp_rn(t.f[i]);
        s.g[i] = g;
        earlyz = earlyz && dep[i] >= 0.0f && dep[i] < 3.0e38f && (g == 0.0f || g >= 1e-30f);
    }
    s.x0 = (float)(r.bbx & 0xFFFF); s.x1 = (float)(r.bbx >> 16);
    s.y0 = (float)(r.bby & 0xFFFF); s.y1 = (float)(r.bby >> 16);
    s.flags = t.flags | (earlyz ? TRI_EARLYZ : 0u);
    s.id = r.id;
    s.slot = slot;
}

// Coverage + depth of one pixel (canvas.rs:673-682).  Tame triangles: sign tests on the
// sign-normalised edge values, divisions only for covered pixels.  Others: the reference's literal
// divide-then-compare.  The quotients e/f are the same either way.
__device__ __forceinline__ bool cover_pixel(const float (&ecx)[3], const float (&ecy)[3], const float (&ek1)[3],
                                            const float (&ek2)[3], const float (&f)[3], uint32_t flags, float da,
                                            float db, float dc, float x, float y, float &depth) {
    float e[3];
#pragma unroll
    for (int i = 0; i < 3; i++) e[i] = FSUB(FADD(FADD(FMUL(ecx[i], x), FMUL(ecy[i], y)), ek1[i]), ek2[i]);
    float bary[3];
    if (!(flags & TRI_SLOW)) {
        const bool in = (e[0] > 0.0f || (e[0] == 0.0f && (flags & 1u))) && (e[1] > 0.0f || (e[1] == 0.0f && (flags & 2u))) &&
                        (e[2] > 0.0f || (e[2] == 0.0f && (flags & 4u)));
        if (!in) return false;
#pragma unroll
        for (int i = 0; i < 3; i++) bary[i] = FDIV(e[i], f[i]);
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) bary[i] = FDIV(e[i], f[i]);
        if (!(bary[0] >= 0.0f && bary[1] >= 0.0f && bary[2] >= 0.0f)) return false;
        if (!((bary[0] > 0.0f || (flags & 1u)) && (bary[1] > 0.0f || (flags & 2u)) && (bary[2] > 0.0f || (flags & 4u))))
            return false;
    }
    depth = FADD(FADD(FMUL(bary[0], da), FMUL(bary[1], db)), FMUL(bary[2], dc)); // canvas.rs:682
    return true;
}

// Order-preserving map float -> uint32 (-0 is folded onto +0: the reference's `<` treats them as
// equal, so the earlier draw must win between them).
__device__ __forceinline__ uint32_t depth_key(float d) {
    const uint32_t b = __float_as_uint(FADD(d, 0.0f));
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ unsigned long long make_key(float d, uint32_t slot) {
    return ((unsigned long long)depth_key(d) << 32) | slot;
}

__global__ void __launch_bounds__(TILE_THREADS) k_tile(const __grid_constant__ FrameUniforms U, const SceneDev S,
                                                       const FrameDev W, uint8_t *__restrict__ color,
                                                       float *__restrict__ depth) {
    __shared__ unsigned long long keys[TILE_PIXELS]; // (depth key, slot), later (depth bits, draw id)
    __shared__ uint32_t colour[TILE_PIXELS];         // r | g << 8 | b << 16 | pad << 24
    __shared__ StagedTri staged[CHUNK];
    __shared__ float u8tab[256]; // (u8 as f32) / 255.0
    pdl_prologue();

    const uint32_t tile = W.tile_order[blockIdx.x]; // heaviest tiles first (k_alloc)
    const uint32_t tile_x = tile % U.tiles_x, tile_y = tile / U.tiles_x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long t_start = W.tile_cycles ? clock64() : 0;
    const int tx0 = (int)tile_x * TILE_W, ty0 = (int)tile_y * TILE_H;

    // (loads issued together; masked afterwards so that they do not wait for the overflow flag)
    const uint32_t overflow = W.counters[2];
    const uint32_t l_begin = W.list_offset[tile], m_begin = W.list_offset[U.n_coarse + tile],
                   s_begin = W.list_offset[2 * U.n_coarse + tile];
    uint32_t l_count = W.list_count[tile], m_count = W.list_count[U.n_coarse + tile],
             s_count = W.list_count[2 * U.n_coarse + tile]; // the fill cursors end at the counts
    const bool usable = overflow == 0;
    if (!usable) l_count = m_count = s_count = 0;
    const RasterRec *__restrict__ rrec = W.rrec;
    const float depth_max = U.depth_max;

    // ---- empty tile: nothing to rasterise, write the clear colour and depth (canvas.rs:425-433) ----
    if (l_count + m_count + s_count == 0 && (S.n_transparent == 0 || !usable)) {
        const int W_ = (int)U.canvas_w, H_ = (int)U.canvas_h;
        const uint32_t clear_px = 255u | (186u << 8) | (155u << 16) | (255u << 24); // memory order b g r pad
        if ((W_ & 3) == 0) {
            for (int q = tid; q < TILE_PIXELS / 4; q += TILE_THREADS) { // 4 pixels (16 B) per store
                const int x = tx0 + (q & (TILE_W / 4 - 1)) * 4, y = ty0 + q / (TILE_W / 4);
                if (x >= W_ || y >= H_) continue;
                *reinterpret_cast<uint4 *>(color + ((size_t)(H_ - 1 - y) * W_ + x) * 4) =
                    make_uint4(clear_px, clear_px, clear_px, clear_px);
                *reinterpret_cast<float4 *>(depth + (size_t)y * W_ + x) = make_float4(depth_max, depth_max, depth_max, depth_max);
            }
        } else {
            for (int p = tid; p < TILE_PIXELS; p += TILE_THREADS) {
                const int x = tx0 + (p & (TILE_W - 1)), y = ty0 + p / TILE_W;
                if (x >= W_ || y >= H_) continue;
                reinterpret_cast<uint32_t *>(color)[(size_t)(H_ - 1 - y) * W_ + x] = clear_px;
                depth[(size_t)y * W_ + x] = depth_max;
            }
        }
        if (W.tile_cycles && tid == 0) W.tile_cycles[tile] = (uint32_t)(clock64() - t_start);
        return;
    }
    fill_u8_table(u8tab, tid, TILE_THREADS); // visible after the barrier that ends phase A

    // ---- phase A: large triangles, every lane tests its own 4x2 block ------------------------------
    {
        // warp -> 16x16 region, lane -> 4x2 block (canvas coordinates: x right, y = depth-buffer row)
        constexpr int WARPS_X = TILE_W / REGION;
        const int rx0 = tx0 + (warp % WARPS_X) * REGION, ry0 = ty0 + (warp / WARPS_X) * REGION_H;
        const int bx0 = rx0 + (lane & 3) * 4, by0 = ry0 + (lane >> 2) * BLK_H;
        const float fx0 = (float)rx0, fy0 = (float)ry0, fx1 = fx0 + (float)(REGION - 1), fy1 = fy0 + (float)(REGION_H - 1);
        float xf[4], yf[BLK_H];
#pragma unroll
        for (int i = 0; i < 4; i++) xf[i] = (float)(bx0 + i);
#pragma unroll
        for (int j = 0; j < BLK_H; j++) yf[j] = (float)(by0 + j);
        float zb[PX];
        uint32_t sl[PX];
#pragma unroll
        for (int i = 0; i < PX; i++) {
            zb[i] = depth_max;
            sl[i] = NO_SLOT;
        }
#pragma unroll 1
        for (uint32_t base = 0; base < l_count; base += CHUNK) {
            const uint32_t n = min((uint32_t)CHUNK, l_count - base);
            __syncthreads();
            if ((uint32_t)tid < n) {
                const uint32_t slot = W.list_refs[l_begin + base + tid];
                stage_triangle(staged + tid, rrec + slot, slot);
            }
            __syncthreads();
#pragma unroll 1
            for (uint32_t k = 0; k < n; k++) {
                const StagedTri &s = staged[k];
                if (s.x1 < fx0 || s.x0 > fx1 || s.y1 < fy0 || s.y0 > fy1) continue; // warp-uniform
                const float lo_x = fmaxf(s.x0, xf[0]), hi_x = fminf(s.x1, xf[3]);
                const float lo_y = fmaxf(s.y0, yf[0]), hi_y = fminf(s.y1, yf[BLK_H - 1]);
                if (lo_x > hi_x || lo_y > hi_y) continue;
                const uint32_t flags = s.flags, slot = s.slot;
                if (!(flags & TRI_SLOW)) {
                    // block-level reject at the best corner of the clipped block (exact, see rect_may_cover)
                    bool any = true;
#pragma unroll
                    for (int e = 0; e < 3; e++) {
                        const float cx = s.ecx[e], cy = s.ecy[e];
                        const float xm = cx >= 0.0f ? hi_x : lo_x, ym = cy >= 0.0f ? hi_y : lo_y;
                        const float em = FSUB(FADD(FADD(FMUL(cx, xm), FMUL(cy, ym)), s.ek1[e]), s.ek2[e]);
                        any = any && em > s.thr[e];
                    }
                    if (!any) continue;
                    float pxs[3][4], pys[3][BLK_H];
#pragma unroll
                    for (int e = 0; e < 3; e++) {
#pragma unroll
                        for (int i = 0; i < 4; i++) pxs[e][i] = FMUL(s.ecx[e], xf[i]);
#pragma unroll
                        for (int j = 0; j < BLK_H; j++) pys[e][j] = FMUL(s.ecy[e], yf[j]);
                    }
#pragma unroll
                    for (int j = 0; j < BLK_H; j++) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            if (xf[i] < lo_x || xf[i] > hi_x || yf[j] < lo_y || yf[j] > hi_y) continue;
                            // f > 0 after sign normalisation: alpha >= 0 <=> e >= 0, alpha > 0 <=> e > 0
                            const float e0 = FSUB(FADD(FADD(pxs[0][i], pys[0][j]), s.ek1[0]), s.ek2[0]);
                            const float e1 = FSUB(FADD(FADD(pxs[1][i], pys[1][j]), s.ek1[1]), s.ek2[1]);
                            const float e2 = FSUB(FADD(FADD(pxs[2][i], pys[2][j]), s.ek1[2]), s.ek2[2]);
                            if (!(e0 > s.thr[0] && e1 > s.thr[1] && e2 > s.thr[2])) continue;
                            const int p = j * 4 + i;
                            // early depth reject (exactly conservative, see TRI_EARLYZ): skip the divisions
                            if ((flags & TRI_EARLYZ) &&
                                fmaf(e2, s.g[2], fmaf(e1, s.g[1], e0 * s.g[0])) * EARLYZ_SCALE > zb[p])
                                continue;
                            const float alpha = FDIV(e0, s.f[0]), beta = FDIV(e1, s.f[1]), gama = FDIV(e2, s.f[2]);
                            const float d = FADD(FADD(FMUL(alpha, s.da), FMUL(beta, s.db)), FMUL(gama, s.dc)); // canvas.rs:682
                            // strict `<` (canvas.rs:923); on equal depth the earlier draw (smaller slot) stays
                            if (d < zb[p] || (d == zb[p] && slot < sl[p] && sl[p] != NO_SLOT)) {
                                zb[p] = d;
                                sl[p] = slot;
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < PX; p++) {
                        const float x = xf[p & 3], y = yf[p >> 2];
                        if (x < lo_x || x > hi_x || y < lo_y || y > hi_y) continue;
                        float d;
                        if (!cover_pixel(s.ecx, s.ecy, s.ek1, s.ek2, s.f, flags, s.da, s.db, s.dc, x, y, d)) continue;
                        if (d < zb[p] || (d == zb[p] && slot < sl[p] && sl[p] != NO_SLOT)) {
                            zb[p] = d;
                            sl[p] = slot;
                        }
                    }
                }
            }
        }
        // ---- merge: publish the block as keys --------------------------------------------------------
#pragma unroll
        for (int p = 0; p < PX; p++) {
            const int lxp = bx0 - tx0 + (p & 3), lyp = by0 - ty0 + (p >> 2);
            keys[lyp * TILE_W + lxp] = sl[p] != NO_SLOT ? make_key(zb[p], sl[p]) : make_key(depth_max, NO_SLOT);
        }
    }
    __syncthreads();

    if (W.tile_cycles && tid == 0) W.tile_cycles[U.n_coarse + tile] = (uint32_t)(clock64() - t_start); // end of phase A
    // ---- phase B1: medium triangles -------------------------------------------------------------------
    // Per chunk of 64 triangles: (1) one thread per triangle stages it and counts the 8x4-pixel blocks of
    // its bbox inside the tile; (2) coarse raster: one thread per (triangle, block) runs the exact block
    // test and queues the blocks that can be covered; (3) fine raster: warps take queued blocks, one
    // pixel per lane, and commit covered fragments with atomicMin on the key.
    {
        __shared__ uint16_t queue[CHUNK * (TILE_W / 8) * (TILE_H / 4)]; // item = tri | bx << 6 | by << 9
        __shared__ uint32_t blk_prefix[CHUNK + 1];
        __shared__ uint32_t q_count;
        const float tx0f = (float)tx0, ty0f = (float)ty0, tx1f = (float)(tx0 + TILE_W - 1), ty1f = (float)(ty0 + TILE_H - 1);
#pragma unroll 1
        for (uint32_t base = 0; base < m_count; base += CHUNK) {
            const uint32_t n = min((uint32_t)CHUNK, m_count - base);
            __syncthreads();
            if ((uint32_t)tid < n) {
                const uint32_t slot = W.list_refs[m_begin + base + tid];
                stage_triangle(staged + tid, rrec + slot, slot);
                const StagedTri &t = staged[tid];
                const float w = FSUB(fminf(t.x1, tx1f), fmaxf(t.x0, tx0f)), h = FSUB(fminf(t.y1, ty1f), fmaxf(t.y0, ty0f));
                blk_prefix[tid + 1] = ((uint32_t)w / 8u + 1u) * ((uint32_t)h / 4u + 1u); // blocks of 8x4 from the bbox corner
            }
            if (tid == 0) {
                blk_prefix[0] = 0;
                q_count = 0;
            }
            __syncthreads();
            if (warp == 0) { // inclusive scan of the (up to 64) block counts
                uint32_t a = lane < n ? blk_prefix[lane + 1] : 0u, b = lane + 32 < n ? blk_prefix[lane + 33] : 0u;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t ua = __shfl_up_sync(0xFFFFFFFFu, a, d), ub = __shfl_up_sync(0xFFFFFFFFu, b, d);
                    if (lane >= d) { a += ua; b += ub; }
                }
                b += __shfl_sync(0xFFFFFFFFu, a, 31);
                if (lane < n) blk_prefix[lane + 1] = a;
                if (lane + 32 < n) blk_prefix[lane + 33] = b;
            }
            __syncthreads();
            // (2) coarse raster
            const uint32_t total = blk_prefix[n];
#pragma unroll 1
            for (uint32_t pbase = 0; pbase < total; pbase += TILE_THREADS) {
                const uint32_t p = pbase + tid;
                bool hit = false;
                uint32_t item = 0;
                if (p < total) {
                    uint32_t lo = 0, hi = n; // largest k with blk_prefix[k] <= p
                    while (hi - lo > 1) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (blk_prefix[mid] <= p) lo = mid; else hi = mid;
                    }
                    const StagedTri &t = staged[lo];
                    const float lx = fmaxf(t.x0, tx0f), hx = fminf(t.x1, tx1f), ly = fmaxf(t.y0, ty0f), hy = fminf(t.y1, ty1f);
                    const uint32_t nbx = (uint32_t)FSUB(hx, lx) / 8u + 1u, local = p - blk_prefix[lo];
                    const uint32_t bxi = local % nbx, byi = local / nbx;
                    const float bx = FADD(lx, (float)(bxi * 8u)), by = FADD(ly, (float)(byi * 4u));
                    hit = rect_may_cover(t, bx, fminf(FADD(bx, 7.0f), hx), by, fminf(FADD(by, 3.0f), hy));
                    item = lo | (bxi << 6) | (byi << 9);
                }
                const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
                uint32_t wbase = 0;
                if (lane == 0 && ballot) wbase = atomicAdd(&q_count, (uint32_t)__popc(ballot));
                wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
                if (hit) queue[wbase + __popc(ballot & ((1u << lane) - 1u))] = (uint16_t)item;
            }
            __syncthreads();
            // (3) fine raster
            const uint32_t nq = q_count;
            const float dxf = (float)(lane & 7), dyf = (float)(lane >> 3);
#pragma unroll 1
            for (uint32_t j = (uint32_t)warp; j < nq; j += TILE_THREADS / 32) {
                const uint32_t item = queue[j];
                const StagedTri &t = staged[item & 63u];
                const float lx = fmaxf(t.x0, tx0f), hx = fminf(t.x1, tx1f), ly = fmaxf(t.y0, ty0f), hy = fminf(t.y1, ty1f);
                const float x = FADD(FADD(lx, (float)(((item >> 6) & 7u) * 8u)), dxf);
                const float y = FADD(FADD(ly, (float)((item >> 9) * 4u)), dyf);
                if (x > hx || y > hy) continue;
                unsigned long long *cell = &keys[(int)FSUB(y, ty0f) * TILE_W + (int)FSUB(x, tx0f)];
                float d;
                if (!(t.flags & TRI_SLOW)) {
                    const float e0 = FSUB(FADD(FADD(FMUL(t.ecx[0], x), FMUL(t.ecy[0], y)), t.ek1[0]), t.ek2[0]);
                    const float e1 = FSUB(FADD(FADD(FMUL(t.ecx[1], x), FMUL(t.ecy[1], y)), t.ek1[1]), t.ek2[1]);
                    const float e2 = FSUB(FADD(FADD(FMUL(t.ecx[2], x), FMUL(t.ecy[2], y)), t.ek1[2]), t.ek2[2]);
                    if (!(e0 > t.thr[0] && e1 > t.thr[1] && e2 > t.thr[2])) continue;
                    // early depth reject in key space (exactly conservative, see TRI_EARLYZ)
                    if ((t.flags & TRI_EARLYZ) &&
                        depth_key(fmaf(e2, t.g[2], fmaf(e1, t.g[1], e0 * t.g[0])) * EARLYZ_SCALE) > (uint32_t)(*cell >> 32))
                        continue;
                    const float alpha = FDIV(e0, t.f[0]), beta = FDIV(e1, t.f[1]), gama = FDIV(e2, t.f[2]);
                    d = FADD(FADD(FMUL(alpha, t.da), FMUL(beta, t.db)), FMUL(gama, t.dc)); // canvas.rs:682
                } else if (!cover_pixel(t.ecx, t.ecy, t.ek1, t.ek2, t.f, t.flags, t.da, t.db, t.dc, x, y, d)) {
                    continue;
                }
                if (!(d < depth_max)) continue;
                const unsigned long long key = make_key(d, t.slot);
                if (key < *cell) atomicMin(cell, key);
            }
        }
    }
    // ---- phase B2: small triangles, one per lane, atomicMin on the key --------------------------
    // (item j of a round goes to lane j / warps of warp j % warps, so a short list spreads over all warps)
    for (uint32_t i = (uint32_t)(lane * (TILE_THREADS / 32) + warp); i < s_count; i += TILE_THREADS) {
        const uint32_t slot = W.list_refs[s_begin + i];
        const RasterRec r = load_raster(rrec + slot);
        const TriEdges t = prepare_edges(r);
        const int lx = max((int)(r.bbx & 0xFFFF), tx0), hx = min((int)(r.bbx >> 16), tx0 + TILE_W - 1);
        const int ly = max((int)(r.bby & 0xFFFF), ty0), hy = min((int)(r.bby >> 16), ty0 + TILE_H - 1);
        float y = (float)ly;
        for (int yi = ly; yi <= hy; yi++, y = FADD(y, 1.0f)) {
            float x = (float)lx;
            for (int xi = lx; xi <= hx; xi++, x = FADD(x, 1.0f)) {
                float d;
                if (!cover_pixel(t.ecx, t.ecy, t.ek1, t.ek2, t.f, t.flags, r.da, r.db, r.dc, x, y, d)) continue;
                if (!(d < depth_max)) continue; // also rejects NaN; equality with the clear depth fails `<`
                const unsigned long long key = make_key(d, slot);
                unsigned long long *cell = &keys[(yi - ty0) * TILE_W + (xi - tx0)];
                if (key < *cell) atomicMin(cell, key);
            }
        }
    }
    __syncthreads();

    if (W.tile_cycles && tid == 0) W.tile_cycles[2 * U.n_coarse + tile] = (uint32_t)(clock64() - t_start); // end of phase B
    // ---- phase C: deferred shading, fused clear ----------------------------------------------------
#pragma unroll 1
    for (int it = 0; it < TILE_PIXELS / TILE_THREADS; it++) {
        const int p = it * TILE_THREADS + tid;
        const uint32_t slot = (uint32_t)keys[p];
        uint32_t c = 155u | (186u << 8) | (255u << 16) | (255u << 24); // azul_bb, pad 255 (canvas.rs:131)
        float d = depth_max;
        uint32_t id = NO_SLOT;
        if (slot != NO_SLOT) {
            const RasterRec r = load_raster(rrec + slot);
            float op;
            c = shade_pixel(S.materials, S.texels, u8tab, r, W.srec + slot, (float)(tx0 + (p & (TILE_W - 1))),
                            (float)(ty0 + p / TILE_W), &d, &op) | (255u << 24);
            id = r.id;
        }
        colour[p] = c;
        keys[p] = ((unsigned long long)__float_as_uint(d) << 32) | id; // own pixel: no sync needed
    }

    // ---- phase D: transparent triangles in draw order (scene/mod.rs:1088-1246) -------------------
    const uint32_t n_tslots = usable ? S.n_transparent * 4u : 0u;
#pragma unroll 1
    for (uint32_t base = 0; base < n_tslots; base += CHUNK) {
        const uint32_t n = min((uint32_t)CHUNK, n_tslots - base);
        __syncthreads();
        if ((uint32_t)tid < n) stage_triangle(staged + tid, W.t_rrec + base + tid, base + tid);
        __syncthreads();
#pragma unroll 1
        for (uint32_t k = 0; k < n; k++) {
            const StagedTri &s = staged[k];
            if (s.x1 < (float)tx0 || s.x0 > (float)(tx0 + TILE_W - 1) || s.y1 < (float)ty0 || s.y0 > (float)(ty0 + TILE_H - 1))
                continue;
#pragma unroll 1
            for (int it = 0; it < TILE_PIXELS / TILE_THREADS; it++) {
                const int p = it * TILE_THREADS + tid;
                const float x = (float)(tx0 + (p & (TILE_W - 1))), y = (float)(ty0 + p / TILE_W);
                if (x < s.x0 || x > s.x1 || y < s.y0 || y > s.y1) continue;
                float d;
                if (!cover_pixel(s.ecx, s.ecy, s.ek1, s.ek2, s.f, s.flags | TRI_SLOW, s.da, s.db, s.dc, x, y, d)) continue;
                const unsigned long long key = keys[p];
                const uint32_t wid = (uint32_t)key;
                if (!(wid == NO_SLOT || s.id > wid)) continue;            // drawn before the opaque winner: overwritten
                if (!(d < __uint_as_float((uint32_t)(key >> 32)))) continue; // canvas.rs:923, depth write is off
                const RasterRec r = load_raster(W.t_rrec + s.slot);
                float d2, op;
                const uint32_t rgb = shade_pixel(S.materials, S.texels, u8tab, r, W.t_srec + s.slot, x, y, &d2, &op);
                // canvas.rs:916-921: opacity < 1 blends with the stored colour, else replaces it
                colour[p] = op < 1.0f ? blend_rgb(colour[p], rgb, op) : (rgb | (255u << 24));
            }
        }
    }

    // ---- phase E: single write-back -------------------------------------------------------------
    const int W_ = (int)U.canvas_w, H_ = (int)U.canvas_h;
#pragma unroll
    for (int it = 0; it < TILE_PIXELS / TILE_THREADS; it++) {
        const int p = it * TILE_THREADS + tid;
        const int x = tx0 + (p & (TILE_W - 1)), y = ty0 + p / TILE_W;
        if (x >= W_ || y >= H_) continue;
        const uint32_t c = colour[p]; // r g b pad -> memory order b g r pad
        reinterpret_cast<uint32_t *>(color)[(size_t)(H_ - 1 - y) * W_ + x] =
            ((c >> 16) & 255u) | (c & 0x0000FF00u) | ((c & 255u) << 16) | (c & 0xFF000000u);
        depth[(size_t)y * W_ + x] = __uint_as_float((uint32_t)(keys[p] >> 32));
    }
    if (W.tile_cycles) {
        __syncthreads();
        if (tid == 0) W.tile_cycles[tile] = (uint32_t)(clock64() - t_start);
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
// launchers
// ------------------------------------------------------------------------------------------
void launch_tile(const FrameUniforms &U, const SceneDev &S, const FrameDev &W, uint8_t *color, float *depth,
                 cudaStream_t stream) {
    const uint32_t stripe_tiles = (U.tile_y_end - U.tile_y_begin) * U.tiles_x;
    if (stripe_tiles) launch_pdl(k_tile, stripe_tiles, TILE_THREADS, stream, U, S, W, color, depth);
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
