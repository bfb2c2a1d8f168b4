// k_tile.cu — per-tile raster / depth / shade kernel (mororo18/draw canvas.rs:577-750, 906-960).
//
// One CTA (TILE_THREADS = 256 threads) per 64x32-pixel tile or pixel window of one (persistent CTAs, k_alloc's work list).  The tile's depth, winning record and colour stay on
// chip (registers, then shared memory) for the whole kernel; colour and depth go to HBM exactly
// once at the end, with the clear fused in.  No tensor cores: nothing here is a contraction.
//
//   phase A  "large" list: triangles are staged through shared memory 64 at a time; every lane owns
//            a 4 x BLK_H pixel block (warp = REGION x REGION_H region) and tests it against each triangle, after a
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

constexpr int CHUNK = 64;      // triangles staged per round
constexpr int CAND_CAP = 1024; // window filter: list references examined per segment
constexpr int TILE_PIXELS = TILE_W * TILE_H;
// phase A geometry (device_types.h): a warp owns a REGION x REGION_H rectangle (4 lanes across, 8 down),
// a lane a 4 x BLK_H block of it
constexpr int PX = 4 * BLK_H;
static_assert(BLK_H >= 1 && TILE_W / REGION * (TILE_H / REGION_H) * 32 == TILE_THREADS, "tile geometry");

// A staged triangle is the first PREP_WORDS words of its PrepRec (device_types.h) in shared memory, at an
// odd stride so that threads reading the same field of different triangles hit different banks; reads in
// the phase-A pixel loop are broadcasts.  Word offsets:
enum : int { S_ECX = 0, S_ECY = 3, S_EK1 = 6, S_EK2 = 9, S_F = 12, S_RF = 15, S_DA = 18, S_DB = 19, S_X0 = 20, S_X1 = 21,
             S_Y0 = 22, S_Y1 = 23, S_DC = 24, S_FLAGS = 25, S_ID = 26, S_SLOT = 27 };
constexpr int STAGE_STRIDE = PREP_WORDS + 1;
static_assert(PREP_WORDS == 28 && (STAGE_STRIDE & 1) == 1 && offsetof(PrepRec, x0) == 4 * S_X0 &&
              offsetof(PrepRec, flags) == 4 * S_FLAGS && offsetof(PrepRec, slot) == 4 * S_SLOT, "PrepRec layout");

// Copies n prepared records into shared memory: one 16-byte load per thread (8 threads per record, the
// eighth idle), so a chunk of 64 is two loads per thread of a 256-thread CTA.
__device__ __forceinline__ void stage_copy(float (*staged)[STAGE_STRIDE], const PrepRec *__restrict__ prep,
                                           const uint32_t *refs, uint32_t ref_stride, uint32_t n, int tid) {
    for (uint32_t i = (uint32_t)tid; i < n * 8u; i += TILE_THREADS) {
        const uint32_t k = i >> 3, q = i & 7u;
        if (q == 7u) continue;
        const uint32_t slot = refs[k * ref_stride];
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(prep + slot) + q);
        float *d = staged[k] + 4 * q;
        d[0] = __uint_as_float(v.x); d[1] = __uint_as_float(v.y); d[2] = __uint_as_float(v.z);
        d[3] = q == 6u ? __uint_as_float(slot) : __uint_as_float(v.w);
    }
}
// Phase D only (transparent records are not binned and have no PrepRec): prepare in place.
__device__ __forceinline__ void stage_triangle(float *dst, const RasterRec *src, uint32_t slot) {
    RasterRec r = load_raster(src);
    if (r.id == NO_SLOT) { // empty transparent slot: an empty bbox makes every lane skip it
        r.bbx = 1u;        // x_min = 1 > x_max = 0
        r.bby = 1u;
    }
    PrepRec p;
    make_prep(r, p);
    p.slot = slot;
    const float *w = reinterpret_cast<const float *>(&p);
#pragma unroll
    for (int i = 0; i < PREP_WORDS; i++) dst[i] = w[i];
}

// Window filter: keeps the references whose bbox meets the window (only their bbox quad is read).
// Returns the number kept; cand[] is valid after the call (ends with a barrier).
__device__ __forceinline__ uint32_t filter_refs(const uint32_t *__restrict__ list, uint32_t list_stride, uint32_t count,
                                                const PrepRec *__restrict__ prep, uint32_t *cand, uint32_t *s_count,
                                                float wx0f, float wx1f, float wy0f, float wy1f, int tid) {
    __syncthreads(); // cand / s_count may still be in use by the previous segment
    if (tid == 0) *s_count = 0;
    __syncthreads();
    const int lane = tid & 31;
    for (uint32_t base = 0; base < count; base += TILE_THREADS) {
        const uint32_t i = base + (uint32_t)tid;
        bool keep = false;
        uint32_t slot = 0;
        if (i < count) {
            slot = list[i * list_stride];
            // medium lists: entries 1..3 of a reference split for k_raster name the same triangle again
            const bool extra_part = list_stride == 2u && ((list[i * 2u + 1u] >> 21) & 3u) != 0u;
            const uint4 bb = __ldg(reinterpret_cast<const uint4 *>(prep + slot) + 5); // x0 x1 y0 y1
            keep = !extra_part && !(__uint_as_float(bb.y) < wx0f || __uint_as_float(bb.x) > wx1f ||
                                    __uint_as_float(bb.w) < wy0f || __uint_as_float(bb.z) > wy1f);
        }
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
        uint32_t wbase = 0;
        if (lane == 0 && ballot) wbase = atomicAdd(s_count, (uint32_t)__popc(ballot));
        wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
        if (keep) cand[wbase + __popc(ballot & ((1u << lane) - 1u))] = slot;
    }
    __syncthreads();
    return *s_count;
}

__device__ __forceinline__ TriRegs tri_from_words(const float *w) {
    TriRegs t;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        t.ecx[i] = w[S_ECX + i]; t.ecy[i] = w[S_ECY + i]; t.ek1[i] = w[S_EK1 + i]; t.ek2[i] = w[S_EK2 + i];
        t.f[i] = w[S_F + i]; t.rf[i] = w[S_RF + i];
    }
    t.da = w[S_DA]; t.db = w[S_DB]; t.dc = w[S_DC];
    t.flags = __float_as_uint(w[S_FLAGS]);
    return t;
}

// Exact block reject on a staged triangle (see rect_may_cover in device_math.cuh).
__device__ __forceinline__ bool staged_may_cover(const float *s, uint32_t flags, float lx, float hx, float ly, float hy) {
    if (flags & TRI_SLOW) return true;
    bool any = true;
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const float cx = s[S_ECX + e], cy = s[S_ECY + e];
        const float xm = cx >= 0.0f ? hx : lx, ym = cy >= 0.0f ? hy : ly;
        const float em = FSUB(FADD(FADD(FMUL(cx, xm), FMUL(cy, ym)), s[S_EK1 + e]), s[S_EK2 + e]);
        any = any && em > ((flags >> e) & 1u ? -0.5f : 0.0f); // edge values are integers: e >= 0 <=> e > -0.5
    }
    return any;
}

// One work item (k_alloc, device_types.h): a tile or one pixel window of a dense tile.
__device__ __forceinline__ void tile_item(const uint32_t item, const FrameUniforms &U, const SceneDev &S, const FrameDev &W,
                                          uint8_t *__restrict__ color, float *__restrict__ depth, const float *u8tab,
                                          const bool usable) {
    __shared__ __align__(16) unsigned long long keys[TILE_PIXELS]; // (depth key, slot), later (depth bits, draw id)
    __shared__ uint32_t colour[TILE_PIXELS];         // r | g << 8 | b << 16 | pad << 24
    __shared__ float staged[CHUNK][STAGE_STRIDE];
    __shared__ uint32_t cand[CAND_CAP];
    __shared__ uint32_t s_cand_count;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W_ = (int)U.canvas_w, H_ = (int)U.canvas_h;
    const float depth_max = U.depth_max;

    const uint32_t tile_x = item & (MAX_TILES_X - 1), tile_y = (item >> 10) & (MAX_TILES_Y - 1);
    const uint32_t tile = tile_y * U.tiles_x + tile_x;
    const bool defer = (item & ITEM_DEFER) != 0; // only add the large triangles to the key page; k_shade does the rest
    const long long t_start = W.tile_cycles ? clock64() : 0;
    const int tx0 = (int)tile_x * TILE_W, ty0 = (int)tile_y * TILE_H;

    // pixel window of the tile this CTA renders (the whole tile unless k_alloc cut the tile up)
    int wx0 = tx0, wy0 = ty0, ww = TILE_W, wh = TILE_H;
    if (TILE_SPLITTABLE) {
        wx0 = tx0 + (int)((item >> 21) & 3u) * REGION;
        wy0 = ty0 + (int)((item >> 23) & 3u) * REGION_H;
        ww = (int)(((item >> 25) & 3u) + 1u) * REGION;
        wh = (int)(((item >> 27) & 3u) + 1u) * REGION_H;
    }
    const bool windowed = ww != TILE_W || wh != TILE_H; // then the lists hold triangles that miss the window
    const int ww_shift = 31 - __clz(ww), n_win = ww * wh; // ww is a power of two
    const float wx0f = (float)wx0, wy0f = (float)wy0, wx1f = (float)(wx0 + ww - 1), wy1f = (float)(wy0 + wh - 1);
    const float tx0f = (float)tx0, ty0f = (float)ty0;

    const uint32_t l_begin = W.list_offset[tile], m_begin = W.list_offset[U.n_coarse + tile],
                   s_begin = W.list_offset[2 * U.n_coarse + tile];
    uint32_t l_count = W.list_count[tile], m_count = W.list_count[U.n_coarse + tile],
             s_count = W.list_count[2 * U.n_coarse + tile]; // the fill cursors end at the counts
    const uint32_t page = W.tile_page[tile];
    if (!usable) l_count = m_count = s_count = 0;
    // the tile has a key page: k_raster has already rasterised its medium and small lists into it
    if (page != NO_PAGE) m_count = s_count = 0;
    const RasterRec *__restrict__ rrec = W.rrec;
    const PrepRec *__restrict__ prep = W.prep;

    // ---- phase A: large triangles, every lane tests its own 4 x BLK_H block -----------------------
    {
        // warp -> REGION x REGION_H region, lane -> 4 x BLK_H block (canvas coordinates: x right, y = depth-buffer row)
        constexpr int WARPS_X = TILE_W / REGION;
        const int rx0 = tx0 + (warp % WARPS_X) * REGION, ry0 = ty0 + (warp / WARPS_X) * REGION_H;
        const int bx0 = rx0 + (lane & 3) * 4, by0 = ry0 + (lane >> 2) * BLK_H;
        const bool warp_in = rx0 >= wx0 && rx0 < wx0 + ww && ry0 >= wy0 && ry0 < wy0 + wh; // regions tile the window
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
        // The tile's key page (k_raster's result): its loads are issued here, together with the stores that
        // leave it empty for the next frame, and consumed after the first chunk of large triangles has been
        // staged, so that the two L2 round trips overlap.  A lane's 4 pixels of a row are 32 contiguous bytes.
        const bool have_page = page != NO_PAGE && usable && warp_in;
        ulonglong2 pk01[BLK_H], pk23[BLK_H];
        if (have_page) {
            unsigned long long *pk = W.key_pages + (size_t)page * TILE_PIXELS;
#pragma unroll
            for (int j = 0; j < BLK_H; j++) {
                ulonglong2 *src = reinterpret_cast<ulonglong2 *>(pk + (by0 - ty0 + j) * TILE_W + (bx0 - tx0));
#if DRAW_PAGE_LD == 1
                pk01[j] = *src;
                pk23[j] = *(src + 1);
#elif DRAW_PAGE_LD == 2
                pk01[j] = *reinterpret_cast<volatile ulonglong2 *>(src);
                pk23[j] = *reinterpret_cast<volatile ulonglong2 *>(src + 1);
#else
                pk01[j] = __ldcg(src);
                pk23[j] = __ldcg(src + 1);
#endif

            }
        }
        bool page_pending = have_page;
        auto merge_page = [&]() { // the page's fragments become the starting depth / winner of the lane's pixels
#pragma unroll
            for (int j = 0; j < BLK_H; j++) {
                const unsigned long long kk[4] = {pk01[j].x, pk01[j].y, pk23[j].x, pk23[j].y};
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (kk[i] != KEY_EMPTY) {
                        zb[j * 4 + i] = depth_from_key((uint32_t)(kk[i] >> 32));
                        sl[j * 4 + i] = (uint32_t)kk[i];
                    }
            }
            page_pending = false;
            if (defer) return; // the page is rewritten below with the merged keys and emptied by k_shade
            // Leave the page empty for the next frame: after every lane of the warp has its keys (the loads
            // above are consumed), the warp's region — REGION_H rows of 128 bytes — is overwritten with whole
            // 128-byte lines, 8 lanes per row.
            __syncwarp();
            unsigned long long *pk = W.key_pages + (size_t)page * TILE_PIXELS;
            constexpr int ROW_QUADS = REGION * 8 / 16; // 16-byte stores per region row
#pragma unroll
            for (int r = lane / ROW_QUADS; r < REGION_H; r += 32 / ROW_QUADS)
                __stcg(reinterpret_cast<ulonglong2 *>(pk + (ry0 - ty0 + r) * TILE_W + (rx0 - tx0)) + lane % ROW_QUADS,
                       make_ulonglong2(KEY_EMPTY, KEY_EMPTY));
        };
#pragma unroll 1
        for (uint32_t seg = 0; seg < l_count; seg += CAND_CAP) {
            const uint32_t seg_n = min((uint32_t)CAND_CAP, l_count - seg);
            const uint32_t *refs = W.list_refs + l_begin + seg;
            uint32_t n_refs = seg_n;
            if (windowed) {
                n_refs = filter_refs(refs, 1u, seg_n, prep, cand, &s_cand_count, wx0f, wx1f, wy0f, wy1f, tid);
                refs = cand;
            }
#pragma unroll 1
            for (uint32_t base = 0; base < n_refs; base += CHUNK) {
                const uint32_t n = min((uint32_t)CHUNK, n_refs - base);
                __syncthreads();
                stage_copy(staged, prep, refs + base, 1u, n, tid);
                __syncthreads();
#if DRAW_TAP_B == 3
                if (W.tile_cycles && tid == 0) atomicMax(&W.tile_cycles[U.n_coarse + tile], (uint32_t)(clock64() - t_start)); // staged
#endif
                if (page_pending) merge_page();
#if DRAW_TAP_B == 3
                if (W.tile_cycles && tid == 0) atomicMax(&W.tile_cycles[2 * U.n_coarse + tile], (uint32_t)(clock64() - t_start)); // page merged
#endif
#pragma unroll 1
                for (uint32_t k = 0; k < (warp_in ? n : 0u); k++) {
                    const float *s = staged[k];
                    const float sx0 = s[S_X0], sx1 = s[S_X1], sy0 = s[S_Y0], sy1 = s[S_Y1];
                    if (sx1 < fx0 || sx0 > fx1 || sy1 < fy0 || sy0 > fy1) continue; // warp-uniform
                    const float lo_x = fmaxf(sx0, xf[0]), hi_x = fminf(sx1, xf[3]);
                    const float lo_y = fmaxf(sy0, yf[0]), hi_y = fminf(sy1, yf[BLK_H - 1]);
                    if (lo_x > hi_x || lo_y > hi_y) continue;
                    const uint32_t flags = __float_as_uint(s[S_FLAGS]), slot = __float_as_uint(s[S_SLOT]);
                    if (!(flags & TRI_SLOW)) {
                        // block-level reject at the best corner of the clipped block (exact, see rect_may_cover)
                        if (!staged_may_cover(s, flags, lo_x, hi_x, lo_y, hi_y)) continue;
                        const float thr0 = (flags & 1u) ? -0.5f : 0.0f, thr1 = (flags & 2u) ? -0.5f : 0.0f,
                                    thr2 = (flags & 4u) ? -0.5f : 0.0f;
                        const float da = s[S_DA], db = s[S_DB], dc = s[S_DC];
                        const float rf0 = s[S_RF], rf1 = s[S_RF + 1], rf2 = s[S_RF + 2];
                        const float g0 = FMUL(da, rf0), g1 = FMUL(db, rf1), g2 = FMUL(dc, rf2);
                        // edge values of the block's pixels (f > 0 after sign normalisation:
                        // alpha >= 0 <=> e >= 0, alpha > 0 <=> e > 0), coverage and early depth reject, branch-free
                        float e0[PX], e1[PX], e2[PX];
                        uint32_t mask = 0;
                        float pys[3][BLK_H];
#pragma unroll
                        for (int e = 0; e < 3; e++)
#pragma unroll
                            for (int j = 0; j < BLK_H; j++) pys[e][j] = FMUL(s[S_ECY + e], yf[j]);
#pragma unroll
                        for (int j = 0; j < BLK_H; j++) {
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                const int p = j * 4 + i;
                                e0[p] = FSUB(FADD(FADD(FMUL(s[S_ECX], xf[i]), pys[0][j]), s[S_EK1]), s[S_EK2]);
                                e1[p] = FSUB(FADD(FADD(FMUL(s[S_ECX + 1], xf[i]), pys[1][j]), s[S_EK1 + 1]), s[S_EK2 + 1]);
                                e2[p] = FSUB(FADD(FADD(FMUL(s[S_ECX + 2], xf[i]), pys[2][j]), s[S_EK1 + 2]), s[S_EK2 + 2]);
                                bool in = xf[i] >= lo_x && xf[i] <= hi_x && yf[j] >= lo_y && yf[j] <= hi_y && e0[p] > thr0 &&
                                          e1[p] > thr1 && e2[p] > thr2;
                                // early depth reject (exactly conservative, see TRI_EARLYZ): skip the divisions
                                if (flags & TRI_EARLYZ)
                                    in = in && !(fmaf(e2[p], g2, fmaf(e1[p], g1, e0[p] * g0)) * EARLYZ_SCALE > zb[p]);
                                mask |= (in ? 1u : 0u) << p;
                            }
                        }
                        if (!mask) continue;
                        if (flags & TRI_FASTDIV) {
                            const float f0 = s[S_F], f1 = s[S_F + 1], f2 = s[S_F + 2];
#pragma unroll
                            for (int p = 0; p < PX; p++) { // all pixels of the block: no branches, the quotients overlap
                                const float alpha = exact_div(e0[p], f0, rf0), beta = exact_div(e1[p], f1, rf1),
                                            gama = exact_div(e2[p], f2, rf2);
                                const float d = FADD(FADD(FMUL(alpha, da), FMUL(beta, db)), FMUL(gama, dc)); // canvas.rs:682
                                // strict `<` (canvas.rs:923); on equal depth the earlier draw (smaller slot) stays
                                const bool take = ((mask >> p) & 1u) && (d < zb[p] || (d == zb[p] && slot < sl[p] && sl[p] != NO_SLOT));
                                zb[p] = take ? d : zb[p];
                                sl[p] = take ? slot : sl[p];
                            }
                        } else {
#pragma unroll
                            for (int p = 0; p < PX; p++) {
                                if (!((mask >> p) & 1u)) continue;
                                const float alpha = FDIV(e0[p], s[S_F]), beta = FDIV(e1[p], s[S_F + 1]), gama = FDIV(e2[p], s[S_F + 2]);
                                const float d = FADD(FADD(FMUL(alpha, da), FMUL(beta, db)), FMUL(gama, dc));
                                if (d < zb[p] || (d == zb[p] && slot < sl[p] && sl[p] != NO_SLOT)) {
                                    zb[p] = d;
                                    sl[p] = slot;
                                }
                            }
                        }
                    } else {
                        const TriRegs t = tri_from_words(s);
#pragma unroll
                        for (int p = 0; p < PX; p++) {
                            const float x = xf[p & 3], y = yf[p >> 2];
                            if (x < lo_x || x > hi_x || y < lo_y || y > hi_y) continue;
                            float d;
                            if (!cover_pixel(t, flags, x, y, d)) continue;
                            if (d < zb[p] || (d == zb[p] && slot < sl[p] && sl[p] != NO_SLOT)) {
                                zb[p] = d;
                                sl[p] = slot;
                            }
                        }
                    }
                }
            }
        }
        if (page_pending) merge_page();
        // ---- merge: publish the block as keys --------------------------------------------------------
#pragma unroll
        for (int p = 0; p < PX; p++) {
            const int lxp = bx0 - tx0 + (p & 3), lyp = by0 - ty0 + (p >> 2);
            keys[lyp * TILE_W + lxp] = sl[p] != NO_SLOT ? make_key(zb[p], sl[p]) : make_key(depth_max, NO_SLOT);
        }
    }
    __syncthreads();

    if (defer) {
        // the window's keys go back to the page, whole rows at a time (16 bytes per lane, consecutive lanes)
        unsigned long long *pk = W.key_pages + (size_t)page * TILE_PIXELS;
        const int row_pairs = ww / 2; // 16-byte key pairs per window row
        for (int q = tid; q < row_pairs * wh; q += TILE_THREADS) {
            const int lx = (wx0 - tx0) + 2 * (q % row_pairs), ly = (wy0 - ty0) + q / row_pairs;
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&keys[ly * TILE_W + lx]);
            __stcg(reinterpret_cast<ulonglong2 *>(pk + ly * TILE_W + lx), v);
        }
        if (W.tile_cycles && tid == 0) atomicMax(&W.tile_cycles[tile], (uint32_t)(clock64() - t_start));
        return;
    }
#ifndef DRAW_TAP_B
    if (W.tile_cycles && tid == 0) atomicMax(&W.tile_cycles[U.n_coarse + tile], (uint32_t)(clock64() - t_start)); // end of phase A
#endif
#if DRAW_TAP_B == 3
    if (W.tile_cycles && tid == 0) atomicMax(&W.tile_cycles[tile], (uint32_t)(clock64() - t_start)); // end of phase A
    if (W.tile_cycles) return;
#endif
    // ---- phase B1: medium triangles -------------------------------------------------------------------
    // Per chunk of 64 triangles: (1) the chunk is staged and one thread per triangle counts the 8x4-pixel
    // blocks of its bbox inside the window; (2) coarse raster: one thread per (triangle, block) runs the
    // exact block test and queues the blocks that can be covered; (3) fine raster: warps take queued
    // blocks, one pixel per lane, and commit covered fragments with atomicMin on the key.
    {
        __shared__ uint16_t queue[CHUNK * (TILE_W / 8) * (TILE_H / 4)]; // item = tri | bx << 6 | by << 9
        __shared__ uint32_t blk_prefix[CHUNK + 1];
        __shared__ uint32_t q_count;
#ifdef DRAW_TAP_B
        long long tap_stage = 0, tap_coarse = 0, tap_fine = 0;
#endif
#if DRAW_TAP_B == 2
        long long tap_filter = 0;
#endif
#pragma unroll 1
        for (uint32_t seg = 0; seg < m_count; seg += CAND_CAP) {
            const uint32_t seg_n = min((uint32_t)CAND_CAP, m_count - seg);
            const uint32_t *refs = reinterpret_cast<const uint32_t *>(W.m_refs + m_begin + seg); // (slot, tile) pairs
            uint32_t n_refs = seg_n, ref_stride = 2u;
            { // always filtered: besides the window test this drops the extra entries of split references
#if DRAW_TAP_B == 2
                const long long tf0 = clock64();
#endif
                n_refs = filter_refs(refs, 2u, seg_n, prep, cand, &s_cand_count, wx0f, wx1f, wy0f, wy1f, tid);
                refs = cand;
                ref_stride = 1u;
#if DRAW_TAP_B == 2
                tap_filter += clock64() - tf0;
#endif
            }
#pragma unroll 1
            for (uint32_t base = 0; base < n_refs; base += CHUNK) {
                const uint32_t n = min((uint32_t)CHUNK, n_refs - base);
                __syncthreads();
#ifdef DRAW_TAP_B
                const long long tb0 = clock64();
#endif
                stage_copy(staged, prep, refs + base * ref_stride, ref_stride, n, tid);
                if (tid == 0) {
                    blk_prefix[0] = 0;
                    q_count = 0;
                }
                __syncthreads();
                if (warp == 0) { // block counts (8x4 blocks from the corner of the bbox clipped to the window) and their inclusive scan
                    uint32_t cnt[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint32_t k = (uint32_t)lane + 32u * h;
                        cnt[h] = 0;
                        if (k < n) {
                            const float *t = staged[k];
                            const float w = FSUB(fminf(t[S_X1], wx1f), fmaxf(t[S_X0], wx0f)), hgt = FSUB(fminf(t[S_Y1], wy1f), fmaxf(t[S_Y0], wy0f));
                            cnt[h] = (w < 0.0f || hgt < 0.0f) ? 0u : ((uint32_t)w / 8u + 1u) * ((uint32_t)hgt / 4u + 1u);
                        }
                    }
                    uint32_t a = cnt[0], b = cnt[1];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t ua = __shfl_up_sync(0xFFFFFFFFu, a, d), ub = __shfl_up_sync(0xFFFFFFFFu, b, d);
                        if (lane >= d) { a += ua; b += ub; }
                    }
                    b += __shfl_sync(0xFFFFFFFFu, a, 31);
                    if ((uint32_t)lane < n) blk_prefix[lane + 1] = a;
                    if ((uint32_t)lane + 32u < n) blk_prefix[lane + 33] = b;
                }
                __syncthreads();
#ifdef DRAW_TAP_B
                const long long tb1 = clock64();
#endif
                // (2) coarse raster
                const uint32_t total = blk_prefix[n];
#pragma unroll 1
                for (uint32_t pbase = 0; pbase < total; pbase += TILE_THREADS) {
                    const uint32_t p = pbase + tid;
                    bool hit = false;
                    uint32_t qitem = 0;
                    if (p < total) {
                        uint32_t lo = 0, hi = n; // largest k with blk_prefix[k] <= p
                        while (hi - lo > 1) {
                            const uint32_t mid = (lo + hi) >> 1;
                            if (blk_prefix[mid] <= p) lo = mid; else hi = mid;
                        }
                        const float *t = staged[lo];
                        const float lx = fmaxf(t[S_X0], wx0f), hx = fminf(t[S_X1], wx1f), ly = fmaxf(t[S_Y0], wy0f), hy = fminf(t[S_Y1], wy1f);
                        const uint32_t nbx = (uint32_t)FSUB(hx, lx) / 8u + 1u, local = p - blk_prefix[lo];
                        const uint32_t bxi = local % nbx, byi = local / nbx;
                        const float bx = FADD(lx, (float)(bxi * 8u)), by = FADD(ly, (float)(byi * 4u));
                        hit = staged_may_cover(t, __float_as_uint(t[S_FLAGS]), bx, fminf(FADD(bx, 7.0f), hx), by, fminf(FADD(by, 3.0f), hy));
                        qitem = lo | (bxi << 6) | (byi << 9);
                    }
                    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
                    uint32_t wbase = 0;
                    if (lane == 0 && ballot) wbase = atomicAdd(&q_count, (uint32_t)__popc(ballot));
                    wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
                    if (hit) queue[wbase + __popc(ballot & ((1u << lane) - 1u))] = (uint16_t)qitem;
                }
                __syncthreads();
#ifdef DRAW_TAP_B
                const long long tb2 = clock64();
#endif
                // (3) fine raster
                const uint32_t nq = q_count;
                const float dxf = (float)(lane & 7), dyf = (float)(lane >> 3);
#pragma unroll 1
                for (uint32_t j = (uint32_t)warp; j < nq; j += TILE_THREADS / 32) {
                    const uint32_t qitem = queue[j];
                    const float *t = staged[qitem & 63u];
                    const float lx = fmaxf(t[S_X0], wx0f), hx = fminf(t[S_X1], wx1f), ly = fmaxf(t[S_Y0], wy0f), hy = fminf(t[S_Y1], wy1f);
                    const float x = FADD(FADD(lx, (float)(((qitem >> 6) & 7u) * 8u)), dxf);
                    const float y = FADD(FADD(ly, (float)((qitem >> 9) * 4u)), dyf);
                    if (x > hx || y > hy) continue;
                    unsigned long long *cell = &keys[(int)FSUB(y, ty0f) * TILE_W + (int)FSUB(x, tx0f)];
                    const uint32_t flags = __float_as_uint(t[S_FLAGS]);
                    float d;
                    if (!(flags & TRI_SLOW)) {
                        const float e0 = FSUB(FADD(FADD(FMUL(t[S_ECX], x), FMUL(t[S_ECY], y)), t[S_EK1]), t[S_EK2]);
                        const float e1 = FSUB(FADD(FADD(FMUL(t[S_ECX + 1], x), FMUL(t[S_ECY + 1], y)), t[S_EK1 + 1]), t[S_EK2 + 1]);
                        const float e2 = FSUB(FADD(FADD(FMUL(t[S_ECX + 2], x), FMUL(t[S_ECY + 2], y)), t[S_EK1 + 2]), t[S_EK2 + 2]);
                        if (!(e0 > ((flags & 1u) ? -0.5f : 0.0f) && e1 > ((flags & 2u) ? -0.5f : 0.0f) && e2 > ((flags & 4u) ? -0.5f : 0.0f)))
                            continue;
                        const float da = t[S_DA], db = t[S_DB], dc = t[S_DC];
                        const float rf0 = t[S_RF], rf1 = t[S_RF + 1], rf2 = t[S_RF + 2];
                        // early depth reject in key space (exactly conservative, see TRI_EARLYZ)
                        if ((flags & TRI_EARLYZ) &&
                            depth_key(fmaf(e2, FMUL(dc, rf2), fmaf(e1, FMUL(db, rf1), e0 * FMUL(da, rf0))) * EARLYZ_SCALE) > (uint32_t)(*cell >> 32))
                            continue;
                        float alpha, beta, gama;
                        if (flags & TRI_FASTDIV) {
                            alpha = exact_div(e0, t[S_F], rf0); beta = exact_div(e1, t[S_F + 1], rf1); gama = exact_div(e2, t[S_F + 2], rf2);
                        } else {
                            alpha = FDIV(e0, t[S_F]); beta = FDIV(e1, t[S_F + 1]); gama = FDIV(e2, t[S_F + 2]);
                        }
                        d = FADD(FADD(FMUL(alpha, da), FMUL(beta, db)), FMUL(gama, dc)); // canvas.rs:682
                    } else {
                        const TriRegs tr = tri_from_words(t);
                        if (!cover_pixel(tr, flags, x, y, d)) continue;
                    }
                    if (!(d < depth_max)) continue;
                    const unsigned long long key = make_key(d, __float_as_uint(t[S_SLOT]));
                    if (key < *cell) atomicMin(cell, key);
                }
#ifdef DRAW_TAP_B
                __syncthreads();
                tap_stage += tb1 - tb0; tap_coarse += tb2 - tb1; tap_fine += clock64() - tb2;
#endif
            }
        }
#if DRAW_TAP_B == 1
        if (W.tile_cycles && tid == 0) { // debug build: phase taps replaced by B1's stage / stage+coarse totals
            atomicMax(&W.tile_cycles[U.n_coarse + tile], (uint32_t)tap_stage);
            atomicMax(&W.tile_cycles[2 * U.n_coarse + tile], (uint32_t)(tap_stage + tap_coarse));
            atomicMax(&W.tile_cycles[tile], (uint32_t)(tap_stage + tap_coarse + tap_fine));
        }
#endif
#if DRAW_TAP_B == 2
        if (W.tile_cycles && tid == 0) atomicMax(&W.tile_cycles[U.n_coarse + tile], (uint32_t)tap_filter); // medium filter
#endif
    }
#if DRAW_TAP_B == 2
    const long long tb2_0 = clock64();
#endif
    // ---- phase B2: small triangles, one per lane, atomicMin on the key --------------------------
    // (item j of a round goes to lane j / warps of warp j % warps, so a short list spreads over all warps)
    for (uint32_t i = (uint32_t)(lane * (TILE_THREADS / 32) + warp); i < s_count; i += TILE_THREADS) {
        const uint32_t slot = W.s_refs[s_begin + i].x;
        const uint4 *pq = reinterpret_cast<const uint4 *>(prep + slot);
        const uint4 bb = __ldg(pq + 5); // x0 x1 y0 y1
        const int lx = max((int)__uint_as_float(bb.x), wx0), hx = min((int)__uint_as_float(bb.y), wx0 + ww - 1);
        const int ly = max((int)__uint_as_float(bb.z), wy0), hy = min((int)__uint_as_float(bb.w), wy0 + wh - 1);
        if (lx > hx || ly > hy) continue; // misses the window
        float w[PREP_WORDS];
#pragma unroll
        for (int q = 0; q < 7; q++) {
            if (q == 5) continue;
            const uint4 v = __ldg(pq + q);
            w[4 * q] = __uint_as_float(v.x); w[4 * q + 1] = __uint_as_float(v.y); w[4 * q + 2] = __uint_as_float(v.z); w[4 * q + 3] = __uint_as_float(v.w);
        }
        w[S_X0] = w[S_X1] = w[S_Y0] = w[S_Y1] = 0.0f;
        const TriRegs t = tri_from_words(w);
        float y = (float)ly;
        for (int yi = ly; yi <= hy; yi++, y = FADD(y, 1.0f)) {
            float x = (float)lx;
            for (int xi = lx; xi <= hx; xi++, x = FADD(x, 1.0f)) {
                float d;
                if (!cover_pixel(t, t.flags, x, y, d)) continue;
                if (!(d < depth_max)) continue; // also rejects NaN; equality with the clear depth fails `<`
                const unsigned long long key = make_key(d, slot);
                unsigned long long *cell = &keys[(yi - ty0) * TILE_W + (xi - tx0)];
                if (key < *cell) atomicMin(cell, key);
            }
        }
    }
    __syncthreads();

#ifndef DRAW_TAP_B
    if (W.tile_cycles && tid == 0) atomicMax(&W.tile_cycles[2 * U.n_coarse + tile], (uint32_t)(clock64() - t_start)); // end of phase B
#endif
#if DRAW_TAP_B == 2
    if (W.tile_cycles && tid == 0) atomicMax(&W.tile_cycles[2 * U.n_coarse + tile], (uint32_t)(clock64() - tb2_0) + W.tile_cycles[U.n_coarse + tile]); // + B2
#endif
    // ---- phase C: deferred shading, fused clear ----------------------------------------------------
    // The records of all the lane's winners are requested first (prefetch into L1: no registers held), so
    // that the pixel loop below pays the L2 round trip once and not once per pixel.
#pragma unroll 1
    for (int q = tid; q < n_win; q += TILE_THREADS) {
        const int x = wx0 + (q & (ww - 1)), y = wy0 + (q >> ww_shift);
        const uint32_t slot = (uint32_t)keys[(y - ty0) * TILE_W + (x - tx0)];
        if (slot == NO_SLOT) continue;
        const char *sp = reinterpret_cast<const char *>(W.srec + slot);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(prep + slot));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(sp));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + sizeof(ShadeRec) - 16));
    }
#pragma unroll 1
    for (int q = tid; q < n_win; q += TILE_THREADS) { // the window's pixels
        const int x = wx0 + (q & (ww - 1)), y = wy0 + (q >> ww_shift);
        const int p = (y - ty0) * TILE_W + (x - tx0);
        const uint32_t slot = (uint32_t)keys[p];
        uint32_t c = 155u | (186u << 8) | (255u << 16) | (255u << 24); // azul_bb, pad 255 (canvas.rs:131)
        float d = depth_max;
        uint32_t id = NO_SLOT;
        if (slot != NO_SLOT) {
            float op;
            c = shade_pixel_prep(S.materials, S.texels, u8tab, prep + slot, W.srec + slot, (float)x, (float)y, &d, &op, &id) |
                (255u << 24);
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
        if ((uint32_t)tid < n) stage_triangle(staged[tid], W.t_rrec + base + tid, base + tid);
        __syncthreads();
#pragma unroll 1
        for (uint32_t k = 0; k < n; k++) {
            const float *s = staged[k];
            if (s[S_X1] < wx0f || s[S_X0] > wx1f || s[S_Y1] < wy0f || s[S_Y0] > wy1f) continue;
            const TriRegs t = tri_from_words(s);
            const uint32_t tslot = __float_as_uint(s[S_SLOT]), tid_draw = __float_as_uint(s[S_ID]);
#pragma unroll 1
            for (int q = tid; q < n_win; q += TILE_THREADS) {
                const int xi = wx0 + (q & (ww - 1)), yi = wy0 + (q >> ww_shift);
                const int p = (yi - ty0) * TILE_W + (xi - tx0);
                const float x = (float)xi, y = (float)yi;
                if (x < s[S_X0] || x > s[S_X1] || y < s[S_Y0] || y > s[S_Y1]) continue;
                float d;
                if (!cover_pixel(t, t.flags | TRI_SLOW, x, y, d)) continue;
                const unsigned long long key = keys[p];
                const uint32_t wid = (uint32_t)key;
                if (!(wid == NO_SLOT || tid_draw > wid)) continue;           // drawn before the opaque winner: overwritten
                if (!(d < __uint_as_float((uint32_t)(key >> 32)))) continue; // canvas.rs:923, depth write is off
                const RasterRec r = load_raster(W.t_rrec + tslot);
                float d2, op;
                const uint32_t rgb = shade_pixel(S.materials, S.texels, u8tab, r, W.t_srec + tslot, x, y, &d2, &op);
                // canvas.rs:916-921: opacity < 1 blends with the stored colour, else replaces it
                colour[p] = op < 1.0f ? blend_rgb(colour[p], rgb, op) : (rgb | (255u << 24));
            }
        }
    }

    // ---- phase E: single write-back -------------------------------------------------------------
#pragma unroll 1
    for (int q = tid; q < n_win; q += TILE_THREADS) {
        const int x = wx0 + (q & (ww - 1)), y = wy0 + (q >> ww_shift);
        const int p = (y - ty0) * TILE_W + (x - tx0);
        if (x >= W_ || y >= H_) continue;
        const uint32_t c = colour[p]; // r g b pad -> memory order b g r pad
        __stcs(reinterpret_cast<uint32_t *>(color) + (size_t)(H_ - 1 - y) * W_ + x,
               ((c >> 16) & 255u) | (c & 0x0000FF00u) | ((c & 255u) << 16) | (c & 0xFF000000u));
        __stcs(depth + (size_t)y * W_ + x, __uint_as_float((uint32_t)(keys[p] >> 32)));
    }
#if !defined(DRAW_TAP_B) || DRAW_TAP_B == 2
    if (W.tile_cycles) {
        __syncthreads();
        if (tid == 0) atomicMax(&W.tile_cycles[tile], (uint32_t)(clock64() - t_start));
    }
#endif
}

// One empty tile written by the whole CTA (clear-in-tile mode, FrameUniforms::clear_in_tile): 16 bytes of
// colour and 16 of depth per thread, fire-and-forget streaming stores that drain while the CTA works on
// its raster item.
__device__ __forceinline__ void clear_tile_cta(const uint32_t et, const int W_, const int H_, const float depth_max,
                                               uint8_t *__restrict__ color, float *__restrict__ depth, const int tid) {
    const int ex0 = (int)(et & (MAX_TILES_X - 1)) * TILE_W, ey0 = (int)(et >> 10) * TILE_H;
    const uint32_t clear_px = 255u | (186u << 8) | (155u << 16) | (255u << 24); // memory order b g r pad
    if ((W_ & 3) == 0) {
        constexpr int QPR = TILE_W / 4; // 16-byte quads per tile row
#pragma unroll
        for (int q = tid; q < QPR * TILE_H; q += TILE_THREADS) {
            const int x = ex0 + (q % QPR) * 4, y = ey0 + q / QPR;
            if (x < W_ && y < H_) {
                __stcs(reinterpret_cast<uint4 *>(color + ((size_t)(H_ - 1 - y) * W_ + x) * 4),
                       make_uint4(clear_px, clear_px, clear_px, clear_px));
                __stcs(reinterpret_cast<float4 *>(depth + (size_t)y * W_ + x), make_float4(depth_max, depth_max, depth_max, depth_max));
            }
        }
    } else {
        for (int p = tid; p < TILE_PIXELS; p += TILE_THREADS) {
            const int x = ex0 + (p & (TILE_W - 1)), y = ey0 + p / TILE_W;
            if (x >= W_ || y >= H_) continue;
            __stcs(reinterpret_cast<uint32_t *>(color) + (size_t)(H_ - 1 - y) * W_ + x, clear_px);
            __stcs(depth + (size_t)y * W_ + x, depth_max);
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_clear_empty : the tiles nothing was binned to get the clear colour and depth (canvas.rs:425-433).
// They are most of a frame's bytes and none of its arithmetic, so they have their own launch: it
// starts as soon as k_alloc has listed them, on the canvas' stream, and streams to HBM while the
// latency-bound rest of the frame (k_bin<fill>, k_raster, the dense tiles of k_tile — and the previous
// frame's) leaves the memory system idle.  One tile per warp, two 256-byte rows (colour) per store
// instruction, pointers stepped by a row pair.
// ------------------------------------------------------------------------------------------
constexpr int CLEAR_THREADS = 256;
#ifndef DRAW_CLEAR_MINB
#define DRAW_CLEAR_MINB 1
#endif
__global__ void __launch_bounds__(CLEAR_THREADS, DRAW_CLEAR_MINB) k_clear_empty(const FrameUniforms *__restrict__ Up, const FrameDev W) {
    const FrameUniforms &U = *Up; // per-frame uniforms, device-resident (one upload per frame; the launches never change)
    uint8_t *__restrict__ color = U.color;
    float *__restrict__ depth = U.depth;
    const CtaTrace trace_(W, 7u);
    pdl_prologue(false);
    const uint32_t n_empty = W.counters[13];
    const int lane = threadIdx.x & 31;
    const int W_ = (int)U.canvas_w, H_ = (int)U.canvas_h;
    const float depth_max = U.depth_max;
    const uint32_t n_warps = gridDim.x * (CLEAR_THREADS / 32);
    for (uint32_t e = blockIdx.x * (CLEAR_THREADS / 32) + (threadIdx.x >> 5); e < n_empty; e += n_warps) {
        const uint32_t et = W.empty_tiles[e];
        const int ex0 = (int)(et & (MAX_TILES_X - 1)) * TILE_W, ey0 = (int)(et >> 10) * TILE_H;
        const uint32_t clear_px = 255u | (186u << 8) | (155u << 16) | (255u << 24); // memory order b g r pad
        if ((W_ & 3) == 0) {
            constexpr int QPR = TILE_W / 4, RPI = 32 / QPR; // 16-byte quads per tile row, rows per warp store
            const int x = ex0 + (lane % QPR) * 4, r0 = lane / QPR;
            const int rows = min(TILE_H, H_ - ey0);
            if (x < W_) {
                uint4 *cp = reinterpret_cast<uint4 *>(color + ((size_t)(H_ - 1 - ey0 - r0) * W_ + x) * 4);
                float4 *dp = reinterpret_cast<float4 *>(depth + (size_t)(ey0 + r0) * W_ + x);
                const ptrdiff_t step = (ptrdiff_t)RPI * (W_ / 4); // in 16-byte units: colour rows go up, depth rows down
                const uint4 cv = make_uint4(clear_px, clear_px, clear_px, clear_px);
                const float4 dv = make_float4(depth_max, depth_max, depth_max, depth_max);
#pragma unroll 4
                for (int r = r0; r < rows; r += RPI, cp -= step, dp += step) {
                    __stcs(cp, cv); // streaming stores: the frame is written once and not read back, keep L2 for the records
                    __stcs(dp, dv);
                }
            }
        } else {
            for (int p = lane; p < TILE_PIXELS; p += 32) {
                const int x = ex0 + (p & (TILE_W - 1)), y = ey0 + p / TILE_W;
                if (x >= W_ || y >= H_) continue;
                __stcs(reinterpret_cast<uint32_t *>(color) + (size_t)(H_ - 1 - y) * W_ + x, clear_px);
                __stcs(depth + (size_t)y * W_ + x, depth_max);
            }
        }
        if (W.tile_cycles && lane == 0) atomicMax(&W.tile_cycles[(et >> 10) * U.tiles_x + (et & (MAX_TILES_X - 1))], 1u);
    }
}

// Persistent CTAs (two per SM): each takes the next item of the work list — heaviest first — until the
// list is exhausted, so that no CTA is launched just to find out that there is nothing to do, and the
// per-CTA set-up is paid once.  Thread 0 keeps one list index and one item in flight ahead of the
// item being processed (the cursor's atomicAdd and the list load are L2 round trips).
#ifdef DRAW_TILE_MAXNREG // leave registers free for a co-resident k_clear_empty CTA (experiments)
__global__ void __maxnreg__(DRAW_TILE_MAXNREG) k_tile(
#else
__global__ void __launch_bounds__(TILE_THREADS, 1024 / TILE_THREADS) k_tile(
#endif
    const FrameUniforms *__restrict__ Up, const SceneDev S,
                                                                            const FrameDev W, const uint32_t n_slots) {
    const FrameUniforms &U = *Up; // per-frame uniforms, device-resident (one upload per frame; the launches never change)
    uint8_t *__restrict__ color = U.color;
    float *__restrict__ depth = U.depth;
    __shared__ uint32_t s_item, s_index;
    __shared__ float u8tab[256]; // (u8 as f32) / 255.0, filled once per CTA (visible after the loop's first barrier)
    fill_u8_table(u8tab, threadIdx.x, TILE_THREADS);
    const CtaTrace trace_(W, 8u);
    pdl_prologue();
    // The first item of a CTA is its own index; the cursor (in a cache line of its own: a load that shares
    // a line with a contended atomic queues behind it) hands out the rest.
    const bool usable = W.counters[2] == 0; // a work buffer overflowed: lists are unusable, the host re-renders
    // The frame's counters (record / reference totals, overflow flags, statistics: final since k_alloc) go to the
    // canvas' pinned host memory as sixteen posted stores — no copy-engine transfer that would queue behind the
    // frames' 33-132 MB read-backs, no extra launch.  Visible to the host once the kernel has completed.
    if (blockIdx.x == 0 && threadIdx.x < 16 && U.status_host) {
        U.status_host[threadIdx.x] = __ldcg(&W.counters[threadIdx.x]);
        __threadfence_system();
    }
    // Clear-in-tile mode: the empty tiles (k_alloc's list) are dealt out evenly over the raster items and
    // written by the item's CTA before it starts on the item; there is no k_clear_empty launch.  The
    // stores need no answer, so they drain to HBM under the latency-bound raster work instead of
    // holding every SM for a launch of their own.
    const uint32_t n_empty = U.clear_in_tile ? W.counters[13] : 0u, n_items = W.counters[15];
    const uint32_t per_item = n_items ? (n_empty + n_items - 1u) / n_items : 0u;
    const int W_ = (int)U.canvas_w, H_ = (int)U.canvas_h;
    uint32_t *cursor = W.counters + ITEM_CURSOR;
    uint32_t item = ITEM_NONE, index = blockIdx.x, ahead = 0; // thread 0 only
    if (threadIdx.x == 0) {
        ahead = gridDim.x + atomicAdd(cursor, 1u);
        item = blockIdx.x < n_slots ? __ldcg(&W.tile_order[blockIdx.x]) : ITEM_NONE;
    }
    while (true) {
        uint32_t next_item = ITEM_NONE, next_ahead = 0;
        if (threadIdx.x == 0) {
            s_item = item;
            s_index = index;
            next_item = ahead < n_slots ? __ldcg(&W.tile_order[ahead]) : ITEM_NONE; // consumed after the item below
            next_ahead = gridDim.x + atomicAdd(cursor, 1u);
        }
        __syncthreads();
        const uint32_t cur = s_item, cur_index = s_index;
        if (cur == ITEM_NONE) break; // items are contiguous; the slots after them hold ITEM_NONE
        // mode 1: before the item; 2: after it; 3: before in even CTAs, after in odd ones (the two CTAs of an SM
        // are then rarely both storing)
        const bool clear_first = U.clear_in_tile == 1u || (U.clear_in_tile == 3u && (blockIdx.x & 1u) == 0u);
        if (clear_first)
            for (uint32_t e = cur_index * per_item, e_end = min(n_empty, e + per_item); e < e_end; e++)
                clear_tile_cta(__ldg(&W.empty_tiles[e]), W_, H_, U.depth_max, color, depth, threadIdx.x);
        tile_item(cur, U, S, W, color, depth, u8tab, usable);
        if (!clear_first)
            for (uint32_t e = cur_index * per_item, e_end = min(n_empty, e + per_item); e < e_end; e++)
                clear_tile_cta(__ldg(&W.empty_tiles[e]), W_, H_, U.depth_max, color, depth, threadIdx.x);
        __syncthreads(); // shared memory (and s_item) are reused by the next item
        item = next_item;
        index = ahead;
        ahead = next_ahead;
    }
    if (n_items == 0u) // nothing to rasterise in this stripe: the CTAs share the empty tiles
        for (uint32_t e = blockIdx.x; e < n_empty; e += gridDim.x)
            clear_tile_cta(__ldg(&W.empty_tiles[e]), W_, H_, U.depth_max, color, depth, threadIdx.x);
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
extern thread_local unsigned g_tile_ctas;
uint32_t tile_grid_items(const FrameUniforms &U); // k_binning.cu
void launch_tile(const FrameUniforms &U, const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream) {
    const uint32_t slots = tile_grid_items(U); // work-list slots (k_alloc fills the unused ones with ITEM_NONE)
    if (slots) launch_pdl(k_tile, min(slots, g_tile_ctas), TILE_THREADS, stream, dU, S, W, slots);
}

thread_local unsigned g_clear_ctas = 148u * 4u; // scene.cpp: DRAW_B200_CLEAR_CTAS
thread_local unsigned g_tile_ctas = 148u * (1024u / TILE_THREADS); // scene.cpp: DRAW_B200_TILE_CTAS (persistent CTAs of k_tile)
void launch_clear_empty(const FrameUniforms &U, const FrameUniforms *dU, const FrameDev &W, cudaStream_t stream) {
    if (tile_grid_items(U)) launch_pdl(k_clear_empty, g_clear_ctas, CLEAR_THREADS, stream, dU, W);
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
