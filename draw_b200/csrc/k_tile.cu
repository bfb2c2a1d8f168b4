// k_tile.cu — per-tile raster / depth / shade kernel (mororo18/draw canvas.rs:577-750, 906-960).
//
// One CTA (TILE_THREADS = 256 threads) per 64x32-pixel tile or pixel window of one (persistent CTAs, k_front's
// work list).  The tile's depth, winning record and colour stay on chip (registers, then shared memory) for
// the whole item; colour and depth go to HBM exactly once at the end, with the clear fused in.  No tensor cores:
// nothing here is a contraction.
//
//   phase A  the tile's key page (k_raster's medium / small fragments) becomes the starting depth / winner of
//            every pixel; "large" list: triangles are staged through shared memory 64 at a time; every lane owns
//            a 4 x BLK_H pixel block (warp = REGION x REGION_H region) and tests it against each triangle, after a
//            warp-level bbox reject and an exact block-level edge reject.  Depth/winner in registers.
//   merge    each lane publishes its 8 pixels as 64-bit keys (depth, slot) in shared memory.
//   phase C  deferred shading, one pixel per lane per step: only the winner of a pixel is shaded
//            (canvas.rs:685-743); the key becomes (exact depth, draw id), colour goes to smem.
//   phase D  the tile's transparent triangles in draw order, blended over the shaded colour (scene/mod.rs:1088-1246).
//   phase E  write-back, 16 bytes per lane and buffer: colour rows y-flipped (canvas.rs:955-956), depth
//            rows not (canvas.rs:413-423).
//
// Draw-order semantics without ordered lists: the reference draws triangles sequentially with a
// strict `<` depth test (canvas.rs:923), so for opaque triangles the surviving fragment of a pixel
// is the minimum of (depth, draw order) — ties go to the earlier triangle.  Record slots are
// allocated in draw order by k_front, so the key (depth, slot) ordered as an unsigned 64-bit
// integer is exactly that minimum, and lists can be consumed in any order by any lane.
// Transparent triangles (depth test on, depth write off, blend with the current colour,
// scene/mod.rs:1088) are only visible over the final opaque winner W of their pixel if drawn after
// it: fragment T is blended iff id(T) > id(W) and depth(T) < depth(W), in draw order.  Their slots
// 4 * ordinal + k ARE the draw order (the painter sort of k_sort permutes the index streams in place), so a
// tile's unordered list of transparent references is put in order on chip: slots are unique, so a window of
// 1024 consecutive slot values holds at most 1024 references, each placed at table[slot - window start].
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
// block_loc: the opaque records' slot -> storage table (record_index, device_math.cuh); nullptr for the transparent
// records, which are stored at their slot.  The staged copy keeps the slot number (the order of the draws).
__device__ __forceinline__ void stage_copy(float (*staged)[STAGE_STRIDE], const PrepRec *__restrict__ prep, const uint32_t *__restrict__ block_loc,
                                           const uint32_t *refs, uint32_t ref_stride, uint32_t n, int tid) {
    for (uint32_t i = (uint32_t)tid; i < n * 8u; i += TILE_THREADS) {
        const uint32_t k = i >> 3, q = i & 7u;
        if (q == 7u) continue;
        const uint32_t slot = refs[k * ref_stride];
        const uint32_t at = block_loc ? __ldg(block_loc + (slot >> SLOT_SHIFT)) + (slot & (SLOT_STRIDE - 1u)) : slot;
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(prep + at) + q);
        float *d = staged[k] + 4 * q;
        d[0] = __uint_as_float(v.x); d[1] = __uint_as_float(v.y); d[2] = __uint_as_float(v.z);
        d[3] = q == 6u ? __uint_as_float(slot) : __uint_as_float(v.w);
    }
}
// Window filter: keeps the references whose bbox meets the window (only their bbox quad is read).
// Returns the number kept; cand[] is valid after the call (ends with a barrier).
__device__ __forceinline__ uint32_t filter_refs(const uint32_t *__restrict__ list, uint32_t list_stride, uint32_t count,
                                                const PrepRec *__restrict__ prep, const uint32_t *__restrict__ block_loc, uint32_t *cand, uint32_t *s_count,
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
            const uint4 bb = __ldg(reinterpret_cast<const uint4 *>(prep + (__ldg(block_loc + (slot >> SLOT_SHIFT)) + (slot & (SLOT_STRIDE - 1u)))) + 5); // x0 x1 y0 y1
            keep = !(__uint_as_float(bb.y) < wx0f || __uint_as_float(bb.x) > wx1f ||
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

// Front-to-back order for a segment of a tile's large references (n <= CAND_CAP): the depth resolve does not depend on
// the order the triangles are tested in (ties go to the smaller slot), but the early depth rejects do — with the
// nearest triangles first, most of the others are turned away per block instead of per pixel.  Key = the smallest
// vertex depth of the record (a lower bound of its depths) over the slot; bitonic sort in `scratch` (the tile's key
// array, not in use before the end of phase A); the slots come back in cand[].  Ends with a barrier.
__device__ __forceinline__ void sort_refs_front_to_back(const uint32_t *refs, uint32_t n, const PrepRec *__restrict__ prep,
                                                        const uint32_t *__restrict__ block_loc, unsigned long long *scratch, uint32_t *cand,
                                                        int tid) {
    uint32_t n2 = 32;
    while (n2 < n) n2 <<= 1;
    for (uint32_t i = (uint32_t)tid; i < n2; i += TILE_THREADS) {
        unsigned long long key = ~0ull;
        if (i < n) {
            const uint32_t slot = refs[i];
            const uint4 *q = reinterpret_cast<const uint4 *>(prep + (__ldg(block_loc + (slot >> SLOT_SHIFT)) + (slot & (SLOT_STRIDE - 1u))));
            const uint4 q4 = __ldg(q + 4), q6 = __ldg(q + 6); // .. da db | dc ..
            const float zmin = fminf(fminf(__uint_as_float(q4.z), __uint_as_float(q4.w)), __uint_as_float(q6.x));
            key = ((unsigned long long)depth_key(zmin) << 32) | slot;
        }
        scratch[i] = key;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= n2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = (uint32_t)tid; i < n2; i += TILE_THREADS) {
                const uint32_t o = i ^ j;
                if (o > i) {
                    const unsigned long long a = scratch[i], b = scratch[o];
                    if ((a > b) == ((i & k) == 0u)) {
                        scratch[i] = b;
                        scratch[o] = a;
                    }
                }
            }
            __syncthreads();
        }
    for (uint32_t i = (uint32_t)tid; i < n; i += TILE_THREADS) cand[i] = (uint32_t)scratch[i];
    __syncthreads();
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

constexpr int TWIN = CAND_CAP; // phase D: slot values per ordering window
constexpr int MAT_CACHE = 128; // materials kept in shared memory by k_tile (8 KB)
// debug taps (FrameDev::tile_cycles, 4 x n_coarse words): SM cycles since the item started at the end of each phase
enum : int { TAP_TOTAL = 0, TAP_A = 1, TAP_C = 2, TAP_D = 3 };

// One work item (k_front, device_types.h): a tile or one pixel window of a dense tile.
__device__ __forceinline__ void tile_item(const uint32_t item, const FrameUniforms &U, const SceneDev &S, const FrameDev &W,
                                          uint8_t *__restrict__ color, float *__restrict__ depth, const float *u8tab,
                                          const MaterialDev *mats, const bool usable) {
    __shared__ __align__(16) unsigned long long keys[TILE_PIXELS]; // (depth key, slot), later (depth bits, draw id)
    __shared__ __align__(16) uint32_t colour[TILE_PIXELS];         // r | g << 8 | b << 16 | pad << 24
    __shared__ float staged[CHUNK][STAGE_STRIDE];
    __shared__ uint32_t cand[CAND_CAP];
    __shared__ uint32_t s_cand_count, s_min, s_max, s_warp_sum[TILE_THREADS / 32], s_strips;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W_ = (int)U.canvas_w, H_ = (int)U.canvas_h;
    const float depth_max = U.depth_max;

    const uint32_t tile_x = item & (MAX_TILES_X - 1), tile_y = (item >> 10) & (MAX_TILES_Y - 1);
    const uint32_t tile = tile_y * U.tiles_x + tile_x;
    const long long t_start = W.tile_cycles ? clock64() : 0;
    const int tx0 = (int)tile_x * TILE_W, ty0 = (int)tile_y * TILE_H;

    // pixel window of the tile this CTA renders (the whole tile unless k_front cut the tile up)
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

    const uint32_t l_begin = W.l_offset[tile], t_begin = W.t_offset[tile];
    const uint32_t l_count = usable ? W.l_count[tile] : 0u, t_count = usable ? W.t_count[tile] : 0u; // the fill cursors end at the counts
    const bool paged = usable && W.ms_weight[tile] != 0u; // k_raster may have put fragments in the tile's key page
    const PrepRec *__restrict__ prep = W.prep;

    // ---- phase A: key page, large triangles; every lane tests its own 4 x BLK_H block -------------
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
        float lane_zmax = depth_max; // the largest of zb[]: kept up to date where zb[] changes, read by every block-level depth reject
        auto refresh_zmax = [&]() {
            lane_zmax = zb[0];
#pragma unroll
            for (int p = 1; p < PX; p++) lane_zmax = fmaxf(lane_zmax, zb[p]);
        };
        // The tile's key page (k_raster's result): its loads are issued here, together with the stores that
        // leave it empty for the next frame, and consumed after the first chunk of large triangles has been
        // staged, so that the two L2 round trips overlap.  A lane's 4 pixels of a row are 32 contiguous bytes.
        // (Issued for every tile, before it is known whether the page holds anything — an untouched page reads as empty
        // keys: the round trip overlaps the one of the tile's list header above instead of following it.)
        const bool have_page = warp_in;
        ulonglong2 pk01[BLK_H], pk23[BLK_H];
        if (have_page) {
            unsigned long long *pk = W.key_pages + (size_t)tile * TILE_PIXELS;
#pragma unroll
            for (int j = 0; j < BLK_H; j++) {
                ulonglong2 *src = reinterpret_cast<ulonglong2 *>(pk + (by0 - ty0 + j) * TILE_W + (bx0 - tx0));
                pk01[j] = __ldcg(src);
                pk23[j] = __ldcg(src + 1);
            }
        }
        bool page_pending = have_page && paged;
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
            refresh_zmax();
            // Leave the page empty for the next frame: after every lane of the warp has its keys (the loads
            // above are consumed), the warp's region — REGION_H rows of 128 bytes — is overwritten with whole
            // 128-byte lines, 8 lanes per row.
            __syncwarp();
            unsigned long long *pk = W.key_pages + (size_t)tile * TILE_PIXELS;
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
                n_refs = filter_refs(refs, 1u, seg_n, prep, W.block_loc, cand, &s_cand_count, wx0f, wx1f, wy0f, wy1f, tid);
                refs = cand;
            }
            if (U.sort_large && n_refs >= U.sort_large) {
                sort_refs_front_to_back(refs, n_refs, prep, W.block_loc, keys, cand, tid);
                refs = cand;
            }
#pragma unroll 1
            for (uint32_t base = 0; base < n_refs; base += CHUNK) {
                const uint32_t n = min((uint32_t)CHUNK, n_refs - base);
                __syncthreads();
                stage_copy(staged, prep, W.block_loc, refs + base, 1u, n, tid);
                __syncthreads();
                if (page_pending) merge_page();
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
                        const float thr0 = (flags & 1u) ? -0.5f : 0.0f, thr1 = (flags & 2u) ? -0.5f : 0.0f,
                                    thr2 = (flags & 4u) ? -0.5f : 0.0f;
                        const float da = s[S_DA], db = s[S_DB], dc = s[S_DC];
                        const float rf0 = s[S_RF], rf1 = s[S_RF + 1], rf2 = s[S_RF + 2];
                        const float g0 = FMUL(da, rf0), g1 = FMUL(db, rf1), g2 = FMUL(dc, rf2);
                        if (flags & TRI_EARLYZ) {
                            // Early depth reject for the whole block.  Every edge value is monotone in x and in y (each
                            // rounding is), so its minimum over the block sits at a corner; the approximate depth
                            // fma(e2, g2, fma(e1, g1, e0 * g0)) is monotone in each e (g >= 0 under TRI_EARLYZ), so L below
                            // bounds it from below for every pixel of the block.  If L * EARLYZ_SCALE exceeds the block's
                            // largest stored depth, the per-pixel test below rejects every pixel: same result, 8 x 3 edge
                            // evaluations not done (overdrawn scenes: most triangles of a tile are hidden).
                            float em[3];
#pragma unroll
                            for (int e = 0; e < 3; e++) {
                                const float cx = s[S_ECX + e], cy = s[S_ECY + e];
                                const float xm = cx >= 0.0f ? lo_x : hi_x, ym = cy >= 0.0f ? lo_y : hi_y;
                                em[e] = FSUB(FADD(FADD(FMUL(cx, xm), FMUL(cy, ym)), s[S_EK1 + e]), s[S_EK2 + e]);
                            }
                            if (fmaf(em[2], g2, fmaf(em[1], g1, em[0] * g0)) * EARLYZ_SCALE > lane_zmax) continue;
                        }
                        // block-level reject at the best corner of the clipped block (exact, see rect_may_cover); after the depth
                        // reject: in an overdrawn tile most triangles are hidden, which is the cheaper thing to find out
                        if (!staged_may_cover(s, flags, lo_x, hi_x, lo_y, hi_y)) continue;
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
                            refresh_zmax();
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
                            refresh_zmax();
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
                        refresh_zmax();
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
    if (W.tile_cycles && tid == 0) atomicMax(&W.tile_cycles[TAP_A * U.n_coarse + tile], (uint32_t)(clock64() - t_start));

    // ---- phase C: deferred shading, fused clear ----------------------------------------------------
    if (tid == 0) s_strips = 0; // phase E: which 64x8 strips of the tile hold something else than the clear colour (ordered by the barriers between)
    // The records of all the lane's winners are requested first (prefetch into L1: no registers held), so
    // that the pixel loop below pays the L2 round trip once and not once per pixel.
    // (Measured alternatives that did not pay: two pixels per lane as independent streams; per-warp staging of the
    // distinct winners' records in shared memory — the phase is bound by instruction issue, ~250 per 32 pixels.)
#pragma unroll 1
    for (int q = tid; q < n_win; q += TILE_THREADS) {
        const int x = wx0 + (q & (ww - 1)), y = wy0 + (q >> ww_shift);
        const int p = (y - ty0) * TILE_W + (x - tx0);
        const uint32_t slot = (uint32_t)keys[p];
        if (slot == NO_SLOT) continue;
        // the winner is decided: from here on the pixel carries where its record is stored instead of the slot number
        // (own pixel in both loops: no barrier)
        const uint32_t at = record_index(W, slot);
        reinterpret_cast<uint32_t *>(keys + p)[0] = at;
        const char *sp = reinterpret_cast<const char *>(W.srec + at);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(prep + at));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(sp));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + sizeof(ShadeRec) - 16));
    }
#pragma unroll 1
    for (int q = tid; q < n_win; q += TILE_THREADS) { // the window's pixels
        const int x = wx0 + (q & (ww - 1)), y = wy0 + (q >> ww_shift);
        const int p = (y - ty0) * TILE_W + (x - tx0);
        const uint32_t at = (uint32_t)keys[p]; // NO_SLOT, or the storage index written above
        uint32_t c = 155u | (186u << 8) | (255u << 16) | (255u << 24); // azul_bb, pad 255 (canvas.rs:131)
        float d = depth_max;
        uint32_t id = NO_SLOT;
        if (at != NO_SLOT) {
            float op;
            c = shade_pixel_prep<true>(mats, S.texels, u8tab, prep + at, W.srec + at, (float)x, (float)y, &d, &op, &id) | (255u << 24);
        }
        colour[p] = c;
        keys[p] = ((unsigned long long)__float_as_uint(d) << 32) | id; // own pixel: no sync needed
    }
    if (W.tile_cycles && tid == 0) atomicMax(&W.tile_cycles[TAP_C * U.n_coarse + tile], (uint32_t)(clock64() - t_start));

    // ---- phase D: the tile's transparent triangles in draw order (scene/mod.rs:1088-1246) ---------
    if (t_count) { // block-uniform
        // slot range of the list
        if (tid == 0) {
            s_min = 0xFFFFFFFFu;
            s_max = 0u;
        }
        __syncthreads();
        {
            uint32_t lo = 0xFFFFFFFFu, hi = 0u;
            for (uint32_t i = (uint32_t)tid; i < t_count; i += TILE_THREADS) {
                const uint32_t sl = W.t_refs[t_begin + i];
                lo = min(lo, sl);
                hi = max(hi, sl);
            }
            lo = __reduce_min_sync(0xFFFFFFFFu, lo);
            hi = __reduce_max_sync(0xFFFFFFFFu, hi);
            if (lane == 0) {
                atomicMin(&s_min, lo);
                atomicMax(&s_max, hi);
            }
        }
        __syncthreads();
        const uint32_t slot_lo = s_min, slot_hi = s_max;
#pragma unroll 1
        for (uint32_t w0 = slot_lo; w0 <= slot_hi; w0 += TWIN) { // ordering windows; w0 + TWIN cannot wrap (slots < 2^32 - TWIN)
            // (1) the window's references at table[slot - w0] (slots are unique: no collisions)
            __syncthreads(); // cand / staged of the previous window are no longer read
            for (int i = tid; i < TWIN; i += TILE_THREADS) cand[i] = NO_SLOT;
            __syncthreads();
            for (uint32_t i = (uint32_t)tid; i < t_count; i += TILE_THREADS) {
                const uint32_t sl = W.t_refs[t_begin + i];
                if (sl - w0 < (uint32_t)TWIN) cand[sl - w0] = sl;
            }
            __syncthreads();
            // (2) ordered compaction in place: every thread owns TWIN / TILE_THREADS consecutive entries
            constexpr int OWN = TWIN / TILE_THREADS;
            uint32_t own[OWN], n_own = 0;
#pragma unroll
            for (int j = 0; j < OWN; j++) {
                own[j] = cand[tid * OWN + j];
                n_own += own[j] != NO_SLOT ? 1u : 0u;
            }
            uint32_t incl = n_own;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += up;
            }
            if (lane == 31) s_warp_sum[warp] = incl;
            __syncthreads(); // also: every thread has read its entries
            uint32_t before = 0, n_ord = 0;
#pragma unroll
            for (int w = 0; w < TILE_THREADS / 32; w++) {
                const uint32_t t = s_warp_sum[w];
                if (w < warp) before += t;
                n_ord += t;
            }
            uint32_t at = before + incl - n_own;
#pragma unroll
            for (int j = 0; j < OWN; j++)
                if (own[j] != NO_SLOT) cand[at++] = own[j];
            // (3) the ordered references, staged 64 at a time; every lane walks them for its own pixels
#pragma unroll 1
            for (uint32_t base = 0; base < n_ord; base += CHUNK) {
                const uint32_t n = min((uint32_t)CHUNK, n_ord - base);
                __syncthreads(); // cand is complete / staged of the previous chunk is no longer read
                stage_copy(staged, W.t_prep, nullptr, cand + base, 1u, n, tid);
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
                        const uint32_t rgb = shade_pixel<true>(mats, S.texels, u8tab, r, W.t_srec + tslot, x, y, &d2, &op);
                        // canvas.rs:916-921: opacity < 1 blends with the stored colour, else replaces it
                        colour[p] = op < 1.0f ? blend_rgb(colour[p], rgb, op) : (rgb | (255u << 24));
                    }
                }
            }
        }
    }
    if (W.tile_cycles && tid == 0) atomicMax(&W.tile_cycles[TAP_D * U.n_coarse + tile], (uint32_t)(clock64() - t_start));

    // ---- phase E: single write-back -------------------------------------------------------------
    __syncthreads(); // a lane stores four neighbouring pixels, shaded by four different lanes
    if ((W_ & 3) == 0) {
        // four pixels per lane: one 16-byte store of colour, one of depth (rows are 16-byte aligned: W % 4 == 0, x % 4 == 0)
        const int qpr = ww >> 2, qpr_shift = ww_shift - 2;
#pragma unroll 1
        for (int q = tid; q < (n_win >> 2); q += TILE_THREADS) {
            const int x = wx0 + ((q & (qpr - 1)) << 2), y = wy0 + (q >> qpr_shift);
            bool content = false;
            if (x < W_ && y < H_) {
                const int p = (y - ty0) * TILE_W + (x - tx0);
                const uint4 c4 = *reinterpret_cast<const uint4 *>(&colour[p]); // r g b pad -> memory order b g r pad
                const ulonglong2 k01 = *reinterpret_cast<const ulonglong2 *>(&keys[p]), k23 = *reinterpret_cast<const ulonglong2 *>(&keys[p + 2]);
                uint4 o;
                o.x = __byte_perm(c4.x, 0, 0x3012); o.y = __byte_perm(c4.y, 0, 0x3012);
                o.z = __byte_perm(c4.z, 0, 0x3012); o.w = __byte_perm(c4.w, 0, 0x3012);
                __stcs(reinterpret_cast<uint4 *>(color + ((size_t)(H_ - 1 - y) * W_ + x) * 4), o);
                __stcs(reinterpret_cast<float4 *>(depth + (size_t)y * W_ + x),
                       make_float4(__uint_as_float((uint32_t)(k01.x >> 32)), __uint_as_float((uint32_t)(k01.y >> 32)),
                                   __uint_as_float((uint32_t)(k23.x >> 32)), __uint_as_float((uint32_t)(k23.y >> 32))));
                constexpr uint32_t CLEAR = 155u | (186u << 8) | (255u << 16) | (255u << 24); // phase C's value for a pixel nothing covers
                content = c4.x != CLEAR || c4.y != CLEAR || c4.z != CLEAR || c4.w != CLEAR;
            }
            // a whole tile: every thread makes the same number of steps and a warp's 32 quads are two rows of one strip
            if (!windowed && __any_sync(0xFFFFFFFFu, content) && lane == 0) atomicOr(&s_strips, 1u << ((y - ty0) >> 3));
        }
    } else {
#pragma unroll 1
        for (int q = tid; q < n_win; q += TILE_THREADS) {
            const int x = wx0 + (q & (ww - 1)), y = wy0 + (q >> ww_shift);
            const int p = (y - ty0) * TILE_W + (x - tx0);
            if (x >= W_ || y >= H_) continue;
            const uint32_t c = colour[p]; // r g b pad -> memory order b g r pad
            __stcs(reinterpret_cast<uint32_t *>(color) + (size_t)(H_ - 1 - y) * W_ + x, __byte_perm(c, 0, 0x3012));
            __stcs(depth + (size_t)y * W_ + x, __uint_as_float((uint32_t)(keys[p] >> 32)));
        }
    }
    if (W.tile_cycles) {
        __syncthreads();
        if (tid == 0) atomicMax(&W.tile_cycles[TAP_TOTAL * U.n_coarse + tile], (uint32_t)(clock64() - t_start));
    }
    // What k_mirror needs to know about the tile (k_mirror.cu): one bit per 64x8 strip that holds something else than the
    // clear colour.  Windows of a split tile, and canvases whose rows are not 16-byte multiples, say "all of it".
    if (U.tile_state) {
        __syncthreads();
        if (tid == 0) U.tile_state[tile] = (windowed || (W_ & 3) != 0) ? (uint8_t)TILE_STRIPS_ALL : (uint8_t)s_strips;
    }
}

// One empty tile written by the whole CTA: 16 bytes of colour and 16 of depth per thread, fire-and-forget
// streaming stores that drain while the CTA works on its raster item (canvas.rs:425-433).
__device__ __forceinline__ void clear_tile_cta(const uint32_t et, const int W_, const int H_, const float depth_max,
                                               uint8_t *__restrict__ color, float *__restrict__ depth, const int tid, const bool write_color) {
    const int ex0 = (int)(et & (MAX_TILES_X - 1)) * TILE_W, ey0 = (int)(et >> 10) * TILE_H;
    const uint32_t clear_px = 255u | (186u << 8) | (155u << 16) | (255u << 24); // memory order b g r pad
    if ((W_ & 3) == 0) {
        // A thread keeps its column of 16-byte quads and steps down the tile: the addresses are computed once per tile
        // (the clears of a 4K frame with a small scene are most of the frame's stores: ~4 000 tiles, 24 instructions per
        // store pair when every quad derived its own address).
        constexpr int QPR = TILE_W / 4, ROWS_PER_STEP = TILE_THREADS / QPR, STEPS = TILE_H / ROWS_PER_STEP; // 16 quads per row, 16 rows per step
        static_assert(TILE_THREADS % QPR == 0 && TILE_H % ROWS_PER_STEP == 0, "clear_tile_cta geometry");
        const int x = ex0 + (tid % QPR) * 4, y0 = ey0 + tid / QPR;
        if (x < W_) {
            uint4 *c = reinterpret_cast<uint4 *>(color + ((size_t)(H_ - 1 - y0) * W_ + x) * 4);
            float4 *d = reinterpret_cast<float4 *>(depth + (size_t)y0 * W_ + x);
            const ptrdiff_t step = (ptrdiff_t)ROWS_PER_STEP * (W_ >> 2); // in 16-byte units; colour rows run upwards (y-flipped)
            const uint4 cq = make_uint4(clear_px, clear_px, clear_px, clear_px);
            const float4 dq = make_float4(depth_max, depth_max, depth_max, depth_max);
#pragma unroll
            for (int k = 0; k < STEPS; k++)
                if (y0 + k * ROWS_PER_STEP < H_) {
                    if (write_color) __stcs(c - k * step, cq);
                    __stcs(d + k * step, dq);
                }
        }
    } else {
        for (int p = tid; p < TILE_PIXELS; p += TILE_THREADS) {
            const int x = ex0 + (p & (TILE_W - 1)), y = ey0 + p / TILE_W;
            if (x >= W_ || y >= H_) continue;
            if (write_color) __stcs(reinterpret_cast<uint32_t *>(color) + (size_t)(H_ - 1 - y) * W_ + x, clear_px);
            __stcs(depth + (size_t)y * W_ + x, depth_max);
        }
    }
}

// Persistent CTAs (three per SM): each takes the next item of the work list — heaviest first — until the
// list is exhausted, so that no CTA is launched just to find out that there is nothing to do, and the
// per-CTA set-up is paid once.  Thread 0 keeps one list index and one item in flight ahead of the
// item being processed (the cursor's atomicAdd and the list load are L2 round trips).  The empty tiles
// (k_front's list) are dealt out evenly over the raster items and written by the item's CTA after the item:
// the stores need no answer, so they drain to HBM under the latency-bound raster work.
__global__ void __launch_bounds__(TILE_THREADS, 1024 / TILE_THREADS) k_tile(const FrameUniforms *__restrict__ Up, const SceneDev S,
                                                                            const FrameDev W) {
    const FrameUniforms &U = *Up; // per-frame uniforms, device-resident (one upload per frame; the launches never change)
    uint8_t *__restrict__ color = U.color;
    float *__restrict__ depth = U.depth;
    __shared__ uint32_t s_item, s_index, s_bucket_end[COST_BUCKETS];
    __shared__ float u8tab[256]; // (u8 as f32) / 255.0, filled once per CTA (visible after the loop's first barrier)
    fill_u8_table(u8tab, threadIdx.x, TILE_THREADS);
    // the material table on chip when it fits (64 B each): one dependent global load less per shaded pixel
    __shared__ __align__(16) MaterialDev s_mats[MAT_CACHE];
    const bool mats_cached = S.n_materials <= (uint32_t)MAT_CACHE;
    if (mats_cached) {
        const uint4 *src = reinterpret_cast<const uint4 *>(S.materials);
        uint4 *dst = reinterpret_cast<uint4 *>(s_mats);
        for (uint32_t i = threadIdx.x; i < S.n_materials * 4u; i += TILE_THREADS) dst[i] = __ldg(src + i);
    }
    const MaterialDev *mats = mats_cached ? s_mats : S.materials;
    const CtaTrace trace_(W, 3u);
    const bool usable = W.counters[CNT_OVERFLOW] == 0; // a work buffer overflowed: lists are unusable, the host re-renders
    // The work list is bucketed by cost (heaviest bucket first): bucket b's items are tile_order[b * bucket_cap ..
    // + counters[CNT_BUCKETS + b]).  Item index -> bucket through the running sums of the bucket sizes.
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int b = 0; b < COST_BUCKETS; b++) {
            run += W.counters[CNT_BUCKETS + b];
            s_bucket_end[b] = run;
        }
    }
    __syncthreads();
    const uint32_t n_items = s_bucket_end[COST_BUCKETS - 1];
    // The frame's counters (record / reference totals, overflow flags, statistics, k_front's phase stamps: final
    // since k_front) go to the canvas' pinned host memory as posted stores — no copy-engine transfer that would
    // queue behind the frames' 33-132 MB read-backs, no extra launch.  Visible to the host once the kernel has completed.
    if (blockIdx.x == 0 && threadIdx.x < N_STATUS_WORDS && U.status_host) {
        uint32_t v = __ldcg(&W.counters[threadIdx.x]);
        if (threadIdx.x == CNT_ITEMS) v = n_items;
        if (threadIdx.x == CNT_REFS_NEEDED)
            v = max(max(__ldcg(&W.counters[CNT_L_PAIRS]), __ldcg(&W.counters[CNT_T_PAIRS])), max(__ldcg(&W.counters[CNT_MEDIUM]), __ldcg(&W.counters[CNT_SMALL])));
        U.status_host[threadIdx.x] = v;
        __threadfence_system();
    }
    const uint32_t n_empty = W.counters[CNT_EMPTY];
    const uint32_t per_item = n_items ? (n_empty + n_items - 1u) / n_items : 0u;
    const int W_ = (int)U.canvas_w, H_ = (int)U.canvas_h;
    // The empty tiles are dealt out evenly over the raster items: item k's CTA writes tiles [k * per_item, (k+1) * per_item)
    // — before the item (FrameUniforms::clear_first: the stores drain to HBM under the item's latency-bound work) or after it.
    auto clear_one = [&](uint32_t e) {
        const uint32_t et = __ldg(&W.empty_tiles[e]); // x | y << 10
        clear_tile_cta(et, W_, H_, U.depth_max, color, depth, threadIdx.x, U.empty_tile_color != 0u);
        if (threadIdx.x == 0 && U.tile_state) U.tile_state[(et >> 10) * U.tiles_x + (et & (MAX_TILES_X - 1))] = 0; // clear colour only
    };
    auto clear_share = [&](uint32_t k) {
        for (uint32_t e = k * per_item, e_end = min(n_empty, e + per_item); e < e_end; e++) clear_one(e);
    };
    auto fetch_item = [&](uint32_t i) -> uint32_t { // thread 0 only
        if (i >= n_items) return ITEM_NONE;
        int b = 0;
        while (s_bucket_end[b] <= i) b++;
        return __ldcg(&W.tile_order[(size_t)b * W.bucket_cap + (i - (b ? s_bucket_end[b - 1] : 0u))]);
    };
    // The first item of a CTA is its own index; the cursor (in a cache line of its own: a load that shares
    // a line with a contended atomic queues behind it) hands out the rest.
    uint32_t *cursor = W.counters + CNT_ITEM_CURSOR;
    uint32_t item = ITEM_NONE, index = blockIdx.x, ahead = 0; // thread 0 only
    if (threadIdx.x == 0) {
        ahead = gridDim.x + atomicAdd(cursor, 1u);
        item = fetch_item(blockIdx.x);
    }
    while (true) {
        uint32_t next_item = ITEM_NONE, next_ahead = 0;
        if (threadIdx.x == 0) {
            s_item = item;
            s_index = index;
            next_item = fetch_item(ahead); // consumed after the item below
            next_ahead = gridDim.x + atomicAdd(cursor, 1u);
        }
        __syncthreads();
        const uint32_t cur = s_item, cur_index = s_index;
        if (cur == ITEM_NONE) break;
        if (U.clear_first) clear_share(cur_index);
        tile_item(cur, U, S, W, color, depth, u8tab, mats, usable);
        if (!U.clear_first) clear_share(cur_index);
        __syncthreads(); // shared memory (and s_item) are reused by the next item
        item = next_item;
        index = ahead;
        ahead = next_ahead;
    }
    if (n_items == 0u) // nothing to rasterise in this launch's rows: the CTAs share the empty tiles
        for (uint32_t e = blockIdx.x; e < n_empty; e += gridDim.x) clear_one(e);
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
uint32_t tile_grid_items(const FrameUniforms &U); // k_front.cu
thread_local unsigned g_tile_ctas = 148u * 3u;    // scene.cpp: DRAW_B200_TILE_CTAS (persistent CTAs of k_tile)
void launch_tile(const FrameUniforms &U, const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream) {
    const uint32_t slots = tile_grid_items(U); // upper bound of the work items
    if (!slots) return;
    k_tile<<<min(slots, g_tile_ctas), TILE_THREADS, 0, stream>>>(dU, S, W);
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
