// k_binning.cu — sorts the frame's raster records into screen-tile lists.
//
// The reference has no such stage (it loops each triangle's bbox directly, canvas.rs:668-670);
// it exists so that k_tile can keep a tile's colour and depth on chip and write HBM once.
//
//   k_bin<false>  per record  count the lists it belongs to
//   k_alloc       per tile    reserve a contiguous range of list_refs for every list; order the tiles
//                             heaviest first for k_tile
//   k_bin<true>   per record  scatter its slot into those lists
//
// Every tile (64x32 px) has three lists (device_types.h), classed by the area of the record's bbox
// inside the tile: "large" records are rasterised by k_tile with every lane testing its own pixels,
// "medium" ones one record per warp, "small" ones one record per lane.
// A tile is only referenced if the triangle can actually cover a pixel in it (exact corner
// test, rect_may_cover), not merely because its bbox touches it.  Records covering many tiles are
// binned by the whole warp (ballot picks them, lanes stride over the tiles).
// Order inside a list is irrelevant: k_tile resolves fragments by (depth, record slot), and slots
// are allocated in draw order.
#include "device_math.cuh"

namespace drawb200 {

// Each class has its own reference array (a tile's list is a contiguous range of it, and the ranges of
// all tiles tile the array without gaps, which is what lets k_raster walk the medium and small
// references of the whole frame as one flat list).  Medium and small references carry their tile.
template <bool FILL>
__device__ __forceinline__ void bin_hit(const FrameDev &W, uint32_t cls, uint32_t list, uint32_t slot, uint32_t tile_xy) {
    const uint32_t pos = atomicAdd(&W.list_count[list], 1u);
    if (FILL) {
        const uint32_t at = W.list_offset[list] + pos;
        if (cls == 0u) W.list_refs[at] = slot;
        else (cls == 1u ? W.m_refs : W.s_refs)[at] = make_uint2(slot, tile_xy);
    }
}
// Estimated k_tile work of one reference, in quarter block-iterations (one iteration = a warp testing an
// 8x4 block of a medium triangle, ~130 instructions): a large triangle costs every warp of the CTA a
// pass, a medium one its 8x4 blocks, a small one a lane.  Summed per tile by k_bin<count> for k_alloc.
constexpr uint32_t COST_LARGE = 80u, COST_MEDIUM_BLOCK = 4u, COST_SMALL = 1u, COST_TILE_BASE = 16u;

// Reference one tile from a record if the triangle can cover a pixel of it; the list class follows
// the area of the bbox clipped to the tile.
template <bool FILL, typename Tri>
__device__ __forceinline__ void bin_tile(const FrameUniforms &U, const FrameDev &W, const Tri &t, int x0, int x1,
                                         int y0, int y1, int tx, int ty, uint32_t slot) {
    const int lx = max(x0, tx * TILE_W), hx = min(x1, tx * TILE_W + TILE_W - 1);
    const int ly = max(y0, ty * TILE_H), hy = min(y1, ty * TILE_H + TILE_H - 1);
    if (!rect_may_cover(t, (float)lx, (float)hx, (float)ly, (float)hy)) return;
    const int area = (hx - lx + 1) * (hy - ly + 1);
    const uint32_t cls = area <= SMALL_AREA ? 2u : (area <= MEDIUM_AREA ? 1u : 0u);
    const uint32_t tile = (uint32_t)ty * U.tiles_x + (uint32_t)tx;
    const uint32_t blocks = (uint32_t)((hx - lx) / 8 + 1) * (uint32_t)((hy - ly) / 4 + 1);
    // A medium reference with many 8x4 blocks is entered 2-4 times, each entry naming a share of the blocks
    // (ref_part bits), so that k_raster's warp-per-reference jobs stay short; k_tile skips the extra entries.
    const uint32_t parts = cls == 1u ? min(4u, (blocks + 7u) / 8u) : 1u;
    for (uint32_t part = 0; part < parts; part++)
        bin_hit<FILL>(W, cls, cls * U.n_coarse + tile, slot, (uint32_t)tx | (uint32_t)ty << 10 | part << 21 | (parts - 1u) << 23);
    if (!FILL) { // [tile]: cost of the large references, [n_coarse + tile]: of the medium and small ones
        if (cls == 0u) atomicAdd(&W.tile_cost[tile], COST_LARGE);
        else atomicAdd(&W.tile_cost[U.n_coarse + tile], cls == 2u ? COST_SMALL : COST_MEDIUM_BLOCK * blocks);
    }
}

// The part of a prepared record binning needs.  k_bin<count> prepares every record once (make_prep)
// and writes the PrepRec k_tile stages from; k_bin<fill> reads the edges back instead of redoing them.
struct BinTri {
    float ecx[3], ecy[3], ek1[3], ek2[3];
    uint32_t flags;
};
template <bool FILL>
__device__ __forceinline__ BinTri bin_prepare(const FrameDev &W, const RasterRec &r, uint32_t slot, bool store) {
    BinTri b;
    if (!FILL) {
        PrepRec p;
        make_prep(r, p);
        if (store) store_prep(W.prep + slot, p);
#pragma unroll
        for (int i = 0; i < 3; i++) { b.ecx[i] = p.ecx[i]; b.ecy[i] = p.ecy[i]; b.ek1[i] = p.ek1[i]; b.ek2[i] = p.ek2[i]; }
        b.flags = p.flags;
    } else {
        const uint4 *q = reinterpret_cast<const uint4 *>(W.prep + slot);
        const uint4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q6 = __ldg(q + 6);
        b.ecx[0] = __uint_as_float(q0.x); b.ecx[1] = __uint_as_float(q0.y); b.ecx[2] = __uint_as_float(q0.z);
        b.ecy[0] = __uint_as_float(q0.w); b.ecy[1] = __uint_as_float(q1.x); b.ecy[2] = __uint_as_float(q1.y);
        b.ek1[0] = __uint_as_float(q1.z); b.ek1[1] = __uint_as_float(q1.w); b.ek1[2] = __uint_as_float(q2.x);
        b.ek2[0] = __uint_as_float(q2.y); b.ek2[1] = __uint_as_float(q2.z); b.ek2[2] = __uint_as_float(q2.w);
        b.flags = q6.y;
    }
    return b;
}

constexpr int BIN_THREADS = 256;
constexpr int WIDE_TILES = 8;        // thread-per-record mode: records covering more tiles are binned by the whole warp
// (FrameUniforms::bin_records_per_warp: up to this many records per warp of the grid, one warp per record)

// Tile range of a record's bbox, clipped to this launch's stripe.
struct TileRange {
    int x0, x1, y0, y1;     // bbox in pixels
    int tx0, tx1, ty0, ty1; // tiles
    __device__ __forceinline__ int count() const { return ty0 <= ty1 ? (tx1 - tx0 + 1) * (ty1 - ty0 + 1) : 0; }
};
__device__ __forceinline__ TileRange tile_range(const FrameUniforms &U, uint32_t bbx, uint32_t bby) {
    TileRange t;
    t.x0 = (int)(bbx & 0xFFFF); t.x1 = (int)(bbx >> 16);
    t.y0 = (int)(bby & 0xFFFF); t.y1 = (int)(bby >> 16);
    t.tx0 = t.x0 / TILE_W; t.tx1 = t.x1 / TILE_W;
    t.ty0 = max(t.y0 / TILE_H, (int)U.tile_y_begin);
    t.ty1 = min(t.y1 / TILE_H, (int)U.tile_y_end - 1);
    return t;
}

template <bool FILL>
__global__ void __launch_bounds__(BIN_THREADS) k_bin(const FrameUniforms *__restrict__ Up, const FrameDev W) {
    const FrameUniforms &U = *Up; // per-frame uniforms, device-resident (one upload per frame; the launches never change)
    const CtaTrace trace_(W, (FILL ? 5u : 3u));
    pdl_prologue(U.pdl_early != 0);
    if (FILL && W.counters[2] != 0) return; // a buffer overflowed: the host re-renders with larger buffers
    uint32_t n = W.counters[0];
    if (n > W.rec_cap) n = W.rec_cap;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_warps = gridDim.x * (BIN_THREADS / 32);

    if (n <= n_warps * U.bin_records_per_warp) {
        // Few records (they may each cover many tiles): one warp per record, lanes stride over its
        // tiles so that the atomics of one record are in flight together.
        for (uint32_t slot = blockIdx.x * (BIN_THREADS / 32) + (threadIdx.x >> 5); slot < n; slot += n_warps) {
            const RasterRec r = load_raster(W.rrec + slot); // same address in every lane: one broadcast load
            const TileRange tr = tile_range(U, r.bbx, r.bby);
            const int total = tr.count();
            if (r.id == NO_SLOT) continue; // reserved by k_setup, not used by k_clip
            if (total == 0) {              // outside the stripe: never referenced, but keep the PrepRec defined
                if (!FILL && lane == 0) bin_prepare<FILL>(W, r, slot, true);
                continue;
            }
            const BinTri t = bin_prepare<FILL>(W, r, slot, lane == 0);
            const int cols = tr.tx1 - tr.tx0 + 1;
            for (int i = (int)lane; i < total; i += 32)
                bin_tile<FILL>(U, W, t, tr.x0, tr.x1, tr.y0, tr.y1, tr.tx0 + i % cols, tr.ty0 + i / cols, slot);
        }
        return;
    }

    // Many records: one thread per record; the rare record covering many tiles is handed to the warp.
    for (uint32_t base = blockIdx.x * BIN_THREADS; base < n; base += gridDim.x * BIN_THREADS) {
        const uint32_t slot = base + threadIdx.x;
        bool valid = slot < n;
        RasterRec r;
        if (valid) r = load_raster(W.rrec + slot);
        else { r.id = NO_SLOT; r.bbx = r.bby = 0; r.ax = r.ay = r.bx = r.by = r.cx = r.cy = 0.0f; }
        if (r.id == NO_SLOT) valid = false; // also: slots reserved by k_setup that k_clip did not use
        const TileRange tr = tile_range(U, r.bbx, r.bby);
        const bool wide = valid && tr.count() > WIDE_TILES;
        BinTri t = {};
        if (valid) t = bin_prepare<FILL>(W, r, slot, true);

        if (valid && !wide) {
            for (int ty = tr.ty0; ty <= tr.ty1; ty++)
                for (int tx = tr.tx0; tx <= tr.tx1; tx++) bin_tile<FILL>(U, W, t, tr.x0, tr.x1, tr.y0, tr.y1, tx, ty, slot);
        }

        uint32_t pending = __ballot_sync(0xFFFFFFFFu, wide);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            BinTri w;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                w.ecx[i] = __shfl_sync(0xFFFFFFFFu, t.ecx[i], src); w.ecy[i] = __shfl_sync(0xFFFFFFFFu, t.ecy[i], src);
                w.ek1[i] = __shfl_sync(0xFFFFFFFFu, t.ek1[i], src); w.ek2[i] = __shfl_sync(0xFFFFFFFFu, t.ek2[i], src);
            }
            w.flags = __shfl_sync(0xFFFFFFFFu, t.flags, src);
            const TileRange wt = tile_range(U, __shfl_sync(0xFFFFFFFFu, r.bbx, src), __shfl_sync(0xFFFFFFFFu, r.bby, src));
            const uint32_t wslot = __shfl_sync(0xFFFFFFFFu, slot, src);
            const int cols = wt.tx1 - wt.tx0 + 1, total = wt.count();
            for (int i = (int)lane; i < total; i += 32)
                bin_tile<FILL>(U, W, w, wt.x0, wt.x1, wt.y0, wt.y1, wt.tx0 + i % cols, wt.ty0 + i / cols, wslot);
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_alloc : one thread per tile.
//  (1) gives the tile's three lists a contiguous range of list_refs.  Lists need not be laid out in
//      tile order, so instead of a global scan each CTA scans its 256 sums locally and reserves its
//      total with one atomicAdd; list_count is reset to serve as the fill cursor (after k_bin<true> it
//      holds the count again, which is what k_tile reads).  total -> counters[1].
//  (2) builds k_tile's work list (the last CTA to finish (1) does it).  A work item is a tile or, for
//      a tile whose estimated cost is far above the frame's average, one of 2/4/8/16 pixel windows of
//      it (device_types.h: make_item): every window gets its own 512-thread CTA, walks the tile's
//      lists and keeps only what falls inside its window, so that a few dense tiles do not decide
//      the duration of the whole launch.  Items are bucket-sorted by the log2 of their cost, heaviest
//      first (k_tile's CTAs are dispatched in index order); empty tiles carry a flag so that their
//      CTA needs no further loads.  Unused slots up to the launch's grid size hold ITEM_NONE.
// ------------------------------------------------------------------------------------------
constexpr int ALLOC_THREADS = 256;
constexpr uint32_t COST_NOT_IN_STRIPE = 0xFFFFFFFFu;
// tile_cost after k_alloc's first part: cost in the low 28 bits, what k_tile does with the tile above them
constexpr uint32_t TILE_KIND_SHIFT = 28, TILE_KIND_SHIFTED = 1u << TILE_KIND_SHIFT;
constexpr uint32_t TILE_KIND_FULL = 0, TILE_KIND_DEFER = 1, TILE_KIND_RASTER_ONLY = 2, TILE_KIND_EMPTY = 3;
constexpr int COST_BUCKETS = 34;

__device__ __forceinline__ int cost_bucket(uint32_t cost) { // 0 = heaviest ... COST_BUCKETS-1 = empty tile
    return cost == 0 ? COST_BUCKETS - 1 : __clz(cost);      // clz in 0..31 (larger cost -> smaller clz)
}
// Number of windows a tile of this cost is cut into: the largest power of two <= cost / target,
// so that the sum over tiles stays <= total cost / target <= TILE_EXTRA_ITEMS.
__device__ __forceinline__ uint32_t tile_splits(uint32_t cost, uint32_t target, uint32_t max_split) {
    if (!TILE_SPLITTABLE || cost < 2u * target) return 1u;
    const uint32_t q = cost / target;
    return min(max_split, 1u << (31 - __clz(q)));
}

__global__ void __launch_bounds__(ALLOC_THREADS) k_alloc(const FrameUniforms *__restrict__ Up, const FrameDev W) {
    const FrameUniforms &U = *Up; // per-frame uniforms, device-resident (one upload per frame; the launches never change)
    __shared__ uint32_t warp_sum[3][ALLOC_THREADS / 32], warp_cost[ALLOC_THREADS / 32];
    __shared__ uint32_t block_base3[3], block_base, is_last;
    __shared__ uint32_t bucket_start[COST_BUCKETS];
    const CtaTrace trace_(W, 4u);
    pdl_prologue(U.pdl_early != 0);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x * ALLOC_THREADS + tid, nc = U.n_coarse;
    const bool valid = tile < nc;
    const uint32_t c0 = valid ? W.list_count[tile] : 0u, c1 = valid ? W.list_count[nc + tile] : 0u,
                   c2 = valid ? W.list_count[2 * nc + tile] : 0u;
    const uint32_t c = c0 + c1 + c2;
    const uint32_t ty = valid ? tile / U.tiles_x : 0u;
    const bool in_stripe = valid && ty >= U.tile_y_begin && ty < U.tile_y_end;
    // A tile with medium or small references gets a key page if one is left: k_raster then rasterises those
    // references into the page for the whole frame at once and k_tile only merges the page (device_types.h).
    uint32_t page = NO_PAGE;
    if (in_stripe && c1 + c2 > 0 && W.page_cap) {
        page = atomicAdd(&W.counters[11], 1u);
        if (page >= W.page_cap) page = NO_PAGE;
    }
    // What k_tile has to do for the tile (TILE_KIND_*), and its estimated cost, summed per class by
    // k_bin<count> (COST_* above).  A tile with a key page and few large triangles is resolved by k_shade,
    // one thread per pixel: k_tile then only adds its large triangles to the page (nothing at all if there
    // are none), which keeps the serial per-tile work — and the longest item — short.
    uint32_t cost = 0, kind = TILE_KIND_EMPTY;
    if (in_stripe && c > 0) {
        const bool defer = page != NO_PAGE && U.has_transparent == 0 && c0 < U.defer_max;
        kind = defer ? (c0 == 0 ? TILE_KIND_RASTER_ONLY : TILE_KIND_DEFER) : TILE_KIND_FULL;
        if (kind != TILE_KIND_RASTER_ONLY) {
            // shading: cost_shade units for a tile's worth of covered pixels; a tile with large triangles is taken
            // as covered, the others by the bbox area of their medium / small references (COST_MEDIUM_BLOCK per 32 px)
            const uint32_t cm = W.tile_cost[nc + tile];
            const uint32_t full = (uint32_t)(TILE_W * TILE_H / 32) * COST_MEDIUM_BLOCK;
            const uint32_t shade = c0 ? U.cost_shade : (uint32_t)((unsigned long long)min(cm, full) * U.cost_shade / full);
            cost = COST_TILE_BASE + W.tile_cost[tile] + (page == NO_PAGE ? cm : 0u) + shade;
        }
        if (cost >= TILE_KIND_SHIFTED) cost = TILE_KIND_SHIFTED - 1u;
        if (defer) W.shade_tiles[atomicAdd(&W.counters[14], 1u)] = (tile % U.tiles_x) | ty << 10;
    }

    // exclusive prefixes of the three counts over the CTA, one range reservation per class
    uint32_t inc0 = c0, inc1 = c1, inc2 = c2, cost_sum = cost;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u0 = __shfl_up_sync(0xFFFFFFFFu, inc0, d), u1 = __shfl_up_sync(0xFFFFFFFFu, inc1, d),
                       u2 = __shfl_up_sync(0xFFFFFFFFu, inc2, d);
        if (lane >= (uint32_t)d) { inc0 += u0; inc1 += u1; inc2 += u2; }
        cost_sum += __shfl_xor_sync(0xFFFFFFFFu, cost_sum, d);
    }
    if (lane == 31) { warp_sum[0][warp] = inc0; warp_sum[1][warp] = inc1; warp_sum[2][warp] = inc2; }
    if (lane == 0) warp_cost[warp] = cost_sum;
    __syncthreads();
    if (tid < 3) {
        uint32_t total = 0;
        for (int w = 0; w < ALLOC_THREADS / 32; w++) {
            const uint32_t t = warp_sum[tid][w];
            warp_sum[tid][w] = total;
            total += t;
        }
        const uint32_t base = total ? atomicAdd(&W.counters[8 + tid], total) : 0u;
        if (total) atomicAdd(&W.counters[1], total); // all classes: frame statistics, buffer growth
        if (total && base + total > W.refs_cap) atomicOr(&W.counters[2], OVERFLOW_REFS);
        block_base3[tid] = base;
    }
    if (tid == 32) {
        uint32_t total_cost = 0;
        for (int w = 0; w < ALLOC_THREADS / 32; w++) total_cost += warp_cost[w];
        if (total_cost) atomicAdd(&W.counters[6], total_cost);
    }
    __syncthreads();
    if (valid) {
        W.list_offset[tile] = block_base3[0] + warp_sum[0][warp] + inc0 - c0;
        W.list_offset[nc + tile] = block_base3[1] + warp_sum[1][warp] + inc1 - c1;
        W.list_offset[2 * nc + tile] = block_base3[2] + warp_sum[2][warp] + inc2 - c2;
        W.list_count[tile] = 0; // become the fill cursors
        W.list_count[nc + tile] = 0;
        W.list_count[2 * nc + tile] = 0;
        W.tile_cost[tile] = in_stripe ? (cost | kind << TILE_KIND_SHIFT) : COST_NOT_IN_STRIPE;
        W.tile_page[tile] = page;
    }
    // ---- the last CTA builds the work list -------------------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(&W.counters[4], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __shared__ uint32_t n_empty;
    for (int b = (int)tid; b < COST_BUCKETS; b += ALLOC_THREADS) bucket_start[b] = 0;
    if (tid == 0) n_empty = 0;
    __syncthreads();
    const uint32_t total_cost = __ldcg(&W.counters[6]);
    const uint32_t target = max(U.split_min_cost, total_cost / U.split_div + 1u); // split_div <= TILE_EXTRA_ITEMS
    const uint32_t max_split = U.split_max;
    const bool list_empties = U.has_transparent == 0; // else every tile runs the full path (cost 0, last bucket)
    constexpr int BATCH = 8; // independent loads in flight per thread
    for (uint32_t base = 0; base < nc; base += ALLOC_THREADS * BATCH) { // warp-uniform trip count: ballots below
        uint32_t cost_k[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            const uint32_t t = base + tid + k * ALLOC_THREADS;
            cost_k[k] = t < nc ? __ldcg(&W.tile_cost[t]) : COST_NOT_IN_STRIPE;
        }
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            const uint32_t t = base + tid + k * ALLOC_THREADS;
            const bool in = cost_k[k] != COST_NOT_IN_STRIPE;
            const uint32_t kind = cost_k[k] >> TILE_KIND_SHIFT, cost = cost_k[k] & (TILE_KIND_SHIFTED - 1u);
            // empty tiles are listed on their own (k_clear_empty writes them): most of the frame's tiles, so one
            // shared-memory atomic per warp, not per tile
            const bool empty = in && kind == TILE_KIND_EMPTY && list_empties;
            const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, empty);
            if (ballot) {
                uint32_t wbase = 0;
                if (lane == 0) wbase = atomicAdd(&n_empty, (uint32_t)__popc(ballot));
                wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
                if (empty) W.empty_tiles[wbase + __popc(ballot & ((1u << lane) - 1u))] = (t % U.tiles_x) | (t / U.tiles_x) << 10;
            }
            if (!in || empty || kind == TILE_KIND_RASTER_ONLY) continue; // raster-only: k_raster + k_shade do it all
            const uint32_t s = tile_splits(cost, target, max_split);
            atomicAdd(&bucket_start[cost_bucket(cost / s)], s);
        }
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int b = 0; b < COST_BUCKETS; b++) {
            const uint32_t n = bucket_start[b];
            bucket_start[b] = run;
            run += n;
        }
        block_base = run; // number of tile / window items
    }
    __syncthreads();
    const uint32_t n_dense = block_base;
    const uint32_t grid_items = (U.tile_y_end - U.tile_y_begin) * U.tiles_x + (TILE_SPLITTABLE ? TILE_EXTRA_ITEMS : 0);
    for (uint32_t i = n_dense + tid; i < grid_items; i += ALLOC_THREADS) W.tile_order[i] = ITEM_NONE;
    if (tid == 0) {
        W.counters[13] = n_empty; // k_clear_empty's work (or k_tile's, FrameUniforms::clear_in_tile)
        W.counters[15] = n_dense; // k_tile's items
    }
    for (uint32_t t0 = tid; t0 < nc; t0 += ALLOC_THREADS * BATCH) {
        uint32_t cost_k[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            const uint32_t t = t0 + k * ALLOC_THREADS;
            cost_k[k] = t < nc ? __ldcg(&W.tile_cost[t]) : COST_NOT_IN_STRIPE;
        }
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (cost_k[k] != COST_NOT_IN_STRIPE) {
                const uint32_t kind = cost_k[k] >> TILE_KIND_SHIFT, cost = cost_k[k] & (TILE_KIND_SHIFTED - 1u);
                if ((kind == TILE_KIND_EMPTY && list_empties) || kind == TILE_KIND_RASTER_ONLY) continue;
                const uint32_t t = t0 + k * ALLOC_THREADS, s = tile_splits(cost, target, max_split);
                const uint32_t tx = t % U.tiles_x, tyy = t / U.tiles_x;
                const uint32_t at = atomicAdd(&bucket_start[cost_bucket(cost / s)], s);
                for (uint32_t i = 0; i < s; i++)
                    W.tile_order[at + i] = make_item(tx, tyy, s, i) | (kind == TILE_KIND_DEFER ? ITEM_DEFER : 0u);
            }
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
thread_local unsigned g_bin_ctas = 148u * 4u; // scene.cpp: DRAW_B200_BIN_CTAS
static int bin_blocks(const FrameDev &) { return (int)g_bin_ctas; }
void launch_bin_count(const FrameUniforms &U, const FrameUniforms *dU, const FrameDev &W, cudaStream_t stream) {
    launch_pdl(k_bin<false>, bin_blocks(W), BIN_THREADS, stream, dU, W);
}
uint32_t tile_grid_items(const FrameUniforms &U) { // k_tile's grid: one CTA per work-list slot
    const uint32_t stripe_tiles = (U.tile_y_end - U.tile_y_begin) * U.tiles_x;
    return stripe_tiles ? stripe_tiles + (TILE_SPLITTABLE ? TILE_EXTRA_ITEMS : 0) : 0;
}
void launch_alloc(const FrameUniforms &U, const FrameUniforms *dU, const FrameDev &W, cudaStream_t stream) {
    launch_pdl(k_alloc, (U.n_coarse + ALLOC_THREADS - 1) / ALLOC_THREADS, ALLOC_THREADS, stream, dU, W);
}
void launch_bin_fill(const FrameUniforms &U, const FrameUniforms *dU, const FrameDev &W, cudaStream_t stream) {
    launch_pdl(k_bin<true>, bin_blocks(W), BIN_THREADS, stream, dU, W);
}

} // namespace drawb200
