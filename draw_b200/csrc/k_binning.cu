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

template <bool FILL>
__device__ __forceinline__ void bin_hit(const FrameDev &W, uint32_t list, uint32_t slot) {
    const uint32_t pos = atomicAdd(&W.list_count[list], 1u);
    if (FILL) W.list_refs[W.list_offset[list] + pos] = slot;
}

// Reference one tile from a record if the triangle can cover a pixel of it; the list class follows
// the area of the bbox clipped to the tile.
template <bool FILL>
__device__ __forceinline__ void bin_tile(const FrameUniforms &U, const FrameDev &W, const TriEdges &t, int x0, int x1,
                                         int y0, int y1, int tx, int ty, uint32_t slot) {
    const int lx = max(x0, tx * TILE_W), hx = min(x1, tx * TILE_W + TILE_W - 1);
    const int ly = max(y0, ty * TILE_H), hy = min(y1, ty * TILE_H + TILE_H - 1);
    if (!rect_may_cover(t, lx, hx, ly, hy)) return;
    const int area = (hx - lx + 1) * (hy - ly + 1);
    const uint32_t cls = area <= SMALL_AREA ? 2u : (area <= MEDIUM_AREA ? 1u : 0u);
    bin_hit<FILL>(W, cls * U.n_coarse + (uint32_t)ty * U.tiles_x + (uint32_t)tx, slot);
}

constexpr int BIN_THREADS = 256;
constexpr int WIDE_TILES = 8;        // thread-per-record mode: records covering more tiles are binned by the whole warp
constexpr int RECORDS_PER_WARP = 8;  // up to this many records per warp of the grid: one warp per record

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
__global__ void __launch_bounds__(BIN_THREADS) k_bin(const __grid_constant__ FrameUniforms U, const FrameDev W) {
    pdl_prologue();
    if (FILL && W.counters[2] != 0) return; // a buffer overflowed: the host re-renders with larger buffers
    uint32_t n = W.counters[0];
    if (n > W.rec_cap) n = W.rec_cap;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_warps = gridDim.x * (BIN_THREADS / 32);

    if (n <= n_warps * RECORDS_PER_WARP) {
        // Few records (they may each cover many tiles): one warp per record, lanes stride over its
        // tiles so that the atomics of one record are in flight together.
        for (uint32_t slot = blockIdx.x * (BIN_THREADS / 32) + (threadIdx.x >> 5); slot < n; slot += n_warps) {
            const RasterRec r = load_raster(W.rrec + slot); // same address in every lane: one broadcast load
            const TileRange tr = tile_range(U, r.bbx, r.bby);
            const int total = tr.count();
            if (total == 0 || r.id == NO_SLOT) continue; // NO_SLOT: reserved by k_setup, not used by k_clip
            const TriEdges t = prepare_edges(r);
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

        if (valid && !wide) {
            const TriEdges t = prepare_edges(r);
            for (int ty = tr.ty0; ty <= tr.ty1; ty++)
                for (int tx = tr.tx0; tx <= tr.tx1; tx++) bin_tile<FILL>(U, W, t, tr.x0, tr.x1, tr.y0, tr.y1, tx, ty, slot);
        }

        uint32_t pending = __ballot_sync(0xFFFFFFFFu, wide);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            RasterRec w;
            w.ax = __shfl_sync(0xFFFFFFFFu, r.ax, src); w.ay = __shfl_sync(0xFFFFFFFFu, r.ay, src);
            w.bx = __shfl_sync(0xFFFFFFFFu, r.bx, src); w.by = __shfl_sync(0xFFFFFFFFu, r.by, src);
            w.cx = __shfl_sync(0xFFFFFFFFu, r.cx, src); w.cy = __shfl_sync(0xFFFFFFFFu, r.cy, src);
            const TileRange wt = tile_range(U, __shfl_sync(0xFFFFFFFFu, r.bbx, src), __shfl_sync(0xFFFFFFFFu, r.bby, src));
            const uint32_t wslot = __shfl_sync(0xFFFFFFFFu, slot, src);
            const TriEdges t = prepare_edges(w);
            const int cols = wt.tx1 - wt.tx0 + 1, total = wt.count();
            for (int i = (int)lane; i < total; i += 32)
                bin_tile<FILL>(U, W, t, wt.x0, wt.x1, wt.y0, wt.y1, wt.tx0 + i % cols, wt.ty0 + i / cols, wslot);
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_alloc : one thread per tile.
//  (1) gives the tile's three lists a contiguous range of list_refs.  Lists need not be laid out in
//      tile order, so instead of a global scan each CTA scans its 256 sums locally and reserves its
//      total with one atomicAdd; list_count is reset to serve as the fill cursor (after k_bin<true> it
//      holds the count again, which is what k_tile reads).  total -> counters[1].
//  (2) orders the stripe's tiles by estimated work, heaviest first, so that k_tile (whose CTAs are
//      dispatched in index order) starts its long tiles first and ends with the empty ones: the last
//      CTA to finish (1) bucket-sorts the per-tile costs by their log2.
// ------------------------------------------------------------------------------------------
constexpr int ALLOC_THREADS = 256;
constexpr uint32_t COST_NOT_IN_STRIPE = 0xFFFFFFFFu;
constexpr int COST_BUCKETS = 34;

__device__ __forceinline__ int cost_bucket(uint32_t cost) { // 0 = heaviest ... COST_BUCKETS-1 = empty tile
    return cost == 0 ? COST_BUCKETS - 1 : __clz(cost);      // clz in 0..31 (larger cost -> smaller clz)
}

__global__ void __launch_bounds__(ALLOC_THREADS) k_alloc(const __grid_constant__ FrameUniforms U, const FrameDev W) {
    __shared__ uint32_t warp_sum[ALLOC_THREADS / 32];
    __shared__ uint32_t block_base, is_last;
    __shared__ uint32_t bucket_start[COST_BUCKETS];
    pdl_prologue();
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x * ALLOC_THREADS + tid, nc = U.n_coarse;
    const bool valid = tile < nc;
    const uint32_t c0 = valid ? W.list_count[tile] : 0u, c1 = valid ? W.list_count[nc + tile] : 0u,
                   c2 = valid ? W.list_count[2 * nc + tile] : 0u;
    const uint32_t c = c0 + c1 + c2;

    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (uint32_t)d) incl += up;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (tid == 0) {
        uint32_t total = 0;
        for (int w = 0; w < ALLOC_THREADS / 32; w++) {
            const uint32_t t = warp_sum[w];
            warp_sum[w] = total;
            total += t;
        }
        const uint32_t base = total ? atomicAdd(&W.counters[1], total) : 0u;
        if (total && base + total > W.refs_cap) atomicOr(&W.counters[2], OVERFLOW_REFS);
        block_base = base;
    }
    __syncthreads();
    if (valid) {
        const uint32_t off = block_base + warp_sum[warp] + incl - c;
        W.list_offset[tile] = off;
        W.list_offset[nc + tile] = off + c0;
        W.list_offset[2 * nc + tile] = off + c0 + c1;
        W.list_count[tile] = 0; // become the fill cursors
        W.list_count[nc + tile] = 0;
        W.list_count[2 * nc + tile] = 0;
        const uint32_t ty = tile / U.tiles_x;
        // rough relative cost of a large / medium / small reference in k_tile
        W.tile_cost[tile] = (ty >= U.tile_y_begin && ty < U.tile_y_end) ? 16u * c0 + 4u * c1 + c2 : COST_NOT_IN_STRIPE;
    }
    // ---- the last CTA orders the tiles ---------------------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(&W.counters[4], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    for (int b = (int)tid; b < COST_BUCKETS; b += ALLOC_THREADS) bucket_start[b] = 0;
    __syncthreads();
    constexpr int BATCH = 8; // independent loads in flight per thread
    for (uint32_t t0 = tid; t0 < nc; t0 += ALLOC_THREADS * BATCH) {
        uint32_t cost[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            const uint32_t t = t0 + k * ALLOC_THREADS;
            cost[k] = t < nc ? __ldcg(&W.tile_cost[t]) : COST_NOT_IN_STRIPE;
        }
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (cost[k] != COST_NOT_IN_STRIPE) atomicAdd(&bucket_start[cost_bucket(cost[k])], 1u);
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int b = 0; b < COST_BUCKETS; b++) {
            const uint32_t n = bucket_start[b];
            bucket_start[b] = run;
            run += n;
        }
    }
    __syncthreads();
    for (uint32_t t0 = tid; t0 < nc; t0 += ALLOC_THREADS * BATCH) {
        uint32_t cost[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            const uint32_t t = t0 + k * ALLOC_THREADS;
            cost[k] = t < nc ? __ldcg(&W.tile_cost[t]) : COST_NOT_IN_STRIPE;
        }
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (cost[k] != COST_NOT_IN_STRIPE)
                W.tile_order[atomicAdd(&bucket_start[cost_bucket(cost[k])], 1u)] = t0 + k * ALLOC_THREADS;
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static int bin_blocks(const FrameDev &) { return 148 * 4; }
void launch_bin_count(const FrameUniforms &U, const FrameDev &W, cudaStream_t stream) {
    launch_pdl(k_bin<false>, bin_blocks(W), BIN_THREADS, stream, U, W);
}
void launch_alloc(const FrameUniforms &U, const FrameDev &W, cudaStream_t stream) {
    launch_pdl(k_alloc, (U.n_coarse + ALLOC_THREADS - 1) / ALLOC_THREADS, ALLOC_THREADS, stream, U, W);
}
void launch_bin_fill(const FrameUniforms &U, const FrameDev &W, cudaStream_t stream) {
    launch_pdl(k_bin<true>, bin_blocks(W), BIN_THREADS, stream, U, W);
}

} // namespace drawb200
