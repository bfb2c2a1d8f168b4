// k_binning.cu — sorts the frame's raster records into screen-tile lists.
//
// The reference has no such stage (it loops each triangle's bbox directly, canvas.rs:668-670);
// it exists so that k_tile can keep a tile's colour and depth on chip and write HBM once.
//
//   k_bin<false>  per record  count the lists it belongs to
//   k_alloc       per list    reserve a contiguous range of list_refs for every list
//   k_bin<true>   per record  scatter its slot into those lists
//
// Every tile (64x32 px) has two lists (device_types.h): "large" records are rasterised by k_tile
// with every lane testing its own pixels, "small" records (bbox <= SMALL_AREA px) one record per
// lane.  A tile is only referenced if the triangle can actually cover a pixel in it (exact corner
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

constexpr int BIN_THREADS = 256;
constexpr int WIDE_TILES = 16; // records covering more coarse tiles than this are binned warp-cooperatively

template <bool FILL>
__global__ void __launch_bounds__(BIN_THREADS) k_bin(const __grid_constant__ FrameUniforms U, const FrameDev W) {
    if (FILL && W.counters[2] != 0) return; // a buffer overflowed: the host re-renders with larger buffers
    uint32_t n = W.counters[0];
    if (n > W.rec_cap) n = W.rec_cap;
    const uint32_t lane = threadIdx.x & 31;

    for (uint32_t base = blockIdx.x * BIN_THREADS; base < n; base += gridDim.x * BIN_THREADS) {
        const uint32_t slot = base + threadIdx.x;
        const bool valid = slot < n;
        RasterRec r;
        if (valid) r = load_raster(W.rrec + slot);
        else { r.bbx = r.bby = 0; r.ax = r.ay = r.bx = r.by = r.cx = r.cy = 0.0f; }
        const int x0 = (int)(r.bbx & 0xFFFF), x1 = (int)(r.bbx >> 16);
        const int y0 = (int)(r.bby & 0xFFFF), y1 = (int)(r.bby >> 16);
        // small: bbox of at most SMALL_AREA pixels -> the tile's "small" list (one lane of k_tile walks it);
        // anything else -> the tile's "large" list (all lanes of k_tile test their own pixels against it)
        const bool small = (x1 - x0 + 1) * (y1 - y0 + 1) <= SMALL_AREA;
        const int tx0 = x0 / TILE_W, tx1 = x1 / TILE_W;
        int ty0 = y0 / TILE_H, ty1 = y1 / TILE_H;
        if (ty0 < (int)U.tile_y_begin) ty0 = (int)U.tile_y_begin;
        if (ty1 > (int)U.tile_y_end - 1) ty1 = (int)U.tile_y_end - 1;
        const int n_coarse_hit = ty0 <= ty1 ? (tx1 - tx0 + 1) * (ty1 - ty0 + 1) : 0;
        const bool wide = valid && n_coarse_hit > WIDE_TILES;

        if (valid && !wide) {
            const TriEdges t = prepare_edges(r);
            const uint32_t list_base = small ? U.n_coarse : 0u;
            for (int ty = ty0; ty <= ty1; ty++)
                for (int tx = tx0; tx <= tx1; tx++) {
                    const int lx = max(x0, tx * TILE_W), hx = min(x1, tx * TILE_W + TILE_W - 1);
                    const int ly = max(y0, ty * TILE_H), hy = min(y1, ty * TILE_H + TILE_H - 1);
                    if (rect_may_cover(t, lx, hx, ly, hy))
                        bin_hit<FILL>(W, list_base + (uint32_t)ty * U.tiles_x + (uint32_t)tx, slot);
                }
        }

        // wide records: one at a time, all 32 lanes stride over its coarse tiles
        uint32_t pending = __ballot_sync(0xFFFFFFFFu, wide);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            RasterRec w;
            w.ax = __shfl_sync(0xFFFFFFFFu, r.ax, src); w.ay = __shfl_sync(0xFFFFFFFFu, r.ay, src);
            w.bx = __shfl_sync(0xFFFFFFFFu, r.bx, src); w.by = __shfl_sync(0xFFFFFFFFu, r.by, src);
            w.cx = __shfl_sync(0xFFFFFFFFu, r.cx, src); w.cy = __shfl_sync(0xFFFFFFFFu, r.cy, src);
            const int wx0 = __shfl_sync(0xFFFFFFFFu, x0, src), wx1 = __shfl_sync(0xFFFFFFFFu, x1, src);
            const int wy0 = __shfl_sync(0xFFFFFFFFu, y0, src), wy1 = __shfl_sync(0xFFFFFFFFu, y1, src);
            const int wtx0 = __shfl_sync(0xFFFFFFFFu, tx0, src), wtx1 = __shfl_sync(0xFFFFFFFFu, tx1, src);
            const int wty0 = __shfl_sync(0xFFFFFFFFu, ty0, src), wty1 = __shfl_sync(0xFFFFFFFFu, ty1, src);
            const uint32_t wslot = __shfl_sync(0xFFFFFFFFu, slot, src);
            const TriEdges t = prepare_edges(w);
            const int cols = wtx1 - wtx0 + 1, total = cols * (wty1 - wty0 + 1);
            for (int i = (int)lane; i < total; i += 32) {
                const int tx = wtx0 + i % cols, ty = wty0 + i / cols;
                const int lx = max(wx0, tx * TILE_W), hx = min(wx1, tx * TILE_W + TILE_W - 1);
                const int ly = max(wy0, ty * TILE_H), hy = min(wy1, ty * TILE_H + TILE_H - 1);
                if (rect_may_cover(t, lx, hx, ly, hy)) bin_hit<FILL>(W, (uint32_t)ty * U.tiles_x + (uint32_t)tx, wslot);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_alloc : gives every list a contiguous range of list_refs.  Lists need not be laid out in
// list order, so instead of a global scan each CTA scans its 256 counts locally and reserves its
// total with one atomicAdd; list_count is reset to serve as the fill cursor (after k_bin<true> it
// holds the count again, which is what k_tile reads).  total -> counters[1].
// ------------------------------------------------------------------------------------------
constexpr int ALLOC_THREADS = 256;

__global__ void __launch_bounds__(ALLOC_THREADS) k_alloc(const FrameDev W, const uint32_t n_lists) {
    __shared__ uint32_t warp_sum[ALLOC_THREADS / 32];
    __shared__ uint32_t block_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t i = blockIdx.x * ALLOC_THREADS + tid;
    const uint32_t c = i < n_lists ? W.list_count[i] : 0u;

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
    if (i < n_lists) {
        W.list_offset[i] = block_base + warp_sum[warp] + incl - c;
        W.list_count[i] = 0; // becomes the fill cursor
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static int bin_blocks(const FrameDev &W) {
    long long b = ((long long)W.rec_cap + BIN_THREADS - 1) / BIN_THREADS;
    if (b > 148 * 8) b = 148 * 8;
    if (b < 1) b = 1;
    return (int)b;
}
void launch_bin_count(const FrameUniforms &U, const FrameDev &W, cudaStream_t stream) {
    k_bin<false><<<bin_blocks(W), BIN_THREADS, 0, stream>>>(U, W);
}
void launch_alloc(const FrameUniforms &U, const FrameDev &W, cudaStream_t stream) {
    k_alloc<<<(U.n_lists + ALLOC_THREADS - 1) / ALLOC_THREADS, ALLOC_THREADS, 0, stream>>>(W, U.n_lists);
}
void launch_bin_fill(const FrameUniforms &U, const FrameDev &W, cudaStream_t stream) {
    k_bin<true><<<bin_blocks(W), BIN_THREADS, 0, stream>>>(U, W);
}

} // namespace drawb200
