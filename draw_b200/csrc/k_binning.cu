// k_binning.cu — sorts the frame's raster records into screen-tile lists.
//
// The reference has no such stage (it loops each triangle's bbox directly, canvas.rs:668-670);
// it exists so that k_tile can keep a tile's colour and depth on chip and write HBM once.
//
//   k_bin<false>  per record  count the lists it belongs to
//   k_scan        one CTA     exclusive scan of the counts -> list offsets
//   k_bin<true>   per record  scatter its slot into those lists
//
// Two list classes (device_types.h): a record whose bbox spans at most 2x2 fine tiles (16x16 px)
// goes to fine-tile lists, read by one warp of k_tile each; any larger record goes to the lists
// of the coarse tiles (64x32 px) it overlaps, read by all 8 warps of that tile's CTA.  A tile is
// only referenced if the triangle can actually cover a pixel in it (exact corner test,
// rect_may_cover), not merely because its bbox touches it.  Records covering many coarse tiles
// are binned by the whole warp (ballot picks them, lanes stride over the tiles).
// Order inside a list is irrelevant: k_tile resolves fragments by (depth, draw id).
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
        const int gx0 = x0 / FINE, gx1 = x1 / FINE, gy0 = y0 / FINE, gy1 = y1 / FINE;
        const bool small = (gx1 - gx0 <= 1) && (gy1 - gy0 <= 1);
        const int tx0 = x0 / TILE_W, tx1 = x1 / TILE_W;
        int ty0 = y0 / TILE_H, ty1 = y1 / TILE_H;
        if (ty0 < (int)U.tile_y_begin) ty0 = (int)U.tile_y_begin;
        if (ty1 > (int)U.tile_y_end - 1) ty1 = (int)U.tile_y_end - 1;
        const int n_coarse_hit = ty0 <= ty1 ? (tx1 - tx0 + 1) * (ty1 - ty0 + 1) : 0;
        const bool wide = valid && !small && n_coarse_hit > WIDE_TILES;

        if (valid && !wide) {
            const TriEdges t = prepare_edges(r);
            if (small) {
                for (int gy = gy0; gy <= gy1; gy++) {
                    const int trow = gy / FINE_PER_TILE_Y;
                    if (trow < (int)U.tile_y_begin || trow >= (int)U.tile_y_end) continue;
                    for (int gx = gx0; gx <= gx1; gx++) {
                        const int lx = max(x0, gx * FINE), hx = min(x1, gx * FINE + FINE - 1);
                        const int ly = max(y0, gy * FINE), hy = min(y1, gy * FINE + FINE - 1);
                        if (rect_may_cover(t, lx, hx, ly, hy))
                            bin_hit<FILL>(W, U.n_coarse + (uint32_t)gy * U.fine_nx + (uint32_t)gx, slot);
                    }
                }
            } else {
                for (int ty = ty0; ty <= ty1; ty++)
                    for (int tx = tx0; tx <= tx1; tx++) {
                        const int lx = max(x0, tx * TILE_W), hx = min(x1, tx * TILE_W + TILE_W - 1);
                        const int ly = max(y0, ty * TILE_H), hy = min(y1, ty * TILE_H + TILE_H - 1);
                        if (rect_may_cover(t, lx, hx, ly, hy)) bin_hit<FILL>(W, (uint32_t)ty * U.tiles_x + (uint32_t)tx, slot);
                    }
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
// k_scan : exclusive scan of list_count -> list_offset, cursors reset, total -> counters[1]
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(SCAN_THREADS) k_scan(const FrameDev W, const uint32_t n_lists) {
    __shared__ uint32_t warp_sum[SCAN_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (n_lists + SCAN_THREADS - 1) / SCAN_THREADS;
    const uint32_t begin = tid * per < n_lists ? tid * per : n_lists;
    const uint32_t end = begin + per < n_lists ? begin + per : n_lists;

    uint32_t local = 0;
    for (uint32_t i = begin; i < end; i++) local += W.list_count[i];

    uint32_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (uint32_t)d) incl += up;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = warp_sum[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= (uint32_t)d) wi += up;
        }
        warp_sum[lane] = wi - w; // exclusive
    }
    __syncthreads();
    uint32_t run = warp_sum[warp] + incl - local;
    for (uint32_t i = begin; i < end; i++) {
        const uint32_t c = W.list_count[i];
        W.list_offset[i] = run;
        W.list_count[i] = 0; // becomes the fill cursor
        run += c;
    }
    if (tid == SCAN_THREADS - 1) {
        W.list_offset[n_lists] = run;
        W.counters[1] = run;
        if (run > W.refs_cap) atomicOr(&W.counters[2], OVERFLOW_REFS);
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
void launch_scan(const FrameUniforms &U, const FrameDev &W, cudaStream_t stream) {
    k_scan<<<1, SCAN_THREADS, 0, stream>>>(W, U.n_lists);
}
void launch_bin_fill(const FrameUniforms &U, const FrameDev &W, cudaStream_t stream) {
    k_bin<true><<<bin_blocks(W), BIN_THREADS, 0, stream>>>(U, W);
}

} // namespace drawb200
