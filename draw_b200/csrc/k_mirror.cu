// k_mirror.cu — the frame's way back to the host (Canvas::as_bytes_slice, mororo18/draw canvas.rs:966-982: the
// reference's frame lives in host memory).
//
// A full read-back is 4*W*H bytes over PCIe per frame — 33 MB at 4K, which caps the renderer at ~1 700 frames/s
// whatever the GPU does.  Most of a frame is usually the clear colour, and it was the clear colour in the frame
// the host mirror already holds: those pixels need not cross the bus again.  k_tile records per tile which of its
// 64x8-pixel strips hold anything else than the clear colour (one bit each, FrameUniforms::tile_state; 0 for a tile
// it only cleared); the canvas remembers the same for the frame its pinned host mirror holds (mirror_state).
// k_mirror copies the strips that are set in either — drawn now, or drawn then and clear since — from the device
// frame to the mirror through its device-mapped address (posted PCIe writes, whole 256-byte rows), and brings
// mirror_state up to date.  The mirror ends up byte-identical to the device frame.  The number of strips copied is
// posted to the frame's status block.
#include "device_math.cuh"

namespace drawb200 {

constexpr int MIRROR_THREADS = 256;

__global__ void __launch_bounds__(MIRROR_THREADS) k_mirror(const uint8_t *__restrict__ color, uint8_t *__restrict__ host_color,
                                                           const uint8_t *__restrict__ tile_state, uint8_t *__restrict__ mirror_state,
                                                           int W_, int H_, int tiles_x, int n_tiles, uint32_t *__restrict__ counters,
                                                           uint32_t *__restrict__ status_word) {
    constexpr int QPR = TILE_W / 4, ROWS_PER_STEP = MIRROR_THREADS / QPR, STEPS = TILE_H / ROWS_PER_STEP;
    static_assert(MIRROR_THREADS % QPR == 0 && TILE_H % ROWS_PER_STEP == 0, "k_mirror geometry");
    const int tid = threadIdx.x;
    uint32_t copied = 0; // CTA-uniform
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t now = tile_state[tile], m = now | mirror_state[tile];
        if (!m) continue; // clear colour on both sides
        copied += (uint32_t)__popc(m);
        const int x = (tile % tiles_x) * TILE_W + (tid % QPR) * 4, r0 = tid / QPR, y0 = (tile / tiles_x) * TILE_H + r0;
        if ((W_ & 3) == 0) {
            if (x < W_) {
                const size_t at = ((size_t)(H_ - 1 - y0) * W_ + x) * 4; // colour rows are y-flipped; both buffers alike
                const ptrdiff_t step = (ptrdiff_t)ROWS_PER_STEP * W_ * 4;
                uint4 q[STEPS];
                bool take[STEPS];
#pragma unroll
                for (int k = 0; k < STEPS; k++) {
                    take[k] = y0 + k * ROWS_PER_STEP < H_ && ((m >> ((r0 + k * ROWS_PER_STEP) / TILE_STRIP_H)) & 1u);
                    if (take[k]) q[k] = __ldcs(reinterpret_cast<const uint4 *>(color + at - k * step));
                }
#pragma unroll
                for (int k = 0; k < STEPS; k++)
                    if (take[k]) *reinterpret_cast<uint4 *>(host_color + at - k * step) = q[k];
            }
        } else { // k_tile reports whole tiles for such canvases
            for (int p = tid; p < TILE_W * TILE_H; p += MIRROR_THREADS) {
                const int px = (tile % tiles_x) * TILE_W + (p & (TILE_W - 1)), py = (tile / tiles_x) * TILE_H + p / TILE_W;
                if (px >= W_ || py >= H_) continue;
                const size_t at = (size_t)(H_ - 1 - py) * W_ + px;
                reinterpret_cast<uint32_t *>(host_color)[at] = reinterpret_cast<const uint32_t *>(color)[at];
            }
        }
        if (tid == 0) mirror_state[tile] = (uint8_t)now;
    }
    // The last CTA to finish posts the frame's total and re-arms the counter.  One 64-bit word (CTAs done << 32 | strips)
    // and one atomic per CTA: no fence — a fence here would make every CTA wait for its PCIe writes to land.
    if (tid == 0) {
        unsigned long long *word = reinterpret_cast<unsigned long long *>(counters);
        const unsigned long long old = atomicAdd(word, (1ull << 32) | copied);
        if ((uint32_t)(old >> 32) == gridDim.x - 1u) {
            *word = 0ull; // every CTA of this launch has added; the next launch follows in stream order
            if (status_word) *status_word = (uint32_t)old + copied;
        }
    }
}

cudaError_t launch_mirror(const uint8_t *color, uint8_t *host_color, const uint8_t *tile_state, uint8_t *mirror_state, int W_, int H_,
                          uint32_t *counters, uint32_t *status_word, cudaStream_t stream, uint64_t *launches) {
    const int tiles_x = (W_ + TILE_W - 1) / TILE_W, tiles_y = (H_ + TILE_H - 1) / TILE_H, n_tiles = tiles_x * tiles_y;
    const int grid = n_tiles < 148 * 8 ? n_tiles : 148 * 8;
    k_mirror<<<grid, MIRROR_THREADS, 0, stream>>>(color, host_color, tile_state, mirror_state, W_, H_, tiles_x, n_tiles, counters, status_word);
    ++*launches;
    return cudaGetLastError();
}

} // namespace drawb200
