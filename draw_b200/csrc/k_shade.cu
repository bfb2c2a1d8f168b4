// k_shade.cu — resolves key pages into pixels: one thread per pixel (canvas.rs:682-743, 906-960).
//
// For the tiles k_alloc marked for it (a key page and few large triangles), the page holds the winning
// fragment of every pixel — written by k_raster (medium and small triangles) and, where the tile has
// large triangles, merged with them by k_tile.  Shading a pixel is three dependent fetches (records,
// material, texels) and ~250 instructions; inside a tile CTA that is a serial, latency-bound tail, here
// every pixel is an independent thread of a full-occupancy grid.  The kernel also writes the clear
// colour where a pixel has no fragment (the fused clear of these tiles) and leaves the page empty for
// the next frame.
#include "shading.cuh"

namespace drawb200 {

constexpr int SHADE_THREADS = 256;
constexpr int SHADE_MATERIAL_CACHE = 128; // materials kept in shared memory (8 KB); larger tables are read from global memory
constexpr int SHADE_UNITS = TILE_W * TILE_H / SHADE_THREADS; // 256-pixel units per tile: whole tile rows
static_assert(SHADE_THREADS % TILE_W == 0 && (TILE_W * TILE_H) % SHADE_THREADS == 0, "a unit is a whole number of tile rows");

__global__ void __launch_bounds__(SHADE_THREADS, 4) k_shade(const FrameUniforms *__restrict__ Up, const SceneDev S, const FrameDev W) {
    const FrameUniforms &U = *Up;
    __shared__ float u8tab[256]; // (u8 as f32) / 255.0
    __shared__ __align__(16) MaterialDev mat_cache[SHADE_MATERIAL_CACHE];
#ifdef DRAW_TAP_SHADE
    const long long tap0 = clock64();
#endif
    fill_u8_table(u8tab, threadIdx.x, SHADE_THREADS);
    // the material table is part of the scene (uploaded by add_object), not of the frame: it can be read before
    // the previous kernels have finished.  Cached, the material fetch is no longer a round trip between the
    // record fetch and the texel fetch.
    const bool cached = S.n_materials <= SHADE_MATERIAL_CACHE;
    if (cached)
        for (uint32_t i = threadIdx.x; i < S.n_materials * (sizeof(MaterialDev) / 16); i += SHADE_THREADS)
            reinterpret_cast<uint4 *>(mat_cache)[i] = __ldg(reinterpret_cast<const uint4 *>(S.materials) + i);
    const CtaTrace trace_(W, 9u);
    pdl_prologue(false);
    __syncthreads();
    if (W.counters[2] != 0) return; // overflow: k_raster and k_tile left the pages untouched, the host re-renders
    const uint32_t n_units = W.counters[14] * SHADE_UNITS;
#ifdef DRAW_TAP_SHADE
    const long long tap1 = clock64() + (n_units & 1 ? 0 : 0) * (long long)u8tab[0];
#endif
    const int W_ = (int)U.canvas_w, H_ = (int)U.canvas_h;
    uint8_t *__restrict__ color = U.color;
    float *__restrict__ depth = U.depth;
    const int lx = threadIdx.x % TILE_W, lr = threadIdx.x / TILE_W;
    // The key of the unit after the current one is fetched while the current one is shaded.
    auto fetch = [&](uint32_t u, int &x, int &y) -> unsigned long long {
        const uint32_t txy = W.shade_tiles[u / SHADE_UNITS];
        const uint32_t tile_x = txy & (MAX_TILES_X - 1), tile_y = txy >> 10;
        const uint32_t page = W.tile_page[tile_y * U.tiles_x + tile_x];
        const int ly = (int)(u % SHADE_UNITS) * (SHADE_THREADS / TILE_W) + lr;
        unsigned long long *cell = W.key_pages + (size_t)page * (TILE_W * TILE_H) + ly * TILE_W + lx;
        const unsigned long long key = __ldcg(cell);
        __stcg(cell, KEY_EMPTY); // a warp covers 32 consecutive keys: whole 128-byte lines
        x = (int)tile_x * TILE_W + lx;
        y = (int)tile_y * TILE_H + ly;
        return key;
    };
    int x = 0, y = 0, nx = 0, ny = 0;
    unsigned long long key = blockIdx.x < n_units ? fetch(blockIdx.x, x, y) : KEY_EMPTY, next_key = KEY_EMPTY;
#ifdef DRAW_TAP_SHADE
    const long long tap2 = clock64() + (long long)(key & 0);
    long long tap3 = 0, tap4 = 0;
    int tap_n = 0;
#endif
#pragma unroll 1
    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x, key = next_key, x = nx, y = ny) {
        if (u + gridDim.x < n_units) next_key = fetch(u + gridDim.x, nx, ny);
#ifdef DRAW_TAP_SHADE
        const long long ta = clock64();
#endif
        if (x >= W_ || y >= H_) continue;
        const uint32_t slot = (uint32_t)key;
        uint32_t c = 155u | (186u << 8) | (255u << 16) | (255u << 24); // azul_bb, pad 255 (canvas.rs:131)
        float d = U.depth_max;
        if (slot != NO_SLOT) {
            float op;
            uint32_t id;
            c = (cached ? shade_pixel_prep<true>(mat_cache, S.texels, u8tab, W.prep + slot, W.srec + slot, (float)x, (float)y, &d, &op, &id)
                        : shade_pixel_prep<false>(S.materials, S.texels, u8tab, W.prep + slot, W.srec + slot, (float)x, (float)y, &d, &op, &id)) |
                (255u << 24);
        }
        // r g b pad -> memory order b g r pad; colour rows are y-flipped (canvas.rs:955-956), depth rows are not
        __stcs(reinterpret_cast<uint32_t *>(color) + (size_t)(H_ - 1 - y) * W_ + x,
               ((c >> 16) & 255u) | (c & 0x0000FF00u) | ((c & 255u) << 16) | (c & 0xFF000000u));
        __stcs(depth + (size_t)y * W_ + x, d);
#ifdef DRAW_TAP_SHADE
        tap3 += clock64() - ta + (long long)(c & 0);
        tap_n++;
#endif
    }
#ifdef DRAW_TAP_SHADE
    tap4 = clock64();
    if (W.tile_cycles && threadIdx.x == 0 && blockIdx.x < 64) {
        uint32_t *o = W.tile_cycles + 8 * blockIdx.x;
        o[0] = (uint32_t)(tap1 - tap0); o[1] = (uint32_t)(tap2 - tap1); o[2] = (uint32_t)tap3; o[3] = (uint32_t)tap_n; o[4] = (uint32_t)(tap4 - tap0);
        unsigned sm; asm volatile("mov.u32 %0, %smid;" : "=r"(sm)); o[5] = sm;
        o[6] = (uint32_t)(tap0 & 0xFFFFFFFF);
    }
#endif
}

void launch_shade(const FrameUniforms &U, const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream) {
    if (U.tile_y_end > U.tile_y_begin && U.defer_max) launch_pdl(k_shade, 148u * 6u, SHADE_THREADS, stream, dU, S, W);
}

} // namespace drawb200
