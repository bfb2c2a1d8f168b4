// k_raster.cu — rasterises the medium and small triangle references of the whole frame into key pages.
//
// The reference loops over every triangle's bbox one after the other (canvas.rs:668-680).  Doing the same per
// screen tile makes a tile with hundreds of small triangles (a rib cage, a skull) one long serial job on one SM
// while most of the GPU has nothing to do.  This kernel takes that part of the work out of the tiles: every
// (triangle, tile) reference of the medium and small classes is an independent work item, spread evenly over
// all SMs, and the depth test is a 64-bit atomicMin on the key (order-preserving depth bits << 32 | record
// slot) of the pixel in the tile's *key page* in global memory (L2-resident: 16 KB per tile).  Slots are
// handed out in draw order, so the minimum key is exactly the fragment the reference's sequential
// strict-`<` depth test keeps (k_tile.cu, header).  k_tile then starts from the page (its depths
// also serve the large triangles' early depth reject), adds the large triangles, shades, writes
// the frame, and leaves the page empty for the next frame.
//
//   medium reference : one warp; lanes test up to 32 of the triangle's 8x4-pixel blocks exactly
//                      (rect_may_cover), then the warp visits the surviving blocks, one pixel per lane
//   small reference  : one lane (bbox of at most 16 pixels)
#include "device_math.cuh"

namespace drawb200 {

// 128-thread CTAs (8 K registers): a k_raster CTA of the next frame fits into what three k_tile CTAs and a k_front CTA leave of an
// SM's register file, instead of waiting for a k_tile CTA to leave (measured on C3, frames back to back: +4 % over 256 threads)
#ifndef DRAW_RASTER_THREADS
#define DRAW_RASTER_THREADS 128
#endif
constexpr int RASTER_THREADS = DRAW_RASTER_THREADS;
#ifndef DRAW_RASTER_MINB
#define DRAW_RASTER_MINB (1024 / DRAW_RASTER_THREADS)
#endif
constexpr int RASTER_MIN_CTAS = DRAW_RASTER_MINB;

__device__ __forceinline__ void commit_fragment(unsigned long long *cell, unsigned long long key) {
#ifndef DRAW_RASTER_PRECHECK
    // a reduction needs no answer: nothing to wait for, unlike a load-compare-atomic sequence whose L2 round
    // trip per block would be the whole duration of a warp's job
    atomicMin(cell, key);
#else
    // L2 is the point of coherence for the atomics: read it there (an L1 line could predate the reset of the page)
    if (key < __ldcg(cell)) atomicMin(cell, key);
#endif
}

#ifdef DRAW_RASTER_MAXNREG
__global__ void __maxnreg__(DRAW_RASTER_MAXNREG) k_raster(
#else
__global__ void __launch_bounds__(RASTER_THREADS, RASTER_MIN_CTAS) k_raster(
#endif
    const FrameUniforms *__restrict__ Up, const FrameDev W) {
    const FrameUniforms &U = *Up; // per-frame uniforms, device-resident (one upload per frame; the launches never change)
    const CtaTrace trace_(W, 2u);
    if (W.counters[CNT_OVERFLOW] != 0) return; // a buffer overflowed: the host re-renders
    // Prologue: the (tile, slot) pairs of the large and transparent classes go to their tiles' lists (k_front has
    // given every tile its offset; k_tile, the next kernel, reads the lists).  Fire-and-forget work spread over the
    // whole grid, under the latency of the raster work below.
    {
        const uint32_t n_l = W.counters[CNT_L_PAIRS], n_t = W.counters[CNT_T_PAIRS];
        const uint32_t gtid = blockIdx.x * RASTER_THREADS + threadIdx.x, gsize = gridDim.x * RASTER_THREADS;
        for (uint32_t i = gtid; i < n_l; i += gsize) {
            const uint2 pr = __ldg(W.l_pairs + i);
            W.list_refs[W.l_offset[pr.x] + atomicAdd(&W.l_count[pr.x], 1u)] = pr.y;
        }
        for (uint32_t i = gtid; i < n_t; i += gsize) {
            const uint2 pr = __ldg(W.t_pairs + i);
            W.t_refs[W.t_offset[pr.x] + atomicAdd(&W.t_count[pr.x], 1u)] = pr.y;
        }
    }
    const uint32_t n_medium = min(W.counters[CNT_MEDIUM], W.refs_cap), n_small = min(W.counters[CNT_SMALL], W.refs_cap);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_warps = gridDim.x * (RASTER_THREADS / 32), gwarp = blockIdx.x * (RASTER_THREADS / 32) + (threadIdx.x >> 5);
    const float depth_max = U.depth_max;

    // ---- medium references: one per warp ---------------------------------------------------------
    auto medium_ref = [&](const uint2 ref) {
        const uint32_t tile_x = ref.y & (MAX_TILES_X - 1), tile_y = (ref.y >> 10) & (MAX_TILES_Y - 1);
        const uint32_t part = (ref.y >> 21) & 3u, parts = ((ref.y >> 23) & 3u) + 1u; // this entry's share of the blocks
        const uint32_t page = tile_y * U.tiles_x + tile_x;
        const uint4 *q = reinterpret_cast<const uint4 *>(W.prep + record_index(W, ref.x));
        const uint4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4), q5 = __ldg(q + 5),
                    q6 = __ldg(q + 6);
        TriRegs t;
        t.ecx[0] = __uint_as_float(q0.x); t.ecx[1] = __uint_as_float(q0.y); t.ecx[2] = __uint_as_float(q0.z);
        t.ecy[0] = __uint_as_float(q0.w); t.ecy[1] = __uint_as_float(q1.x); t.ecy[2] = __uint_as_float(q1.y);
        t.ek1[0] = __uint_as_float(q1.z); t.ek1[1] = __uint_as_float(q1.w); t.ek1[2] = __uint_as_float(q2.x);
        t.ek2[0] = __uint_as_float(q2.y); t.ek2[1] = __uint_as_float(q2.z); t.ek2[2] = __uint_as_float(q2.w);
        t.f[0] = __uint_as_float(q3.x); t.f[1] = __uint_as_float(q3.y); t.f[2] = __uint_as_float(q3.z);
        t.rf[0] = __uint_as_float(q3.w); t.rf[1] = __uint_as_float(q4.x); t.rf[2] = __uint_as_float(q4.y);
        t.da = __uint_as_float(q4.z); t.db = __uint_as_float(q4.w); t.dc = __uint_as_float(q6.x);
        t.flags = q6.y;
        // bbox clipped to the tile (floats: exact, < 65536)
        const int tx0 = (int)tile_x * TILE_W, ty0 = (int)tile_y * TILE_H;
        const int lx = max((int)__uint_as_float(q5.x), tx0), hx = min((int)__uint_as_float(q5.y), tx0 + TILE_W - 1);
        const int ly = max((int)__uint_as_float(q5.z), ty0), hy = min((int)__uint_as_float(q5.w), ty0 + TILE_H - 1);
        if (lx > hx || ly > hy) return;
        const uint32_t nbx = (uint32_t)(hx - lx) / 8u + 1u, nby = (uint32_t)(hy - ly) / 4u + 1u, n_blocks = nbx * nby; // <= 64
        unsigned long long *page_keys = W.key_pages + (size_t)page * (TILE_W * TILE_H);
#pragma unroll 1
        for (uint32_t b0 = 0; b0 * parts + part < n_blocks; b0 += 32) {
            // (1) exact block test, one block per lane (blocks part, part + parts, ...)
            const uint32_t b = (b0 + lane) * parts + part;
            bool hit = false;
            if (b < n_blocks) {
                const int bx = lx + (int)(b % nbx) * 8, by = ly + (int)(b / nbx) * 4;
                hit = rect_may_cover(t, (float)bx, (float)min(bx + 7, hx), (float)by, (float)min(by + 3, hy));
            }
            uint32_t mask = __ballot_sync(0xFFFFFFFFu, hit);
            // (2) the surviving blocks, one pixel per lane
            while (mask) {
                const uint32_t bb = (b0 + (uint32_t)(__ffs(mask) - 1)) * parts + part;
                mask &= mask - 1;
                const int px = lx + (int)(bb % nbx) * 8 + (int)(lane & 7), py = ly + (int)(bb / nbx) * 4 + (int)(lane >> 3);
                if (px > hx || py > hy) continue;
                float d;
                if (!cover_pixel(t, t.flags, (float)px, (float)py, d)) continue;
                if (!(d < depth_max)) continue; // also rejects NaN; equality with the clear depth fails `<`
                commit_fragment(page_keys + (py - ty0) * TILE_W + (px - tx0), make_key(d, ref.x));
            }
        }
    };
#ifndef DRAW_RASTER_PIPE
#pragma unroll 1
    for (uint32_t i = gwarp; i < n_medium; i += n_warps) medium_ref(__ldg(W.m_refs + i)); // same address in every lane: one broadcast load
#else
    // software pipeline: the next reference is loaded before this one is processed and its record and page
    // entry are requested (L1 prefetch) before the loop comes back to them
    if (gwarp < n_medium) {
        uint2 ref = __ldg(W.m_refs + gwarp);
#pragma unroll 1
        for (uint32_t i = gwarp; i < n_medium; i += n_warps) {
            const uint32_t i_next = i + n_warps;
            const uint2 ref_next = i_next < n_medium ? __ldg(W.m_refs + i_next) : ref;
            if (i_next < n_medium && lane == 0) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(W.prep + record_index(W, ref_next.x)));
            }
            medium_ref(ref);
            ref = ref_next;
        }
    }
#endif

    // ---- small references: one per lane ----------------------------------------------------------
    const uint32_t n_threads = gridDim.x * RASTER_THREADS;
#pragma unroll 1
    for (uint32_t i = blockIdx.x * RASTER_THREADS + threadIdx.x; i < n_small; i += n_threads) {
        const uint2 ref = __ldg(W.s_refs + i);
        const uint32_t tile_x = ref.y & (MAX_TILES_X - 1), tile_y = (ref.y >> 10) & (MAX_TILES_Y - 1);
        const uint32_t page = tile_y * U.tiles_x + tile_x;
        const RasterRec r = load_raster(W.rrec + record_index(W, ref.x)); // 48 bytes; the edge set-up is cheaper than reading the PrepRec
        PrepRec p;
        make_prep(r, p);
        const TriRegs t = tri_from_prep(p);
        const int tx0 = (int)tile_x * TILE_W, ty0 = (int)tile_y * TILE_H;
        const int lx = max((int)(r.bbx & 0xFFFF), tx0), hx = min((int)(r.bbx >> 16), tx0 + TILE_W - 1);
        const int ly = max((int)(r.bby & 0xFFFF), ty0), hy = min((int)(r.bby >> 16), ty0 + TILE_H - 1);
        unsigned long long *page_keys = W.key_pages + (size_t)page * (TILE_W * TILE_H);
        float y = (float)ly;
        for (int yi = ly; yi <= hy; yi++, y = FADD(y, 1.0f)) {
            float x = (float)lx;
            for (int xi = lx; xi <= hx; xi++, x = FADD(x, 1.0f)) {
                float d;
                if (!cover_pixel(t, t.flags, x, y, d)) continue;
                if (!(d < depth_max)) continue;
                commit_fragment(page_keys + (yi - ty0) * TILE_W + (xi - tx0), make_key(d, ref.x));
            }
        }
    }
}

// Fills freshly allocated key pages with KEY_EMPTY (host: ensure_work_buffers).
__global__ void __launch_bounds__(256) k_fill_u64(unsigned long long *__restrict__ dst, size_t n, unsigned long long value) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = value;
}

thread_local unsigned g_raster_ctas = 148u * 16u; // scene.cpp: DRAW_B200_RASTER_CTAS
void launch_raster(const FrameUniforms *dU, const FrameDev &W, cudaStream_t stream) {
    k_raster<<<g_raster_ctas, RASTER_THREADS, 0, stream>>>(dU, W);
}
cudaError_t launch_fill_u64(unsigned long long *dst, size_t n, unsigned long long value, cudaStream_t stream) {
    k_fill_u64<<<148 * 4, 256, 0, stream>>>(dst, n, value);
    return cudaGetLastError();
}

} // namespace drawb200
