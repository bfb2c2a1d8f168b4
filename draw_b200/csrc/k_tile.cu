// k_tile.cu — per-tile raster / depth / shade kernel (mororo18/draw canvas.rs:577-750, 906-960).
//
// One CTA per coarse tile (64x32 px), 8 warps; warp w owns the fine tile (w & 3, w >> 2) of 16x16
// px and lane l owns the 4x2 pixel block at (4 * (l & 3), 2 * (l >> 2)) inside it.  Depth and the
// winning record of every pixel live in registers for the whole kernel; colour and depth are
// written to HBM exactly once at the end (clear fused in).  No atomics and no tensor cores.
//
//   phase 1a  coarse-tile list   staged through shared memory 64 triangles at a time, all warps
//   phase 1b  fine-tile list     each warp stages and consumes its own list, 32 at a time
//   shading   deferred: only the winning triangle of a pixel is shaded (canvas.rs:685-743)
//   phase 2   transparent triangles, in draw order, blended over the shaded colour
//   write     colour rows y-flipped (canvas.rs:955-956), depth rows not (canvas.rs:413-423)
//
// Draw-order semantics without ordered lists: the reference draws triangles sequentially with a
// strict `<` depth test, so for opaque triangles the surviving fragment of a pixel is the
// minimum of (depth, draw id) — ties go to the earlier triangle.  Lists are therefore consumed
// in any order and depth ties are broken by draw id.  Transparent triangles (depth test on,
// depth write off, blend with the current colour, scene/mod.rs:1088) are only visible over the
// final opaque winner W of their pixel if drawn after it: fragment T is blended iff
// id(T) > id(W) and depth(T) < depth(W), in draw order.
#include "shading.cuh"

namespace drawb200 {

constexpr int TILE_THREADS = 256;
constexpr int TILE_WARPS = TILE_THREADS / 32;
constexpr int CHUNK = 64;  // coarse-list triangles staged per round (whole CTA)
constexpr int FCHUNK = 32; // fine-list triangles staged per round (per warp)
constexpr int PX = 8;      // pixels per lane: 4 wide x 2 tall

// One triangle prepared for the pixel loop (25 words; stride 25 is conflict-free for staging writes,
// and every read in the pixel loop is a broadcast).
struct StagedTri {
    float ecx[3], ecy[3], ek1[3], ek2[3], f[3];
    float da, db, dc;
    int x0, x1, y0, y1;
    uint32_t flags, id, slot;
};

__device__ __forceinline__ void stage_triangle(StagedTri &s, const RasterRec &r, uint32_t slot) {
    const TriEdges t = prepare_edges(r);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        s.ecx[i] = t.ecx[i]; s.ecy[i] = t.ecy[i]; s.ek1[i] = t.ek1[i]; s.ek2[i] = t.ek2[i]; s.f[i] = t.f[i];
    }
    s.da = r.da; s.db = r.db; s.dc = r.dc;
    s.x0 = (int)(r.bbx & 0xFFFF); s.x1 = (int)(r.bbx >> 16);
    s.y0 = (int)(r.bby & 0xFFFF); s.y1 = (int)(r.bby >> 16);
    s.flags = t.flags;
    s.id = r.id;
    s.slot = slot;
}

// Literal coverage + depth of one pixel (canvas.rs:673-682) with the reference's divisions.  Used
// for non-tame triangles and for the transparent phase.  Staged coefficients may be
// sign-normalised; the quotients e/f are unchanged by that.
__device__ __noinline__ bool cover_literal(const StagedTri &s, float x, float y, float &depth) {
    float bary[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float e = FSUB(FADD(FADD(FMUL(s.ecx[i], x), FMUL(s.ecy[i], y)), s.ek1[i]), s.ek2[i]);
        bary[i] = FDIV(e, s.f[i]);
    }
    if (!(bary[0] >= 0.0f && bary[1] >= 0.0f && bary[2] >= 0.0f)) return false;
    if (!((bary[0] > 0.0f || (s.flags & 1u)) && (bary[1] > 0.0f || (s.flags & 2u)) && (bary[2] > 0.0f || (s.flags & 4u))))
        return false;
    depth = FADD(FADD(FMUL(bary[0], s.da), FMUL(bary[1], s.db)), FMUL(bary[2], s.dc));
    return true;
}

// Depth test with the draw-order tie rule.
__device__ __forceinline__ void depth_update(float d, uint32_t id, uint32_t slot, float &zb, uint32_t &sl,
                                             const RasterRec *__restrict__ rrec) {
    if (d < zb) {
        zb = d;
        sl = slot;
    } else if (d == zb && sl != NO_SLOT) {
        if (id < __ldg(&rrec[sl].id)) { // equal depth: the earlier draw wins (strict `<`, canvas.rs:923)
            zb = d;
            sl = slot;
        }
    }
}

// Rasterises staged triangle s against this lane's 4x2 block at (bx0, by0) inside the warp's fine
// tile at (fx0, fy0).
__device__ __forceinline__ void raster_triangle(const StagedTri &s, int fx0, int fy0, int bx0, int by0,
                                                float (&zb)[PX], uint32_t (&sl)[PX],
                                                const RasterRec *__restrict__ rrec) {
    if (s.x1 < fx0 || s.x0 > fx0 + FINE - 1 || s.y1 < fy0 || s.y0 > fy0 + FINE - 1) return; // warp-uniform
    const int lo_x = max(s.x0, bx0), hi_x = min(s.x1, bx0 + 3);
    const int lo_y = max(s.y0, by0), hi_y = min(s.y1, by0 + 1);
    if (lo_x > hi_x || lo_y > hi_y) return;
    const uint32_t flags = s.flags, id = s.id, slot = s.slot;
    if (!(flags & TRI_SLOW)) {
        // block-level reject at the best corner of the clipped block (exact, see rect_may_cover)
        bool any = true;
#pragma unroll
        for (int e = 0; e < 3; e++) {
            const float cx = s.ecx[e], cy = s.ecy[e];
            const float xm = (float)(cx >= 0.0f ? hi_x : lo_x), ym = (float)(cy >= 0.0f ? hi_y : lo_y);
            const float em = FSUB(FADD(FADD(FMUL(cx, xm), FMUL(cy, ym)), s.ek1[e]), s.ek2[e]);
            any = any && (em > 0.0f || (em == 0.0f && (flags & (1u << e))));
        }
        if (!any) return;
        float pxs[3][4], pys[3][2];
#pragma unroll
        for (int e = 0; e < 3; e++) {
#pragma unroll
            for (int i = 0; i < 4; i++) pxs[e][i] = FMUL(s.ecx[e], (float)(bx0 + i));
#pragma unroll
            for (int j = 0; j < 2; j++) pys[e][j] = FMUL(s.ecy[e], (float)(by0 + j));
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int x = bx0 + i, y = by0 + j;
                if (x < lo_x || x > hi_x || y < lo_y || y > hi_y) continue;
                // f > 0 after sign normalisation: alpha >= 0 <=> e >= 0, alpha > 0 <=> e > 0
                const float e0 = FSUB(FADD(FADD(pxs[0][i], pys[0][j]), s.ek1[0]), s.ek2[0]);
                const float e1 = FSUB(FADD(FADD(pxs[1][i], pys[1][j]), s.ek1[1]), s.ek2[1]);
                const float e2 = FSUB(FADD(FADD(pxs[2][i], pys[2][j]), s.ek1[2]), s.ek2[2]);
                const bool in = (e0 > 0.0f || (e0 == 0.0f && (flags & 1u))) &&
                                (e1 > 0.0f || (e1 == 0.0f && (flags & 2u))) &&
                                (e2 > 0.0f || (e2 == 0.0f && (flags & 4u)));
                if (!in) continue;
                const float alpha = FDIV(e0, s.f[0]), beta = FDIV(e1, s.f[1]), gama = FDIV(e2, s.f[2]);
                const float d = FADD(FADD(FMUL(alpha, s.da), FMUL(beta, s.db)), FMUL(gama, s.dc)); // canvas.rs:682
                depth_update(d, id, slot, zb[j * 4 + i], sl[j * 4 + i], rrec);
            }
        }
    } else {
#pragma unroll
        for (int p = 0; p < PX; p++) {
            const int x = bx0 + (p & 3), y = by0 + (p >> 2);
            if (x < lo_x || x > hi_x || y < lo_y || y > hi_y) continue;
            float d;
            if (!cover_literal(s, (float)x, (float)y, d)) continue;
            depth_update(d, id, slot, zb[p], sl[p], rrec);
        }
    }
}

__global__ void __launch_bounds__(TILE_THREADS) k_tile(const __grid_constant__ FrameUniforms U, const SceneDev S,
                                                       const FrameDev W, uint8_t *__restrict__ color,
                                                       float *__restrict__ depth) {
    __shared__ StagedTri st_coarse[CHUNK];
    __shared__ StagedTri st_fine[TILE_WARPS][FCHUNK];

    const uint32_t tile_x = blockIdx.x % U.tiles_x;
    const uint32_t tile_y = U.tile_y_begin + blockIdx.x / U.tiles_x;
    const uint32_t tile = tile_y * U.tiles_x + tile_x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // warp -> fine tile, lane -> 4x2 block (canvas coordinates: x right, y = depth-buffer row)
    const int gx = (int)tile_x * FINE_PER_TILE_X + (warp & 3), gy = (int)tile_y * FINE_PER_TILE_Y + (warp >> 2);
    const int fx0 = gx * FINE, fy0 = gy * FINE;
    const int bx0 = fx0 + (lane & 3) * 4, by0 = fy0 + (lane >> 2) * 2;

    float zb[PX];
    uint32_t sl[PX];
#pragma unroll
    for (int i = 0; i < PX; i++) {
        zb[i] = U.depth_max;
        sl[i] = NO_SLOT;
    }

    const bool usable = W.counters[2] == 0;

    // ---- phase 1a: the coarse tile's list, all warps --------------------------------------------
    {
        const uint32_t begin = usable ? W.list_offset[tile] : 0u, end = usable ? W.list_offset[tile + 1] : 0u;
        for (uint32_t base = begin; base < end; base += CHUNK) {
            const int n = (int)min((uint32_t)CHUNK, end - base);
            __syncthreads();
            if (tid < n) {
                const uint32_t slot = W.list_refs[base + tid];
                stage_triangle(st_coarse[tid], load_raster(W.rrec + slot), slot);
            }
            __syncthreads();
            for (int k = 0; k < n; k++) raster_triangle(st_coarse[k], fx0, fy0, bx0, by0, zb, sl, W.rrec);
        }
    }
    // ---- phase 1b: this warp's fine-tile list ---------------------------------------------------
    {
        const uint32_t list = U.n_coarse + (uint32_t)gy * U.fine_nx + (uint32_t)gx;
        const uint32_t begin = usable ? W.list_offset[list] : 0u, end = usable ? W.list_offset[list + 1] : 0u;
        StagedTri *mine = st_fine[warp];
        for (uint32_t base = begin; base < end; base += FCHUNK) {
            const int n = (int)min((uint32_t)FCHUNK, end - base);
            __syncwarp();
            if (lane < n) {
                const uint32_t slot = W.list_refs[base + lane];
                stage_triangle(mine[lane], load_raster(W.rrec + slot), slot);
            }
            __syncwarp();
            for (int k = 0; k < n; k++) raster_triangle(mine[k], fx0, fy0, bx0, by0, zb, sl, W.rrec);
        }
    }

    // ---- deferred shading of the opaque winners, fused clear ------------------------------------
    // colour as r | g << 8 | b << 16 | pad << 24 ; clear = azul_bb (155,186,255), pad 255 (canvas.rs:131)
    uint32_t col[PX];
    uint32_t wid[PX]; // draw id of the opaque winner (for the transparent phase), NO_SLOT = none
#pragma unroll
    for (int p = 0; p < PX; p++) {
        col[p] = 155u | (186u << 8) | (255u << 16) | (255u << 24);
        wid[p] = NO_SLOT;
        if (sl[p] != NO_SLOT) {
            const RasterRec r = load_raster(W.rrec + sl[p]);
            float d, op;
            const uint32_t rgb = shade_pixel(S, r, W.srec + sl[p], (float)(bx0 + (p & 3)), (float)(by0 + (p >> 2)), &d, &op);
            col[p] = rgb | (255u << 24);
            wid[p] = r.id;
        }
    }

    // ---- phase 2: transparent triangles in draw order (scene/mod.rs:1088-1246) ------------------
    const uint32_t n_tslots = usable ? S.n_transparent * 4u : 0u;
    for (uint32_t base = 0; base < n_tslots; base += CHUNK) {
        const int n = (int)min((uint32_t)CHUNK, n_tslots - base);
        __syncthreads();
        if (tid < n) {
            RasterRec r = load_raster(W.t_rrec + base + tid);
            if (r.id == NO_SLOT) { // empty slot: an empty bbox makes every lane skip it
                r.bbx = 1u;        // x_min = 1 > x_max = 0
                r.bby = 1u;
            }
            stage_triangle(st_coarse[tid], r, base + tid);
        }
        __syncthreads();
        for (int k = 0; k < n; k++) {
            const StagedTri &s = st_coarse[k];
            if (s.x1 < fx0 || s.x0 > fx0 + FINE - 1 || s.y1 < fy0 || s.y0 > fy0 + FINE - 1) continue;
            const int lo_x = max(s.x0, bx0), hi_x = min(s.x1, bx0 + 3);
            const int lo_y = max(s.y0, by0), hi_y = min(s.y1, by0 + 1);
            if (lo_x > hi_x || lo_y > hi_y) continue;
#pragma unroll
            for (int p = 0; p < PX; p++) {
                const int x = bx0 + (p & 3), y = by0 + (p >> 2);
                if (x < lo_x || x > hi_x || y < lo_y || y > hi_y) continue;
                float d;
                if (!cover_literal(s, (float)x, (float)y, d)) continue;
                if (!(wid[p] == NO_SLOT || s.id > wid[p])) continue; // drawn before the opaque winner: overwritten
                if (!(d < zb[p])) continue;                          // canvas.rs:923, depth write is off
                const RasterRec r = load_raster(W.t_rrec + s.slot);
                float d2, op;
                const uint32_t rgb = shade_pixel(S, r, W.t_srec + s.slot, (float)x, (float)y, &d2, &op);
                // canvas.rs:916-921: opacity < 1 blends with the stored colour, else replaces it
                col[p] = op < 1.0f ? blend_rgb(col[p], rgb, op) : (rgb | (255u << 24));
            }
        }
    }

    // ---- single write-back ------------------------------------------------------------------------
    const int W_ = (int)U.canvas_w, H_ = (int)U.canvas_h;
    const bool vec_ok = (W_ & 3) == 0;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int y = by0 + j;
        if (y >= H_ || bx0 >= W_) continue;
        uint32_t px[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t c = col[j * 4 + i]; // r g b pad -> memory order b g r pad
            px[i] = ((c >> 16) & 255u) | (c & 0x0000FF00u) | ((c & 255u) << 16) | (c & 0xFF000000u);
        }
        const size_t crow = (size_t)(H_ - 1 - y) * W_ + bx0, drow = (size_t)y * W_ + bx0;
        if (vec_ok) {
            *reinterpret_cast<uint4 *>(color + crow * 4) = make_uint4(px[0], px[1], px[2], px[3]);
            *reinterpret_cast<float4 *>(depth + drow) = make_float4(zb[j * 4], zb[j * 4 + 1], zb[j * 4 + 2], zb[j * 4 + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (bx0 + i < W_) {
                    reinterpret_cast<uint32_t *>(color)[crow + i] = px[i];
                    depth[drow + i] = zb[j * 4 + i];
                }
            }
        }
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
    if (stripe_tiles) k_tile<<<stripe_tiles, TILE_THREADS, 0, stream>>>(U, S, W, color, depth);
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
