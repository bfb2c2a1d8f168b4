// device_types.h — plain structs shared by the host library (scene.cpp) and the kernels.
#pragma once
#include <stdint.h>
#include <vector_types.h>

namespace drawb200 {

// Screen partition.  A CTA of k_tile owns one tile of TILE_W x TILE_H pixels and keeps its depth /
// winner / colour on chip.  Each tile has three lists of raster records, classed by the area of
// (record bbox intersected with the tile):
//   large   list id = tile                  > MEDIUM_AREA px : every lane tests its own pixels
//   medium  list id = n_coarse + tile       <= MEDIUM_AREA px: one record per warp, lane per pixel
//   small   list id = 2 * n_coarse + tile   <= SMALL_AREA px : one record per lane
// REGION is the square each warp owns during the large-record phase (8 pixels per lane).
#ifndef DRAW_TILE_W
#define DRAW_TILE_W 64
#endif
#ifndef DRAW_TILE_H
#define DRAW_TILE_H 32
#endif
constexpr int TILE_W = DRAW_TILE_W, TILE_H = DRAW_TILE_H, REGION = 16;
#ifndef DRAW_MEDIUM_AREA
#define DRAW_MEDIUM_AREA 1024
#endif
#ifndef DRAW_SMALL_AREA
#define DRAW_SMALL_AREA 8
#endif
constexpr int SMALL_AREA = DRAW_SMALL_AREA, MEDIUM_AREA = DRAW_MEDIUM_AREA;
constexpr int LISTS_PER_TILE = 3;
constexpr uint32_t NO_SLOT = 0xFFFFFFFFu, NO_PAGE = 0xFFFFFFFFu;
constexpr unsigned long long KEY_EMPTY = ~0ull;
#ifndef DRAW_TILE_THREADS
#define DRAW_TILE_THREADS 256
#endif
constexpr int TILE_THREADS = DRAW_TILE_THREADS;
// k_tile phase A geometry: a warp owns a REGION x REGION_H rectangle (4 lanes across, 8 down), a lane
// a 4 x BLK_H block of it.
constexpr int BLK_H = TILE_W * TILE_H / TILE_THREADS / 4;
constexpr int REGION_H = 8 * BLK_H;

// k_tile work items, built by k_alloc (heaviest first).  An item is a tile plus a pixel window of it in
// units of warp regions (a 4x4 grid of REGION x REGION_H rectangles): the whole tile, or one of the
// 2 / 4 / 8 / 16 windows a dense tile is cut into.
//   bits 0-9 tile x | 10-20 tile y | 21-22 window x0 | 23-24 window y0 | 25-26 window w-1 | 27-28 window h-1 | 29 ITEM_DEFER
// Tiles with nothing binned to them are not items: they are listed in FrameDev::empty_tiles (as
// x | y << 10, count in counters[13]) and written by k_clear_empty.
constexpr int REGIONS_X = TILE_W / REGION, REGIONS_Y = TILE_H / REGION_H;
constexpr bool TILE_SPLITTABLE = REGIONS_X == 4 && (REGIONS_Y == 4 || REGIONS_Y == 2);
#ifndef DRAW_TILE_MAX_SPLIT
#define DRAW_TILE_MAX_SPLIT 8
#endif
#ifndef DRAW_TILE_SPLIT_MIN_COST
#define DRAW_TILE_SPLIT_MIN_COST 256
#endif
constexpr int TILE_MAX_SPLIT = DRAW_TILE_MAX_SPLIT < REGIONS_X * REGIONS_Y ? DRAW_TILE_MAX_SPLIT : REGIONS_X * REGIONS_Y; // 1, 2, 4, 8 or 16
constexpr int TILE_SPLIT_MIN_COST = DRAW_TILE_SPLIT_MIN_COST; // windows are not made cheaper than this (k_alloc cost units)
constexpr int TILE_EXTRA_ITEMS = 1024;                        // work-list slots beyond one per tile
#ifndef DRAW_TILE_SPLIT_DIV
#define DRAW_TILE_SPLIT_DIV 296
#endif
constexpr int TILE_SPLIT_DIV = DRAW_TILE_SPLIT_DIV;           // a window should cost about total / this (<= TILE_EXTRA_ITEMS)
constexpr uint32_t ITEM_NONE = 0xFFFFFFFFu;
constexpr uint32_t ITEM_DEFER = 1u << 29; // k_tile only adds the large triangles to the tile's key page; k_shade shades it
constexpr int N_COUNTERS = 64, ITEM_CURSOR = 32; // FrameDev::counters; k_tile's item cursor sits in the second 128-byte line
constexpr uint32_t MAX_TILES_X = 1u << 10, MAX_TILES_Y = 1u << 11;
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t make_item(uint32_t tx, uint32_t ty, uint32_t splits, uint32_t i) {
    const uint32_t sx = splits >= 8 ? 4u : (splits >= 2 ? 2u : 1u), sy = splits / sx; // sy <= REGIONS_Y as splits <= TILE_MAX_SPLIT
    const uint32_t rw = (uint32_t)REGIONS_X / sx, rh = (uint32_t)REGIONS_Y / sy, rx0 = (i % sx) * rw, ry0 = (i / sx) * rh;
    return tx | ty << 10 | rx0 << 21 | ry0 << 23 | (rw - 1u) << 25 | (rh - 1u) << 27;
}

// Per-frame constants.  They live in device memory (one small upload per frame) and every kernel takes a
// pointer to them, so that a frame's launches are identical from frame to frame and can be replayed as a
// CUDA graph (scene.cpp).
struct FrameUniforms {
    float m[16];        // matrix_transf, row-major (scene/mod.rs:817-899)
    float planes[6][4]; // near, far, right, left, top, bottom as (nx,ny,nz,k) (scene/mod.rs:481-593)
    float cam[3];
    float light[3];
    float off_x, off_y; // canvas offset (canvas.rs:382-385)
    float depth_max;    // canvas.rs:403
    uint32_t canvas_w, canvas_h;
    uint32_t tiles_x, tiles_y;         // whole canvas, in coarse tiles
    uint32_t tile_y_begin, tile_y_end; // coarse tile rows rendered by this launch (sort-first stripe)
    uint32_t n_coarse;                 // tiles_x * tiles_y
    uint32_t n_lists;                  // LISTS_PER_TILE * n_coarse: large, medium, small lists
    uint32_t split_min_cost, split_div, split_max; // k_alloc's tile splitting policy (defaults: TILE_SPLIT_*)
    uint32_t defer_max;                // tiles with a key page and fewer large references than this are shaded by k_shade
    uint32_t *status_host;             // pinned host memory of the canvas (device-mapped): k_tile copies counters[0..15] there
    uint8_t *color;                    // the canvas: BGRA8, row 0 = top (canvas.rs:955-956)
    float *depth;                      // depth buffer, row 0 = y 0 (canvas.rs:413-423)
    uint32_t has_transparent;          // transparent triangles are not binned: no tile may take the empty-tile path
    uint32_t pdl_early;                // geometry / binning kernels trigger their dependents at once (device_math.cuh)
    uint32_t cost_shade;               // k_alloc's cost of shading one fully covered tile (same units as COST_* in k_binning.cu)
    uint32_t bin_records_per_warp;     // k_bin: with at most this many records per warp of its grid a warp takes a record, else a thread
    uint32_t clear_in_tile;            // k_tile's CTAs write the empty tiles between their items; no k_clear_empty launch
};

// Texture (scene/mod.rs:206-216) with both TextureMaps flattened into the texel pool.  Maps are
// stored with 4 bytes per texel (3-component images are padded at upload) so that a texel is one
// aligned 32-bit load; offsets are byte offsets into the pool, multiples of 4.  64 bytes, read by
// shade_pixel as four uint4 — keep the field order.
struct alignas(16) MaterialDev {
    float ka[3], kd[3], ks[3];
    float alpha;
    uint32_t ka_off, ka_w;
    uint32_t kd_off, kd_w, kd_h;
    uint32_t ka_h;
};

// Scene geometry, SoA, all objects concatenated (indices are global after upload).
struct SceneDev {
    const float *px, *py, *pz;       // positions           [n_vertices]
    const float *nx, *ny, *nz;       // normals             [n_normals]
    const float *tu, *tv;            // uv                  [n_uvs]
    const uint32_t *idx[9];          // v0 v1 v2 t0 t1 t2 n0 n1 n2, one stream each [n_triangles]
    const uint32_t *tri_mat;         // material id | (transparent << 31), draw order [n_triangles]
    const uint32_t *tri_tslot;       // transparent triangles: ordinal among them; else unused (may be null)
    const MaterialDev *materials;
    const uint8_t *texels;
    uint32_t n_vertices, n_triangles, n_transparent, n_materials;
};

// What the tile kernel needs to rasterise one screen triangle (48 B).
struct alignas(16) RasterRec {
    float ax, ay, bx, by, cx, cy; // snapped vertex centres (canvas.rs:585-587)
    float da, db, dc;             // per-vertex depth (scene/mod.rs:924)
    uint32_t id;                  // draw id = 4 * draw-order index + clip output; NO_SLOT = empty
    uint32_t bbx, bby;            // x_min | x_max << 16, y_min | y_max << 16 (canvas.rs:640-658)
};

// A raster record prepared for coverage and depth tests (128 B = 8 x uint4, one per thread when a
// tile CTA stages it).  Written once per record by k_bin<count>, read by every tile (and window) the
// record is binned to, so the edge set-up is not repeated per reference.  Field order is fixed:
// quad 5 (x0 x1 y0 y1) is what the window filter of k_tile reads on its own.
struct alignas(16) PrepRec {
    float ecx[3], ecy[3], ek1[3], ek2[3]; // sign-normalised edge functions (device_math.cuh: prepare_edges)
    float f[3];                           // their values at the opposite vertex (> 0 when tame)
    float rf[3];                          // RN(1 / f): exact_div, early depth reject
    float da, db;
    float x0, x1, y0, y1;                 // bbox as floats (exact: < 65536)
    float dc;
    uint32_t flags;                       // TRI_* bits
    uint32_t id;                          // draw id
    uint32_t slot;                        // record slot (filled in when staged)
    uint32_t pad[4];
};
static_assert(sizeof(PrepRec) == 128, "PrepRec is staged as 8 uint4");
constexpr int PREP_WORDS = 28; // meaningful words; shared-memory copies use this (odd-ish) stride + 1

// What shading needs for the winning triangle of a pixel (144 B).
struct alignas(16) ShadeRec {
    float n[3][3]; // normal    per corner
    float l[3][3]; // light     per corner
    float h[3][3]; // halfway   per corner
    float uv[3][2];
    uint32_t material;
    uint32_t pad[2];
};

// Per-frame work buffers (owned by the scene, sized for the scene and the canvas).
struct FrameDev {
    float *v_lx, *v_ly, *v_lz;  // light     (scene/mod.rs:922)
    float *v_hx, *v_hy, *v_hz;  // halfway   (:925)
    float *v_depth;             // :924
    float *v_sx, *v_sy;         // screen xy after the divide (:1047-1058), before the canvas offset
    uint32_t *v_flags;          // 2 bits per plane: bit 2p = f>0, bit 2p+1 = f<=0
    RasterRec *rrec;            // opaque records, slots in draw order [rec_cap]
    PrepRec *prep;              // the same records prepared for rasterisation (k_bin<count>) [rec_cap]
    ShadeRec *srec;
    RasterRec *t_rrec;          // transparent records, slot = 4*ordinal + k, in draw order [4*n_transparent]
    ShadeRec *t_srec;
    uint32_t *list_count;       // per list (large, medium, small per tile): count, then fill cursor [n_lists]
    uint32_t *list_offset;      // first entry of each list in list_refs [n_lists + 1]
    uint32_t *list_refs;        // large lists: record slots [refs_cap]
    uint2 *m_refs, *s_refs;     // medium / small lists: (record slot, tile x | y << 10) [refs_cap each]
    uint32_t *tile_page;        // key page of each tile, or NO_PAGE [n_coarse]
    unsigned long long *key_pages; // page p = TILE_W * TILE_H keys (depth key << 32 | slot), all ones = empty; k_raster
                                   // fills them with atomicMin, k_tile consumes and resets them [page_cap pages]
    uint32_t page_cap;
    uint32_t *counters;         // [0] records  [1] refs  [2] overflow bits  [3] k_setup CTA ticket  [4] k_alloc CTAs done  [5] clip queue length
                                // [6] total tile cost  [8..10] large / medium / small references  [11] key pages handed out  [13] empty tiles  [14] tiles for k_shade  [15] k_tile items  [32] k_tile item cursor  [N_COUNTERS]
    uint32_t *tile_cost;        // estimated k_tile work per tile [n_coarse]
    uint32_t *tile_order;       // k_tile work items (make_item), heaviest first, padded with ITEM_NONE [n_coarse + TILE_EXTRA_ITEMS]
    uint32_t *shade_tiles;      // tiles k_shade resolves from their key page, as x | y << 10 (count in counters[14]) [n_coarse]
    uint32_t *empty_tiles;      // tiles with empty lists as x | y << 10 [n_coarse]
    unsigned long long *scan_desc; // k_setup chained-scan descriptors [ceil(n_triangles / 256)]
    uint2 *clip_queue;          // (triangle, first reserved slot | NO_SLOT) of triangles to clip [n_triangles]
    uint32_t rec_cap, refs_cap;
    uint32_t *tile_cycles;      // debug: SM cycles spent by each coarse tile's CTA (null = off) [n_coarse]
    uint4 *trace;               // debug: one record per CTA of every frame kernel (device_math.cuh: CtaTrace), null = off
    uint32_t *trace_count;      // records written (may exceed trace_cap: the excess is dropped)
    uint32_t trace_cap, trace_tag; // tag = work set of the frame
};

enum : uint32_t { OVERFLOW_RECORDS = 1u, OVERFLOW_REFS = 2u };

constexpr int N_FRAME_KERNELS = 10; // k_vertex k_setup k_clip k_bin<count> k_alloc k_bin<fill> k_raster k_clear_empty k_tile k_shade

} // namespace drawb200
