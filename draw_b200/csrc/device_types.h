// device_types.h — plain structs shared by the host library (scene.cpp) and the kernels.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <vector_types.h>

namespace drawb200 {

// Screen partition.  A CTA of k_tile owns one tile of TILE_W x TILE_H pixels (or a window of one) and keeps
// its depth / winner / colour on chip.  A (record, tile) reference is classed by the area of the record's
// bbox inside the tile:
//   large   > MEDIUM_AREA px : per-tile list (FrameDev::list_refs); k_tile, every lane tests its own pixels
//   medium  <= MEDIUM_AREA px: frame-wide list m_refs; k_raster, one reference per warp, lane per pixel
//   small   <= SMALL_AREA px : frame-wide list s_refs; k_raster, one reference per lane
// Medium and small fragments go to the tile's key page (64-bit atomicMin in L2); k_tile starts from the page.
// Transparent records have per-tile lists of their own (t_refs), consumed in draw order by k_tile.
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
#define DRAW_SMALL_AREA 16
#endif
constexpr int SMALL_AREA = DRAW_SMALL_AREA, MEDIUM_AREA = DRAW_MEDIUM_AREA;
constexpr uint32_t NO_SLOT = 0xFFFFFFFFu;
// k_mirror's granularity: a tile is TILE_H / 8 strips of TILE_W x 8 pixels (whole rows of the tile: 256-byte PCIe writes)
constexpr int TILE_STRIP_H = 8, TILE_STRIPS = TILE_H / TILE_STRIP_H;
constexpr uint32_t TILE_STRIPS_ALL = (1u << TILE_STRIPS) - 1u;
static_assert(TILE_STRIPS <= 8, "one state byte per tile");
// Opaque records: slot number = block * SLOT_STRIDE + number inside k_front's 128-triangle block (at most 4 records per
// triangle); storage index = FrameDev::block_loc[block] + number inside the block (record_index, device_math.cuh).
constexpr uint32_t SLOT_SHIFT = 9, SLOT_STRIDE = 1u << SLOT_SHIFT;
constexpr unsigned long long KEY_EMPTY = ~0ull;
#ifndef DRAW_TILE_THREADS
#define DRAW_TILE_THREADS 256
#endif
constexpr int TILE_THREADS = DRAW_TILE_THREADS;
// k_tile phase A geometry: a warp owns a REGION x REGION_H rectangle (4 lanes across, 8 down), a lane
// a 4 x BLK_H block of it.
constexpr int BLK_H = TILE_W * TILE_H / TILE_THREADS / 4;
constexpr int REGION_H = 8 * BLK_H;

// k_tile work items, built by k_front (heaviest first).  An item is a tile plus a pixel window of it in
// units of warp regions (a 4x4 grid of REGION x REGION_H rectangles): the whole tile, or one of the
// 2 / 4 / 8 / 16 windows a dense tile is cut into.
//   bits 0-9 tile x | 10-20 tile y | 21-22 window x0 | 23-24 window y0 | 25-26 window w-1 | 27-28 window h-1
// Tiles with nothing to draw are not items: they are listed in FrameDev::empty_tiles (as x | y << 10,
// count in counters[CNT_EMPTY]) and written by k_tile's CTAs between their items.
constexpr int REGIONS_X = TILE_W / REGION, REGIONS_Y = TILE_H / REGION_H;
constexpr bool TILE_SPLITTABLE = REGIONS_X == 4 && (REGIONS_Y == 4 || REGIONS_Y == 2);
#ifndef DRAW_TILE_MAX_SPLIT
#define DRAW_TILE_MAX_SPLIT 8
#endif
#ifndef DRAW_TILE_SPLIT_MIN_COST
#define DRAW_TILE_SPLIT_MIN_COST 256
#endif
constexpr int TILE_MAX_SPLIT = DRAW_TILE_MAX_SPLIT < REGIONS_X * REGIONS_Y ? DRAW_TILE_MAX_SPLIT : REGIONS_X * REGIONS_Y; // 1, 2, 4, 8 or 16
constexpr int TILE_SPLIT_MIN_COST = DRAW_TILE_SPLIT_MIN_COST; // windows are not made cheaper than this (k_front cost units)
constexpr int TILE_EXTRA_ITEMS = 1024;                        // work-list slots beyond one per tile
#ifndef DRAW_TILE_SPLIT_DIV
#define DRAW_TILE_SPLIT_DIV 296
#endif
constexpr int TILE_SPLIT_DIV = DRAW_TILE_SPLIT_DIV;           // a window should cost about total / this (<= TILE_EXTRA_ITEMS)
constexpr uint32_t ITEM_NONE = 0xFFFFFFFFu;
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
    uint32_t tiles_x, tiles_y;         // whole canvas, in tiles
    // Tile rows rendered by this launch (sort-first partition): rows ty in [tile_y_begin, tile_y_end) with
    // ty % row_step == row_phase.  The whole canvas is (0, tiles_y, 1, 0).
    uint32_t tile_y_begin, tile_y_end, row_step, row_phase;
    uint32_t n_coarse;                 // tiles_x * tiles_y
    uint32_t split_min_cost, split_div, split_max; // k_front's tile splitting policy (defaults: TILE_SPLIT_*)
    uint32_t bar_base;                 // value of the work set's grid-barrier counter when k_front starts (host-tracked)
    uint32_t empty_tile_color;         // 0: tiles nothing is binned to keep their colour bytes (the caller has cleared the buffer: draw_canvas_set_empty_tile_color)
    uint32_t sort_large;               // k_tile: a tile's large references are tested nearest first when there are at least this many (0: never)
    uint32_t clear_first;              // k_tile: a CTA writes its share of the empty tiles before its raster item (else after it)
    uint8_t *tile_state;               // canvas-owned, one byte per tile: k_tile's mask of the 64x8 strips that hold something else than the clear colour (k_mirror.cu)
    uint32_t *status_host;             // pinned host memory of the canvas (device-mapped): k_tile copies counters[0..31] there
    uint8_t *color;                    // the canvas: BGRA8, row 0 = top (canvas.rs:955-956)
    float *depth;                      // depth buffer, row 0 = y 0 (canvas.rs:413-423)
};
#if defined(__CUDACC__)
__host__ __device__
#endif
inline bool row_is_mine(const FrameUniforms &U, uint32_t ty) {
    return ty >= U.tile_y_begin && ty < U.tile_y_end && ty % U.row_step == U.row_phase;
}
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t rows_mine(const FrameUniforms &U) { // number of tile rows this launch renders
    uint32_t first = U.tile_y_begin + (U.row_phase + U.row_step - U.tile_y_begin % U.row_step) % U.row_step;
    return first < U.tile_y_end ? (U.tile_y_end - 1 - first) / U.row_step + 1 : 0;
}

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

// Scene geometry, all objects concatenated (indices are global after upload).  Positions exist twice: as SoA
// streams for the per-vertex pass (coalesced 128-bit loads, four vertices per thread) and packed as float4 for
// the per-triangle gathers (one 16-byte load = one 32-byte sector per corner instead of three).  Normals and
// uvs are only ever gathered: packed only.
struct SceneDev {
    const float *px, *py, *pz;       // positions, SoA                [n_vertices]
    const float4 *pos4;              // positions, packed (x, y, z, 0) [n_vertices]
    const float4 *nrm4;              // normals, packed (x, y, z, 0)   [n_normals]
    const float2 *uv2;               // uv                             [n_uvs]
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
    float4 *vA;                 // per vertex: screen x, y after the divide (scene/mod.rs:1047-1058, before the canvas
                                // offset), depth (:924), plane-side flags as bits (2 per plane: bit 2p = f>0, bit 2p+1 = f<=0)
    float4 *vLH;                // per vertex, 2 x float4 = one 32-byte sector: light xyz (:922), halfway xyz (:925), 0, 0
    RasterRec *rrec;            // opaque records, slots in draw order [rec_cap]
    PrepRec *prep;              // the same records prepared for rasterisation [rec_cap]
    ShadeRec *srec;
    RasterRec *t_rrec;          // transparent records, slot = 4*ordinal + k, in draw order [4*n_transparent]
    PrepRec *t_prep;
    ShadeRec *t_srec;
    // per tile [n_coarse]
    uint32_t *l_count, *t_count;   // large / transparent references: count (k_front P1), then fill cursor (k_raster), then count again
    uint32_t *l_offset, *t_offset; // first entry of the tile's list in list_refs / t_refs
    uint32_t *ms_weight;           // medium / small references of the tile, weighed by their 8x4 blocks: non-zero = k_raster
                                   // may have put fragments in the tile's key page
    // reference lists
    uint2 *l_pairs, *t_pairs;   // (tile, slot) pairs appended by k_front, scattered into the tiles' lists by k_raster's prologue [refs_cap]
    uint32_t *list_refs;        // large lists: record slots [refs_cap]
    uint32_t *t_refs;           // transparent lists: ordered slots 4*ordinal + k, unordered inside a tile's list [refs_cap]
    uint4 *huge_jobs;           // records covering more than k_front's HUGE_TILES tiles: (bbx, bby, slot | transparent << 31, first tile row in the frame-wide row space)
    uint32_t huge_cap;
    uint2 *m_refs, *s_refs;     // medium / small references of the whole frame: (record slot, tile x | y << 10 | part bits) [refs_cap each]
    unsigned long long *key_pages; // page of tile t = TILE_W * TILE_H keys (depth key << 32 | slot) at t * TILE_W * TILE_H, all ones =
                                   // empty; k_raster fills them with atomicMin, k_tile consumes and resets them
    uint32_t *counters;         // CNT_* below [N_COUNTERS]
    uint32_t *tile_order;       // k_tile work items (make_item), bucketed by cost: bucket b's items are tile_order[b * bucket_cap ..
                                // + counters[CNT_BUCKETS + b]), bucket 0 the heaviest [COST_BUCKETS * bucket_cap]
    uint32_t bucket_cap;        // tiles + TILE_EXTRA_ITEMS: any one bucket can hold every item
    uint32_t *empty_tiles;      // tiles with nothing to draw as x | y << 10 [n_coarse]
    uint32_t *block_loc;        // where the records of k_front's 128-triangle block b start in rrec / prep / srec [ceil(n_triangles / 128)]
    uint32_t rec_cap, refs_cap;
    uint32_t *tile_cycles;      // debug: SM cycles spent by each tile's CTA (null = off) [n_coarse]
    uint4 *trace;               // debug: one record per CTA of every frame kernel (device_math.cuh: CtaTrace), null = off
    uint32_t *trace_count;      // records written (may exceed trace_cap: the excess is dropped)
    uint32_t trace_cap, trace_tag; // tag = work set of the frame
};
// FrameDev::counters.  Words 0..31 are posted to the canvas' pinned status block by k_tile.
enum : int {
    CNT_RECORDS = 0, CNT_REFS_NEEDED = 1, CNT_OVERFLOW = 2, CNT_TICKET = 3, CNT_NONEMPTY = 4, CNT_L_PAIRS = 5, CNT_T_PAIRS = 6,
    CNT_L_CURSOR = 7, CNT_T_CURSOR = 8, CNT_MEDIUM = 9, CNT_SMALL = 10, CNT_EMPTY = 13, CNT_ITEMS = 15,
    CNT_HUGE = 30,      // 64-bit (8-byte aligned): huge records queued << 32 | their tile rows
    CNT_PHASE_NS = 16,  // 6 words: global-timer stamps of k_front's phases (block 0), low 32 bits; + 8: durations of the first block's sub-phases
    CNT_BUCKETS = 32,   // COST_BUCKETS words
    CNT_MIRROR_STRIPS = 29, // status word only: strips k_mirror copied to the host mirror (posted by k_mirror, after k_tile's block)
    CNT_ITEM_CURSOR = 96, // k_tile's item cursor, in a 128-byte line of its own
    CNT_BARRIER = 128,    // k_front's grid barrier, in a line of its own; never reset (FrameUniforms::bar_base)
    N_COUNTERS = 160, N_STATUS_WORDS = 32
};
constexpr int COST_BUCKETS = 34;

enum : uint32_t { OVERFLOW_RECORDS = 1u, OVERFLOW_REFS = 2u, OVERFLOW_STALL = 4u, OVERFLOW_HUGE = 8u }; // STALL: a grid barrier of k_front timed out (never in a cooperative launch)

constexpr int N_FRAME_KERNELS = 4; // k_sort_transparent k_front k_raster k_tile

// Canvas::draw_triangle batches (k_overlay.cu): one draw command's triangles, texture and clipping rectangle.
constexpr int OVERLAY_BIN = 64;             // pixels per bin edge
constexpr size_t OVERLAY_REC_BYTES = 80;    // per triangle, written by k_overlay_setup
constexpr uint32_t OVERLAY_MAX_BATCH = 32768; // triangles per launch pair (sizes the bin masks)
struct OverlayClip {
    unsigned long long c[4]; // x0, y0, x1, y1 as given to Rectangle::from_coords
    unsigned long long has;  // 0: Option::None (the whole screen)
};
struct OverlayParams {
    const void *verts; // 3 draw_vertex2d per triangle
    void *recs;        // n x OVERLAY_REC_BYTES
    uint32_t *masks;   // [bins_x * bins_y][words]: bit i of a bin = triangle i's rectangle touches it
    uint32_t *bin_any; // [bins_x * bins_y]
    uint32_t n, words, bins_x, bins_y;
    uint32_t width, height;
    // the draw commands of the submission: command c owns the triangles [cmd_first[c], cmd_first[c + 1]) (numbered over the
    // whole submission) and clips them to cmd_clip[c]
    const unsigned long long *cmd_first; // [n_cmds + 1]
    const OverlayClip *cmd_clip;         // [n_cmds]
    uint32_t n_cmds;
    uint32_t first;                      // number of this batch's first triangle in the submission
    const void *texels; // RGBA8, row 0 = top
    uint32_t tex_w, tex_h;
    uint32_t *color;
    float *depth;
    uint32_t depth_update;
};

} // namespace drawb200
