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
constexpr int SMALL_AREA = 8, MEDIUM_AREA = 1024;
constexpr int LISTS_PER_TILE = 3;
constexpr uint32_t NO_SLOT = 0xFFFFFFFFu;

// Per-frame constants, passed to every kernel by value (__grid_constant__): no upload, no sync.
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
    uint32_t n_vertices, n_triangles, n_transparent;
};

// What the tile kernel needs to rasterise one screen triangle (48 B).
struct alignas(16) RasterRec {
    float ax, ay, bx, by, cx, cy; // snapped vertex centres (canvas.rs:585-587)
    float da, db, dc;             // per-vertex depth (scene/mod.rs:924)
    uint32_t id;                  // draw id = 4 * draw-order index + clip output; NO_SLOT = empty
    uint32_t bbx, bby;            // x_min | x_max << 16, y_min | y_max << 16 (canvas.rs:640-658)
};

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
    RasterRec *rrec;            // opaque records, unordered slots [rec_cap]
    ShadeRec *srec;
    RasterRec *t_rrec;          // transparent records, slot = 4*ordinal + k, in draw order [4*n_transparent]
    ShadeRec *t_srec;
    uint32_t *list_count;       // per list (large, medium, small per tile): count, then fill cursor [n_lists]
    uint32_t *list_offset;      // first entry of each list in list_refs [n_lists + 1]
    uint32_t *list_refs;        // record slots grouped by list [refs_cap]
    uint32_t *counters;         // [0] records  [1] refs  [2] overflow bits  [3] k_setup CTA ticket  [4] k_alloc CTAs done  [5] clip queue length  [8]
    uint32_t *tile_cost;        // estimated k_tile work per tile [n_coarse]
    uint32_t *tile_order;       // tiles of the stripe, heaviest first [n_coarse]
    unsigned long long *scan_desc; // k_setup chained-scan descriptors [ceil(n_triangles / 256)]
    uint2 *clip_queue;          // (triangle, first reserved slot | NO_SLOT) of triangles to clip [n_triangles]
    uint32_t rec_cap, refs_cap;
    uint32_t *tile_cycles;      // debug: SM cycles spent by each coarse tile's CTA (null = off) [n_coarse]
};

enum : uint32_t { OVERFLOW_RECORDS = 1u, OVERFLOW_REFS = 2u };

constexpr int N_FRAME_KERNELS = 7; // k_vertex k_setup k_clip k_bin<count> k_alloc k_bin<fill> k_tile

} // namespace drawb200
