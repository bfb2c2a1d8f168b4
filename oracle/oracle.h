/*
 * oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference renderer
 * (mororo18/draw, src/renderer/{linalg.rs,canvas.rs,scene/mod.rs,scene/mesh.rs}) used as
 * the parity checker and as the CPU baseline of bench.py.  Nothing under draw_b200/ (the
 * product) may include, link or call it; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path
 *  and its Rust toolchain is absent here, so this oracle is pinned only
 * by (i) a second, independent numpy-float32 restatement (oracle/np_oracle.py) that must
 * agree bit-for-bit on small frames and (ii) known-answer values derived from the
 * reference source.
 */
#ifndef DRAW_ORACLE_H
#define DRAW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;
typedef struct orc_canvas orc_canvas;

/* Texture (scene/mod.rs:206-216) + its two TextureMaps (scene/mod.rs:102-110). */
typedef struct orc_material {
    float ka[3], kd[3], ks[3];
    float alpha;
    const uint8_t *map_ka; /* NULL => TextureMap::default() (1x1x3 white, scene/mod.rs:128) */
    uint32_t map_ka_w, map_ka_h, map_ka_comp;
    const uint8_t *map_kd;
    uint32_t map_kd_w, map_kd_h, map_kd_comp;
} orc_material;

/* One IndexedMesh (mesh.rs:31-35): tris = 9 u32 per triangle, (v0 v1 v2, t0 t1 t2, n0 n1 n2). */
typedef struct orc_mesh {
    const uint32_t *tris;
    size_t n_tris;
    uint32_t texture_idx;
} orc_mesh;

typedef struct orc_stats {
    uint64_t input_tris;     /* triangles entering the per-triangle body */
    uint64_t culled_tris;    /* rejected by back-face cull */
    uint64_t emitted_tris;   /* triangles handed to draw_triangle_with_attributes */
    uint64_t bbox_pixels;    /* pixels visited by the bbox loops */
    uint64_t covered_frags;  /* pixels passing the inside test (incl. overdraw) */
    uint64_t written_frags;  /* fragments passing the depth test */
} orc_stats;

/* Scene::new (scene/mod.rs:760) */
orc_scene *orc_scene_new(size_t width, size_t height);
void orc_scene_free(orc_scene *);
/* Object::new (object.rs:34) + Scene::add_obj (scene/mod.rs:788); arrays are copied.
 * positions/normals/uvs are 3 floats each (uv as Vec3, z ignored by the renderer). */
int orc_scene_add_object(orc_scene *, const float *positions, size_t n_pos,
                         const float *normals, size_t n_nrm, const float *uvs, size_t n_uv,
                         const orc_mesh *meshes, size_t n_meshes,
                         const orc_material *materials, size_t n_materials);
/* scene.camera = Camera::new(pos, dir, ratio) (scene/mod.rs:297), ratio = scene W/H */
void orc_scene_set_camera(orc_scene *, const float pos[3], const float dir[3]);
void orc_scene_set_light(orc_scene *, const float pos[3]);
/* Camera::move_* (scene/mod.rs:381-405): 0 up 1 down 2 left 3 right 4 foward 5 backward */
void orc_scene_camera_move(orc_scene *, int which, float dist);
/* Scene::move_camera_direction (scene/mod.rs:803) */
void orc_scene_move_camera_direction(orc_scene *, int dx, int dy);
void orc_scene_get_camera(orc_scene *, float pos[3], float dir[3]);
/* matrix_transf (row-major 16) and planes near,far,right,left,top,bottom as (nx,ny,nz,k) */
void orc_scene_uniforms(orc_scene *, float m[16], float planes[24]);
/* per-vertex visual info of object obj after the last render: 10 floats/vertex
 * (light3, eye3, halfway3, depth) (scene/mod.rs:256-261) */
int orc_scene_vertex_visual(orc_scene *, size_t obj, float *out, size_t n_vertices);

/* Canvas::new / init_depth / apply_offset / resize / clear (canvas.rs:366-433) */
orc_canvas *orc_canvas_new(size_t width, size_t height);
void orc_canvas_free(orc_canvas *);
void orc_canvas_init_depth(orc_canvas *, float depth);
void orc_canvas_apply_offset(orc_canvas *, int x, int y);
void orc_canvas_resize(orc_canvas *, size_t width, size_t height);
void orc_canvas_clear(orc_canvas *);
const uint8_t *orc_canvas_bytes(orc_canvas *, size_t *len); /* as_bytes_slice (canvas.rs:974) */
const float *orc_canvas_depth(orc_canvas *, size_t *len);
/* winner id per pixel (depth-buffer layout, not y-flipped): draw index of the triangle that
 * last wrote the pixel's colour, 0xFFFFFFFF = none.  Bookkeeping only. */
const uint32_t *orc_canvas_winner(orc_canvas *, size_t *len);

/* VertexSimpleAttributes (canvas.rs:185-191): screen_coord, texture_coord, color (Color::Custom rgb), alpha. */
typedef struct orc_vertex2d {
    float x, y, u, v;
    uint8_t r, g, b, pad;
    float alpha;
} orc_vertex2d;
/* Canvas::draw_triangle (canvas.rs:435-575), the GUI's 2-D path: texture = an RGBA map (Texture::map_kd read with
 * get_rgba_slice, scene/mod.rs:137-152); clip = Rectangle::from_coords(x0, y0, x1, y1) or NULL for None. */
void orc_canvas_draw_triangle(orc_canvas *, const orc_vertex2d v[3], const uint8_t *rgba, uint32_t w, uint32_t h,
                              const uint64_t clip[4]);
/* Canvas::enable_depth_update / disable_depth_update (canvas.rs:395-401) */
void orc_canvas_set_depth_update(orc_canvas *, int enabled);

/* Scene::render (scene/mod.rs:901).  count_stats != 0 also fills the counters. */
void orc_scene_render(orc_scene *, orc_canvas *, int count_stats);
void orc_scene_stats(orc_scene *, orc_stats *out);

#ifdef __cplusplus
}
#endif
#endif
