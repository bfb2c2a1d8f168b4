/*
 * draw_b200.h — C ABI of the B200-native renderer (libdraw_b200.so).
 *
 * This is the drop-in boundary for the one hot path of mororo18/draw: everything under
 * `Scene::render(&mut Canvas)` (src/renderer/scene/mod.rs:901) up to the bytes returned by
 * `Canvas::as_bytes_slice()` (src/renderer/canvas.rs:974).  The reference has no FFI of its
 * own; the seam is the pair of public Rust types `Scene` / `Canvas` used by
 * `Application` (src/app/mod.rs:44-45,65-84,196-202).  Each entry point below names the
 * reference item it replaces.  INTEGRATION.md shows the Rust `extern "C"` block a maintainer
 * would add to bind them.
 *
 * Conventions
 *   - Plain C: opaque handles, pointers and sizes.  No C++ or torch types cross the boundary.
 *   - Every function returns a draw_status (0 = ok, negative = error) unless noted; on error
 *     `draw_last_error()` returns a thread-local message.  The reference panics instead
 *     (assert!/unwrap/expect); no exception or abort crosses this boundary.
 *   - Inputs are borrowed for the duration of the call and copied (to device memory where
 *     needed); the caller keeps ownership of every array it passes in.
 *   - Handles are not thread-safe (the reference is single-threaded, `&mut` everywhere);
 *     distinct handles may be used from distinct threads.
 *   - A scene and a canvas live on the CUDA device that was current when they were created
 *     and must be used together on that device.  There is no CPU fallback: every compute
 *     entry point fails with DRAW_ERR_NO_DEVICE when no CUDA device is usable.
 *   - float arrays are IEEE binary32; positions / normals / uvs are 3 floats per element
 *     (uv as the reference's Vec3, z is ignored by the renderer, object.rs:151-155).
 */
#ifndef DRAW_B200_H
#define DRAW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRAW_B200_VERSION 200 /* 0.2.0 */

typedef enum draw_status {
    DRAW_OK = 0,
    DRAW_ERR_INVALID_ARGUMENT = -1,
    DRAW_ERR_NO_DEVICE = -2,
    DRAW_ERR_CUDA = -3,
    DRAW_ERR_OUT_OF_MEMORY = -4,
    DRAW_ERR_IO = -5,
    DRAW_ERR_INTERNAL = -6
} draw_status;

typedef struct draw_scene draw_scene;   /* Scene  (scene/mod.rs:749-757) */
typedef struct draw_canvas draw_canvas; /* Canvas (canvas.rs:353-363)    */
typedef struct draw_object draw_object; /* Object (object.rs:18-31), host-side, from the loader */

/* TextureMap (scene/mod.rs:102-110).  pixels == NULL means TextureMap::default(): 1x1x3 white
 * (scene/mod.rs:128-135).  components is 3 or 4; row 0 is the top of the image. */
typedef struct draw_texture_map {
    const uint8_t *pixels;
    uint32_t width, height, components;
} draw_texture_map;

/* Texture (scene/mod.rs:206-216): Phong coefficients, opacity and the two maps. */
typedef struct draw_material {
    const char *name; /* may be NULL */
    float ka[3], kd[3], ks[3];
    float alpha; /* < 1 puts the mesh in the transparent pass (object.rs:45-53) */
    draw_texture_map map_ka, map_kd;
} draw_material;

/* IndexedMesh (mesh.rs:31-35).  triangles holds 9 indices per triangle:
 * (v0 v1 v2, t0 t1 t2, n0 n1 n2) = the reference's (vertex, texture, normal) index triples. */
typedef struct draw_mesh {
    const char *name; /* may be NULL */
    const uint32_t *triangles;
    size_t n_triangles;
    uint32_t material_idx; /* index into draw_object_desc.materials (IndexedMesh.texture_idx) */
} draw_mesh;

/* The arguments of Object::new (object.rs:34-41). */
typedef struct draw_object_desc {
    const char *name; /* may be NULL */
    const float *positions; size_t n_positions;
    const float *normals;   size_t n_normals;
    const float *uvs;       size_t n_uvs;
    const draw_mesh *meshes;         size_t n_meshes;
    const draw_material *materials;  size_t n_materials;
} draw_object_desc;

/* Camera::move_* (scene/mod.rs:381-405) */
typedef enum draw_camera_dir {
    DRAW_CAMERA_UP = 0, DRAW_CAMERA_DOWN = 1, DRAW_CAMERA_LEFT = 2,
    DRAW_CAMERA_RIGHT = 3, DRAW_CAMERA_FOWARD = 4, DRAW_CAMERA_BACKWARD = 5
} draw_camera_dir;

/* Counters of the last completed frame (device-side bookkeeping, read back at sync). */
typedef struct draw_frame_stats {
    uint32_t input_triangles;   /* triangles in the scene's draw list */
    uint32_t setup_records;     /* record slots used: triangles that survived cull / reject / clip / zero-area (4 per clipped one) */
    uint32_t tile_refs;         /* (tile, triangle) pairs produced by binning, all classes */
    uint32_t large_refs, medium_refs, small_refs, transparent_refs; /* by class */
    uint32_t overflow;          /* non-zero: a device buffer was too small, frame was re-rendered */
    uint32_t empty_tiles;       /* tiles of this render's rows nothing was binned to (they only get the clear colour and depth) */
    uint32_t work_items;        /* k_tile work items: non-empty tiles, dense ones cut into windows */
    uint32_t mirror_kbytes;     /* KiB of the frame copied to the pinned host mirror behind the render (draw_canvas_enable_host_mirror):
                                   the whole frame, or the 64x8-pixel strips that changed; 0 without the mirror */
    uint32_t front_phase_ns[7]; /* k_front, CTA 0: vertex phase, barrier, triangle phase, barrier (+ huge-record phase), tile phase;
                                   then the triangle phase of the slowest CTA and the huge-record phase */
    uint32_t front_block_ns[5]; /* k_front, first 256-triangle block: set-up, slot scan, record write, binning, clip path */
} draw_frame_stats;

/* ---- library ------------------------------------------------------------------------ */
int draw_version(void);                 /* returns DRAW_B200_VERSION */
const char *draw_last_error(void);      /* thread-local, never NULL */
int draw_device_count(int *out_count);  /* CUDA devices visible to this process */
int draw_set_device(int device);        /* device used by subsequent *_create calls */

/* ---- Scene (scene/mod.rs) ------------------------------------------------------------ */
/* Scene::new(width, height) :760 — default camera (0,0,150) looking at the origin, light at
 * (0,300,300).  width/height size the viewport transform (:818-819, :889-894). */
int draw_scene_create(size_t width, size_t height, draw_scene **out);
void draw_scene_destroy(draw_scene *scene);
/* Object::new :34 + Scene::add_obj :788.  Copies the geometry to the device.  out_id may be
 * NULL; ids count up from 0 like ObjectInfo.id. */
int draw_scene_add_object(draw_scene *scene, const draw_object_desc *desc, uint32_t *out_id);
/* scene.camera = Camera::new(pos, dir, ratio) :297, ratio = scene width / height (:768). */
int draw_scene_set_camera(draw_scene *scene, const float pos[3], const float dir[3]);
int draw_scene_get_camera(const draw_scene *scene, float pos[3], float dir[3]);
/* Camera::set_pos :377 */
int draw_scene_set_camera_pos(draw_scene *scene, const float pos[3]);
/* Camera::move_up/down/left/right/foward/backward :381-405 */
int draw_scene_camera_move(draw_scene *scene, draw_camera_dir dir, float dist);
/* Scene::move_camera_direction(dx, dy) :803 */
int draw_scene_move_camera_direction(draw_scene *scene, int dx, int dy);
/* light_source is a private field initialised at :766; settable here for completeness */
int draw_scene_set_light(draw_scene *scene, const float pos[3]);
/* Scene::render(&mut Canvas) :901 — THE hot path.  Enqueues the frame on the canvas' stream
 * and returns without waiting; the first host read (map_host / read_depth / sync) waits. */
int draw_scene_render(draw_scene *scene, draw_canvas *canvas);
/* Not in the reference (its first frame allocates nothing): sets up everything later frames of this scene on
 * this canvas' geometry need — every frame-in-flight slot's streams and work buffers, its key pages, the
 * captured frame graph — and waits for it, so that the first draw_scene_render costs what the others do.
 * Optional: the first render after draw_scene_add_object does the same set-up without waiting. */
int draw_scene_prepare(draw_scene *scene, draw_canvas *canvas);
/* Debug / parity taps: matrix_transf (:817-899, row-major 16 floats) and the six view planes
 * (:481-593) as near, far, right, left, top, bottom, each (nx, ny, nz, k). */
int draw_scene_get_uniforms(draw_scene *scene, float matrix[16], float planes[24]);
/* Per-vertex visual info of the last rendered frame (:917-926) for vertices
 * [first, first+count) of the scene's concatenated vertex list: 7 floats per vertex
 * (light3, halfway3, depth).  Waits for the frame. */
int draw_scene_read_vertex_visual(draw_scene *scene, draw_canvas *canvas, size_t first, size_t count,
                                  float *out);
int draw_scene_counts(const draw_scene *scene, size_t *n_objects, size_t *n_triangles,
                      size_t *n_vertices);
/* Kernels launched by this scene so far (bench.py reports the delta over the timed region). */
int draw_scene_launch_count(const draw_scene *scene, uint64_t *out);

/* Measurement tap: when enabled, every frame records CUDA events between its kernels on the frame's
 * stream; last_kernel_times waits for the frame and returns the device time in ms of
 * k_sort_transparent, k_front, k_raster, k_tile (DESIGN.md describes them). */
int draw_scene_set_kernel_timing(draw_scene *scene, int enabled);
int draw_scene_last_kernel_times(draw_scene *scene, draw_canvas *canvas, float ms[4]);

/* Debug tap: per tile of the last frame, three arrays of n_coarse words each: large references, medium / small
 * weight (8x4 blocks of bbox; non-zero = the tile has a key page to merge), transparent references.
 * out == NULL only returns the number of tiles. */
int draw_scene_debug_list_counts(draw_scene *scene, draw_canvas *canvas, uint32_t *out, size_t n, size_t *n_coarse);

/* Debug tap: enable != 0 makes later frames record SM cycles per tile in k_tile; out (n <= 4 * number of tiles)
 * receives the last frame's values — whole item, end of phase A, end of phase C, end of phase D — or NULL to only toggle. */
int draw_scene_debug_tile_cycles(draw_scene *scene, draw_canvas *canvas, int enable, uint32_t *out, size_t n);
/* Debug timeline: while enabled, thread 0 of every CTA of the frame kernels appends one record of four
 * uint32 (kernel id | SM << 8 | work set << 24, CTA index, start ns, end ns; low 32 bits of the GPU's global
 * timer).  A call first copies up to cap_records records gathered so far into out (if not NULL), then
 * enables or disables recording and empties the buffer.  Waits for the device. */
int draw_scene_debug_trace(draw_scene *scene, int enable, uint32_t *out, size_t cap_records, size_t *n_records);

/* ---- Canvas (canvas.rs) -------------------------------------------------------------- */
/* Canvas::new(width, height) :366 — colour BGRA8 (Pixel, :51-59), black; no depth yet. */
int draw_canvas_create(size_t width, size_t height, draw_canvas **out);
void draw_canvas_destroy(draw_canvas *canvas);
int draw_canvas_init_depth(draw_canvas *canvas, float depth);     /* :403 */
int draw_canvas_apply_offset(draw_canvas *canvas, int x, int y);  /* :382 */
int draw_canvas_resize(draw_canvas *canvas, size_t width, size_t height); /* :387 */
int draw_canvas_clear(draw_canvas *canvas);                       /* :425 */
int draw_canvas_enable_depth_update(draw_canvas *canvas);         /* :399 (render sets it itself) */
int draw_canvas_disable_depth_update(draw_canvas *canvas);        /* :395 */
int draw_canvas_size(const draw_canvas *canvas, size_t *width, size_t *height);
/* Canvas::as_bytes_slice / as_ptr / size_bytes :966-982.  Waits for the frame, copies it to a
 * pinned host mirror if it changed, and returns that mirror: width*height*4 bytes, B,G,R,pad
 * per pixel, row 0 = top.  Valid until the next render / resize / destroy on this canvas. */
int draw_canvas_map_host(draw_canvas *canvas, const uint8_t **out_bytes, size_t *out_len);
/* The reference's frame always lives in host memory (Canvas::frame, canvas.rs:353).  With the host mirror enabled every
 * draw_scene_render also brings the pinned mirror up to date behind its kernels, so the PCIe transfer of frame k overlaps
 * the rendering of frame k+1 on another canvas and draw_canvas_map_host only waits.  The refresh is a copy of the whole
 * frame (4*W*H bytes) or — when the frame is a whole-canvas render, most of it is clear colour and the mirror's content is
 * known — of the tiles that differ from what the mirror holds: drawn in this frame, or drawn in the mirror's frame and
 * cleared since (k_mirror.cu, in strips of 64x8 pixels; draw_frame_stats.mirror_kbytes says how much).  Either way the mirror is byte-identical to
 * the device frame.  Off by default. */
int draw_canvas_enable_host_mirror(draw_canvas *canvas, int enabled);
/* depth_frame (get_pixel_depth :413): width*height floats, row index = canvas y (not flipped). */
int draw_canvas_read_depth(draw_canvas *canvas, float *dst, size_t n_floats);
/* Wait for everything enqueued on the canvas' stream. */
int draw_canvas_sync(draw_canvas *canvas);
int draw_canvas_last_frame_stats(draw_canvas *canvas, draw_frame_stats *out);

/* ---- Canvas::draw_triangle (canvas.rs:435-575), the 2-D path of the GUI -------------- */
/* VertexSimpleAttributes (canvas.rs:185-191): screen_coord (pixels, y up like the canvas), texture_coord, color
 * (Color::Custom([r, g, b])) and alpha. */
typedef struct draw_vertex2d {
    float x, y, u, v;
    uint8_t r, g, b, pad;
    float alpha;
} draw_vertex2d;
/* The arguments of Rectangle::from_coords(x0, y0, x1, y1) (canvas.rs:315-330), in canvas pixels. */
typedef struct draw_rect {
    uint64_t x0, y0, x1, y1;
} draw_rect;
/* The `texture: Option<&Texture>` argument reduced to what the path reads, texture.map_kd through get_rgba_slice
 * (scene/mod.rs:137-152): an RGBA8 map (components must be 4), copied to the device once. */
typedef struct draw_texture draw_texture;
int draw_texture_create(const draw_texture_map *map_kd, draw_texture **out);
void draw_texture_destroy(draw_texture *texture);
/* n_triangles calls of Canvas::draw_triangle(a, b, c, Some(texture), clipping_rect) in order (vertices holds 3 per
 * triangle; clipping_rect may be NULL = None) — one draw command of src/app/gui.rs:382-485.  Pixels are written over
 * the canvas' frame with depth 0.0 under the canvas' depth-update switch, blended in submission order; the partition
 * set by draw_canvas_set_stripe / set_tile_rows does not apply.  Waits for a frame still being rendered, then enqueues
 * on the canvas' stream and returns; the vertices have been copied when it returns.  Needs a depth buffer
 * (draw_canvas_init_depth), as the reference's depth test does. */
int draw_canvas_draw_triangles(draw_canvas *canvas, const draw_vertex2d *vertices, size_t n_triangles, const draw_texture *texture,
                               const draw_rect *clipping_rect);
/* One draw command of Gui::render (src/app/gui.rs:397-481, ig::DrawCmd::Elements): n_triangles consecutive triangles of a
 * submission and the clipping rectangle they are drawn with (has_clip == 0: None). */
typedef struct draw_command2d {
    size_t n_triangles;
    int has_clip;
    draw_rect clip;
} draw_command2d;
/* A whole GUI frame at once: the draw commands of src/app/gui.rs:389-485 that share a texture (ImGui's font atlas), in
 * submission order — command k draws the next commands[k].n_triangles triangles of `vertices` with its own clipping
 * rectangle; the counts must add up to n_triangles.  Same result as one draw_canvas_draw_triangles call per command, in two
 * kernel launches per 32 768 triangles instead of two per command. */
int draw_canvas_draw_commands(draw_canvas *canvas, const draw_vertex2d *vertices, size_t n_triangles, const draw_command2d *commands,
                              size_t n_commands, const draw_texture *texture);

/* ---- device-side plumbing (not in the reference; used by the multi-GPU drivers) ------ */
/* Device pointers of the colour (BGRA8, y-flipped rows) and depth (f32) buffers. */
int draw_canvas_device_ptrs(draw_canvas *canvas, void **out_color, void **out_depth);
/* Render into caller-owned device memory instead (e.g. a torch tensor, or a peer GPU's
 * framebuffer opened through CUDA IPC).  NULL restores the canvas' own buffer.  The colour
 * buffer must hold width*height*4 bytes; depth width*height floats. */
int draw_canvas_bind_external(draw_canvas *canvas, void *color_dev, void *depth_dev);
/* Application::export_frame_as(Png) (app/mod.rs:316-360) without the file dialog: waits for the frame, swaps B and R
 * (the frame is BGRA, the file RGBA with the pad byte as alpha) and writes it with draw_image_write_png. */
int draw_canvas_export_png(draw_canvas *canvas, const char *path);
/* Application::export_frame_as(Jpeg): the same swap, then draw_image_write_jpg with the quality the reference ends up
 * passing — width * 4, which stbi_write_jpg clamps to 100 (app/mod.rs:362-378). */
int draw_canvas_export_jpeg(draw_canvas *canvas, const char *path);
/* Enqueue on an existing CUDA stream (a cudaStream_t passed as void*); NULL = own stream. */
int draw_canvas_set_stream(draw_canvas *canvas, void *cuda_stream);
/* Device-side join: makes cuda_stream (a cudaStream_t passed as void*) wait for everything enqueued so far
 * for this canvas (its last render included) without blocking the host.  Lets canvases render on their own
 * streams — frames on different canvases then overlap — while a caller's stream consumes or times them. */
int draw_canvas_stream_wait(draw_canvas *canvas, void *cuda_stream);
/* Sort-first partition: render only canvas rows y in [y0, y1) (canvas y = depth-buffer row;
 * colour row = height-1-y).  Rows outside are left untouched.  (0, height) = whole frame.
 * y0 and y1 must be multiples of the tile height (draw_tile_size, 32) or equal to height. */
int draw_canvas_set_stripe(draw_canvas *canvas, size_t y0, size_t y1);
/* Sort-first partition, interleaved: render only the tile rows ty (rows of draw_tile_size() pixels, counted from
 * canvas y = 0) with ty % step == phase, inside the stripe.  Interleaving balances the ranks when the scene sits
 * mid-screen.  (0, 1) = every row.  Reset by draw_canvas_resize. */
int draw_canvas_set_tile_rows(draw_canvas *canvas, uint32_t phase, uint32_t step);
/* Fused sort-first gather: with enabled == 0 a render leaves the COLOUR of the tiles nothing is drawn in untouched
 * (their depth is still reset) — the owner of the framebuffer has filled it with the clear colour (draw_canvas_clear)
 * before the peers' kernels store into it, so only pixels of tiles that hold geometry cross NVLink.  Default: enabled
 * (Canvas::clear semantics, canvas.rs:425-433: every pixel of the rendered rows is written). */
int draw_canvas_set_empty_tile_color(draw_canvas *canvas, int enabled);
int draw_tile_size(void);
/* Peer access for the fused sort-first gather (one process per GPU): the owner exports its
 * canvas' own colour buffer as a 64-byte CUDA IPC handle; another process on the same node opens
 * it (peer access is enabled lazily) and passes the pointer to draw_canvas_bind_external, so that
 * its tile kernel stores pixels straight into the owner's framebuffer over NVLink. */
int draw_canvas_ipc_export(draw_canvas *canvas, uint8_t handle[64]);
int draw_ipc_open(const uint8_t handle[64], void **out_dev_ptr);
int draw_ipc_close(void *dev_ptr);
/* Device-side completion flags of that gather (no host in the loop, no collective): a small zero-filled device
 * buffer of 32-bit words owned by one rank (draw_device_alloc) and opened by the others (draw_ipc_export /
 * draw_ipc_open).  draw_flag_signal enqueues, behind everything enqueued so far for the canvas, a system-scope
 * release store of `value` into one word (local or a peer's); draw_flags_wait enqueues on the canvas' stream a
 * wait until each of the n_flags (<= 64) consecutive words has reached `value` (words only grow; wrap-safe).  The
 * wait gives up after 5 s and sets *error_word_dev (may be NULL) to 1 instead of hanging the device. */
int draw_device_alloc(size_t bytes, void **out_dev_ptr);
int draw_device_free(void *dev_ptr);
int draw_ipc_export(void *dev_ptr, uint8_t handle[64]);
int draw_flag_signal(void *flag_dev, uint32_t value, draw_canvas *canvas);
int draw_flags_wait(const void *flags_dev, uint32_t n_flags, uint32_t value, void *error_word_dev, draw_canvas *canvas);

/* ---- Object loader (object.rs:73-454), host only ------------------------------------- */
/* Object::load_from_file :106.  Texture images referenced by the MTL are decoded by the
 * caller-supplied callback (the reference uses stb_image, scene/mod.rs:174-202, which is not
 * part of this library); a NULL callback leaves maps at the 1x1 default.  The callback returns 0 on
 * success and hands over a malloc()ed buffer of width*height*components bytes (components 3 or 4, row 0
 * = top); the library free()s it. */
typedef int (*draw_image_loader)(const char *path, void *user, uint8_t **out_pixels, uint32_t *out_w,
                                 uint32_t *out_h, uint32_t *out_components);
int draw_object_load_obj(const char *path, draw_image_loader loader, void *user, draw_object **out);
/* TextureMap::load_from_file (scene/mod.rs:174-202) for the two formats among the reference's assets: decodes a PNG
 * file (RGB, RGBA or palette; 8/16 bits; non-interlaced) or a JPEG file (three components, baseline or progressive
 * Huffman; stb_image's IDCT / upsampling / colour arithmetic, draw_b200/csrc/jpeg_decode.cpp) to
 * width*height*components bytes, components 3 or 4, row 0 = top.  Other formats fail with
 * DRAW_ERR_INVALID_ARGUMENT.  The buffer is malloc()ed; release it with draw_image_free. */
int draw_image_load(const char *path, uint8_t **out_pixels, uint32_t *out_w, uint32_t *out_h, uint32_t *out_components);
void draw_image_free(uint8_t *pixels);
/* stbi_write_png in Application::write_img (app/mod.rs:362-378): writes width*height*components bytes (components 3
 * or 4, row 0 = top) as a PNG file.  Lossless: the file decodes to exactly the bytes given. */
int draw_image_write_png(const char *path, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t components);
/* stbi_write_jpg in Application::write_img: width*height*components bytes (components 3 or 4; a fourth byte is
 * ignored) as a baseline JFIF file.  quality as stbi_write_jpg reads it: 0 means 90, values are clamped to 1..100;
 * only qualities above 90 (no chroma subsampling — all the reference's export reaches) are written, others fail with
 * DRAW_ERR_INVALID_ARGUMENT.  Lossy: see draw_b200/csrc/jpeg_encode.cpp for what is and is not pinned. */
int draw_image_write_jpg(const char *path, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t components, int quality);
/* draw_image_load as a draw_image_loader callback (user is ignored), for draw_object_load_obj. */
int draw_image_loader_builtin(const char *path, void *user, uint8_t **out_pixels, uint32_t *out_w, uint32_t *out_h,
                              uint32_t *out_components);
void draw_object_free(draw_object *obj);
/* Borrow the loaded object as a desc (valid until draw_object_free). */
int draw_object_desc_of(const draw_object *obj, draw_object_desc *out);

#ifdef __cplusplus
}
#endif
#endif /* DRAW_B200_H */
