//! Raw bindings to `include/draw_b200.h` (C ABI version 200), one `extern "C"` item per exported symbol.
//!
//! The reference (`mororo18/draw`) has no FFI of its own; this crate is what its maintainer would add so that
//! `Application` (src/app/mod.rs:44-45, 65-84, 196, 200-202) can drive the GPU renderer.  The safe wrapper with
//! the reference's own type names (`Scene`, `Canvas`, `Camera`, `Object`) is the `draw-b200` crate next to this one.
//! tests/test_abi_cpu.py checks that every function the header declares is bound here.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_float, c_int, c_void};

pub const DRAW_B200_VERSION: c_int = 200;

pub const DRAW_OK: c_int = 0;
pub const DRAW_ERR_INVALID_ARGUMENT: c_int = -1;
pub const DRAW_ERR_NO_DEVICE: c_int = -2;
pub const DRAW_ERR_CUDA: c_int = -3;
pub const DRAW_ERR_OUT_OF_MEMORY: c_int = -4;
pub const DRAW_ERR_IO: c_int = -5;
pub const DRAW_ERR_INTERNAL: c_int = -6;

/// Camera::move_* (scene/mod.rs:381-405)
pub const DRAW_CAMERA_UP: c_int = 0;
pub const DRAW_CAMERA_DOWN: c_int = 1;
pub const DRAW_CAMERA_LEFT: c_int = 2;
pub const DRAW_CAMERA_RIGHT: c_int = 3;
pub const DRAW_CAMERA_FOWARD: c_int = 4;
pub const DRAW_CAMERA_BACKWARD: c_int = 5;

/// Scene (scene/mod.rs:749-757), opaque.
#[repr(C)]
pub struct draw_scene {
    _private: [u8; 0],
}
/// Canvas (canvas.rs:353-363), opaque.
#[repr(C)]
pub struct draw_canvas {
    _private: [u8; 0],
}
/// Object (object.rs:18-31) as produced by the library's loader, opaque.
#[repr(C)]
pub struct draw_object {
    _private: [u8; 0],
}

/// Texture as Canvas::draw_triangle reads it (map_kd, RGBA8, device resident), opaque.
#[repr(C)]
pub struct draw_texture {
    _private: [u8; 0],
}
/// VertexSimpleAttributes (canvas.rs:185-191).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct draw_vertex2d {
    pub x: f32,
    pub y: f32,
    pub u: f32,
    pub v: f32,
    pub r: u8,
    pub g: u8,
    pub b: u8,
    pub pad: u8,
    pub alpha: f32,
}
/// The arguments of Rectangle::from_coords (canvas.rs:315-330).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct draw_rect {
    pub x0: u64,
    pub y0: u64,
    pub x1: u64,
    pub y1: u64,
}
/// One draw command of Gui::render (src/app/gui.rs:397-481): a run of triangles and its clipping rectangle.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct draw_command2d {
    pub n_triangles: usize,
    pub has_clip: c_int,
    pub clip: draw_rect,
}
/// TextureMap (scene/mod.rs:102-110); `pixels == NULL` is TextureMap::default() (1x1x3 white).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct draw_texture_map {
    pub pixels: *const u8,
    pub width: u32,
    pub height: u32,
    pub components: u32,
}

/// Texture (scene/mod.rs:206-216).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct draw_material {
    pub name: *const c_char,
    pub ka: [c_float; 3],
    pub kd: [c_float; 3],
    pub ks: [c_float; 3],
    pub alpha: c_float,
    pub map_ka: draw_texture_map,
    pub map_kd: draw_texture_map,
}

/// IndexedMesh (mesh.rs:31-35): 9 indices per triangle, (v0 v1 v2, t0 t1 t2, n0 n1 n2).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct draw_mesh {
    pub name: *const c_char,
    pub triangles: *const u32,
    pub n_triangles: usize,
    pub material_idx: u32,
}

/// The arguments of Object::new (object.rs:34-41).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct draw_object_desc {
    pub name: *const c_char,
    pub positions: *const c_float,
    pub n_positions: usize,
    pub normals: *const c_float,
    pub n_normals: usize,
    pub uvs: *const c_float,
    pub n_uvs: usize,
    pub meshes: *const draw_mesh,
    pub n_meshes: usize,
    pub materials: *const draw_material,
    pub n_materials: usize,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct draw_frame_stats {
    pub input_triangles: u32,
    pub setup_records: u32,
    pub tile_refs: u32,
    pub large_refs: u32,
    pub medium_refs: u32,
    pub small_refs: u32,
    pub transparent_refs: u32,
    pub overflow: u32,
    pub empty_tiles: u32,
    pub work_items: u32,
    pub mirror_kbytes: u32,
    pub front_phase_ns: [u32; 7],
    pub front_block_ns: [u32; 5],
}

/// Decodes one texture file for `draw_object_load_obj`: returns 0 and a malloc()ed buffer of
/// width*height*components bytes (components 3 or 4, row 0 = top) that the library free()s.
pub type draw_image_loader = Option<
    unsafe extern "C" fn(
        path: *const c_char,
        user: *mut c_void,
        out_pixels: *mut *mut u8,
        out_w: *mut u32,
        out_h: *mut u32,
        out_components: *mut u32,
    ) -> c_int,
>;

extern "C" {
    // ---- library
    pub fn draw_version() -> c_int;
    pub fn draw_last_error() -> *const c_char;
    pub fn draw_device_count(out_count: *mut c_int) -> c_int;
    pub fn draw_set_device(device: c_int) -> c_int;

    // ---- Scene (scene/mod.rs)
    pub fn draw_scene_create(width: usize, height: usize, out: *mut *mut draw_scene) -> c_int; // Scene::new :760
    pub fn draw_scene_destroy(scene: *mut draw_scene);
    pub fn draw_scene_add_object(scene: *mut draw_scene, desc: *const draw_object_desc, out_id: *mut u32) -> c_int; // Object::new + add_obj :788
    pub fn draw_scene_set_camera(scene: *mut draw_scene, pos: *const c_float, dir: *const c_float) -> c_int; // Camera::new :297
    pub fn draw_scene_get_camera(scene: *const draw_scene, pos: *mut c_float, dir: *mut c_float) -> c_int;
    pub fn draw_scene_set_camera_pos(scene: *mut draw_scene, pos: *const c_float) -> c_int; // Camera::set_pos :377
    pub fn draw_scene_camera_move(scene: *mut draw_scene, dir: c_int, dist: c_float) -> c_int; // Camera::move_* :381-405
    pub fn draw_scene_move_camera_direction(scene: *mut draw_scene, dx: c_int, dy: c_int) -> c_int; // :803
    pub fn draw_scene_set_light(scene: *mut draw_scene, pos: *const c_float) -> c_int;
    pub fn draw_scene_render(scene: *mut draw_scene, canvas: *mut draw_canvas) -> c_int; // Scene::render :901
    pub fn draw_scene_prepare(scene: *mut draw_scene, canvas: *mut draw_canvas) -> c_int;
    pub fn draw_scene_get_uniforms(scene: *mut draw_scene, matrix: *mut c_float, planes: *mut c_float) -> c_int;
    pub fn draw_scene_read_vertex_visual(scene: *mut draw_scene, canvas: *mut draw_canvas, first: usize, count: usize, out: *mut c_float) -> c_int;
    pub fn draw_scene_counts(scene: *const draw_scene, n_objects: *mut usize, n_triangles: *mut usize, n_vertices: *mut usize) -> c_int;
    pub fn draw_scene_launch_count(scene: *const draw_scene, out: *mut u64) -> c_int;
    pub fn draw_scene_set_kernel_timing(scene: *mut draw_scene, enabled: c_int) -> c_int;
    pub fn draw_scene_last_kernel_times(scene: *mut draw_scene, canvas: *mut draw_canvas, ms: *mut c_float) -> c_int;
    pub fn draw_scene_debug_list_counts(scene: *mut draw_scene, canvas: *mut draw_canvas, out: *mut u32, n: usize, n_coarse: *mut usize) -> c_int;
    pub fn draw_scene_debug_tile_cycles(scene: *mut draw_scene, canvas: *mut draw_canvas, enable: c_int, out: *mut u32, n: usize) -> c_int;
    pub fn draw_scene_debug_trace(scene: *mut draw_scene, enable: c_int, out: *mut u32, cap_records: usize, n_records: *mut usize) -> c_int;

    // ---- Canvas (canvas.rs)
    pub fn draw_canvas_create(width: usize, height: usize, out: *mut *mut draw_canvas) -> c_int; // Canvas::new :366
    pub fn draw_canvas_destroy(canvas: *mut draw_canvas);
    pub fn draw_canvas_init_depth(canvas: *mut draw_canvas, depth: c_float) -> c_int; // :403
    pub fn draw_canvas_apply_offset(canvas: *mut draw_canvas, x: c_int, y: c_int) -> c_int; // :382
    pub fn draw_canvas_resize(canvas: *mut draw_canvas, width: usize, height: usize) -> c_int; // :387
    pub fn draw_canvas_clear(canvas: *mut draw_canvas) -> c_int; // :425
    pub fn draw_canvas_enable_depth_update(canvas: *mut draw_canvas) -> c_int; // :399
    pub fn draw_canvas_disable_depth_update(canvas: *mut draw_canvas) -> c_int; // :395
    pub fn draw_canvas_size(canvas: *const draw_canvas, width: *mut usize, height: *mut usize) -> c_int;
    pub fn draw_canvas_map_host(canvas: *mut draw_canvas, out_bytes: *mut *const u8, out_len: *mut usize) -> c_int; // as_bytes_slice :974
    pub fn draw_canvas_enable_host_mirror(canvas: *mut draw_canvas, enabled: c_int) -> c_int;
    pub fn draw_canvas_read_depth(canvas: *mut draw_canvas, dst: *mut c_float, n_floats: usize) -> c_int; // get_pixel_depth :413
    pub fn draw_canvas_sync(canvas: *mut draw_canvas) -> c_int;
    pub fn draw_canvas_last_frame_stats(canvas: *mut draw_canvas, out: *mut draw_frame_stats) -> c_int;

    // ---- Canvas::draw_triangle (canvas.rs:435-575)
    pub fn draw_texture_create(map_kd: *const draw_texture_map, out: *mut *mut draw_texture) -> c_int;
    pub fn draw_texture_destroy(texture: *mut draw_texture);
    pub fn draw_canvas_draw_triangles(canvas: *mut draw_canvas, vertices: *const draw_vertex2d, n_triangles: usize,
                                      texture: *const draw_texture, clipping_rect: *const draw_rect) -> c_int;
    pub fn draw_canvas_draw_commands(canvas: *mut draw_canvas, vertices: *const draw_vertex2d, n_triangles: usize,
                                     commands: *const draw_command2d, n_commands: usize, texture: *const draw_texture) -> c_int;

    // ---- device-side plumbing (multi-GPU drivers)
    pub fn draw_canvas_device_ptrs(canvas: *mut draw_canvas, out_color: *mut *mut c_void, out_depth: *mut *mut c_void) -> c_int;
    pub fn draw_canvas_bind_external(canvas: *mut draw_canvas, color_dev: *mut c_void, depth_dev: *mut c_void) -> c_int;
    pub fn draw_canvas_export_png(canvas: *mut draw_canvas, path: *const c_char) -> c_int; // export_frame_as(Png), app/mod.rs:316-360
    pub fn draw_canvas_export_jpeg(canvas: *mut draw_canvas, path: *const c_char) -> c_int; // export_frame_as(Jpeg)
    pub fn draw_canvas_set_stream(canvas: *mut draw_canvas, cuda_stream: *mut c_void) -> c_int;
    pub fn draw_canvas_stream_wait(canvas: *mut draw_canvas, cuda_stream: *mut c_void) -> c_int;
    pub fn draw_canvas_set_stripe(canvas: *mut draw_canvas, y0: usize, y1: usize) -> c_int;
    pub fn draw_canvas_set_tile_rows(canvas: *mut draw_canvas, phase: u32, step: u32) -> c_int;
    pub fn draw_canvas_set_empty_tile_color(canvas: *mut draw_canvas, enabled: c_int) -> c_int;
    pub fn draw_tile_size() -> c_int;
    pub fn draw_canvas_ipc_export(canvas: *mut draw_canvas, handle: *mut u8) -> c_int;
    pub fn draw_ipc_open(handle: *const u8, out_dev_ptr: *mut *mut c_void) -> c_int;
    pub fn draw_ipc_close(dev_ptr: *mut c_void) -> c_int;
    pub fn draw_device_alloc(bytes: usize, out_dev_ptr: *mut *mut c_void) -> c_int;
    pub fn draw_device_free(dev_ptr: *mut c_void) -> c_int;
    pub fn draw_ipc_export(dev_ptr: *mut c_void, handle: *mut u8) -> c_int;
    pub fn draw_flag_signal(flag_dev: *mut c_void, value: u32, canvas: *mut draw_canvas) -> c_int;
    pub fn draw_flags_wait(flags_dev: *const c_void, n_flags: u32, value: u32, error_word_dev: *mut c_void, canvas: *mut draw_canvas) -> c_int;

    // ---- Object loader (object.rs:73-454), images (scene/mod.rs:174-202, app/mod.rs:362-378); host only
    pub fn draw_object_load_obj(path: *const c_char, loader: draw_image_loader, user: *mut c_void, out: *mut *mut draw_object) -> c_int; // Object::load_from_file :106
    pub fn draw_image_load(path: *const c_char, out_pixels: *mut *mut u8, out_w: *mut u32, out_h: *mut u32, out_components: *mut u32) -> c_int;
    pub fn draw_image_free(pixels: *mut u8);
    pub fn draw_image_write_png(path: *const c_char, pixels: *const u8, width: u32, height: u32, components: u32) -> c_int;
    pub fn draw_image_write_jpg(path: *const c_char, pixels: *const u8, width: u32, height: u32, components: u32, quality: c_int) -> c_int;
    pub fn draw_image_loader_builtin(path: *const c_char, user: *mut c_void, out_pixels: *mut *mut u8, out_w: *mut u32, out_h: *mut u32, out_components: *mut u32) -> c_int;
    pub fn draw_object_free(obj: *mut draw_object);
    pub fn draw_object_desc_of(obj: *const draw_object, out: *mut draw_object_desc) -> c_int;
}
