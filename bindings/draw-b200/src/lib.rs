//! Safe wrapper over `draw-b200-sys` that keeps the reference renderer's API (mororo18/draw `src/renderer`):
//! the only change in the application is the `use` line,
//!
//! ```ignore
//! use draw_b200::{Camera, Canvas, Object, Scene};     // was: crate::renderer::{scene::Scene, canvas::Canvas, ...}
//! ```
//!
//! and `Application::run` stays as it is (src/app/mod.rs:80-84, 196, 200-202):
//!
//! ```ignore
//! canvas.init_depth(100000.0);
//! canvas.apply_offset(off_x, off_y);
//! scene.render(&mut canvas);
//! current_frame.copy_from_slice(canvas.as_bytes_slice());
//! ```
//!
//! Where the reference panics (assert!/unwrap/expect, e.g. canvas.rs:914 "Depth not initialized", mesh.rs:46-48,
//! object.rs:46) the C ABI returns a status and a message; this wrapper turns them back into panics so that the
//! application sees the behaviour it was written against.  Not compiled in this repository's image (no Rust
//! toolchain): it is the source a maintainer of the reference adds.
use draw_b200_sys as sys;
use std::ffi::{CStr, CString};
use std::os::raw::c_int;

fn check(rc: c_int) {
    if rc != sys::DRAW_OK {
        let msg = unsafe { CStr::from_ptr(sys::draw_last_error()) }.to_string_lossy().into_owned();
        panic!("draw_b200 error {}: {}", rc, msg);
    }
}

/// linalg.rs:146-151
#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct Vec3 {
    pub x: f32,
    pub y: f32,
    pub z: f32,
}
impl Vec3 {
    pub fn new(x: f32, y: f32, z: f32) -> Self {
        Self { x, y, z }
    }
}

/// scene/mod.rs:102-110; `None` pixels is TextureMap::default().
#[derive(Clone, Default)]
pub struct TextureMap {
    pub img: Vec<u8>,
    pub width: usize,
    pub height: usize,
    pub components: usize,
}
impl TextureMap {
    /// TextureMap::load_from_file (scene/mod.rs:174-202) through the library's decoder.
    pub fn load_from_file(path: &str) -> Self {
        let c = CString::new(path).unwrap();
        let (mut px, mut w, mut h, mut comp) = (std::ptr::null_mut(), 0u32, 0u32, 0u32);
        check(unsafe { sys::draw_image_load(c.as_ptr(), &mut px, &mut w, &mut h, &mut comp) });
        let n = (w * h * comp) as usize;
        let img = unsafe { std::slice::from_raw_parts(px, n) }.to_vec();
        unsafe { sys::draw_image_free(px) };
        Self { img, width: w as usize, height: h as usize, components: comp as usize }
    }
    fn as_sys(&self) -> sys::draw_texture_map {
        sys::draw_texture_map {
            pixels: if self.img.is_empty() { std::ptr::null() } else { self.img.as_ptr() },
            width: self.width as u32,
            height: self.height as u32,
            components: self.components as u32,
        }
    }
}

/// scene/mod.rs:206-216, defaults :243-253.
#[derive(Clone)]
pub struct Texture {
    pub name: String,
    pub map_ka: TextureMap,
    pub map_kd: TextureMap,
    pub ka: Vec3,
    pub kd: Vec3,
    pub ks: Vec3,
    pub alpha: f32,
}
impl Default for Texture {
    fn default() -> Self {
        Self {
            name: String::new(),
            map_ka: TextureMap::default(),
            map_kd: TextureMap::default(),
            ka: Vec3::new(0.9, 0.9, 0.9),
            kd: Vec3::new(0.4, 0.4, 0.4),
            ks: Vec3::new(0.5, 0.5, 0.5),
            alpha: 1.0,
        }
    }
}

/// mesh.rs:13, 31-35
pub type IndexedTriangle = [usize; 3];
#[derive(Clone)]
pub struct IndexedMesh {
    pub name: String,
    pub triangles: Vec<(IndexedTriangle, IndexedTriangle, IndexedTriangle)>, // (vertex, texture, normal)
    pub texture_idx: usize,
}

/// object.rs:18-31
#[derive(Clone)]
pub struct Object {
    pub name: String,
    pub vertices: Vec<Vec3>,
    pub normals_vertices: Vec<Vec3>,
    pub texture_vertices: Vec<Vec3>,
    pub meshes: Vec<IndexedMesh>,
    pub textures: Vec<Texture>,
}
impl Object {
    /// Object::load_from_file (object.rs:106): the library's loader restates the reference's rules (rescale to
    /// radius 100, normalised normals, quad split, dummy uvs, generated smooth normals, material fallbacks).
    pub fn load_from_file(path: &str) -> Self {
        let c = CString::new(path).unwrap();
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::draw_object_load_obj(c.as_ptr(), Some(sys::draw_image_loader_builtin), std::ptr::null_mut(), &mut h) });
        let mut d: sys::draw_object_desc = unsafe { std::mem::zeroed() };
        check(unsafe { sys::draw_object_desc_of(h, &mut d) });
        let vecs = |p: *const f32, n: usize| -> Vec<Vec3> {
            (0..n).map(|i| unsafe { Vec3::new(*p.add(3 * i), *p.add(3 * i + 1), *p.add(3 * i + 2)) }).collect()
        };
        let name_of = |p: *const std::os::raw::c_char| {
            if p.is_null() { String::new() } else { unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned() }
        };
        let map_of = |m: &sys::draw_texture_map| {
            if m.pixels.is_null() {
                TextureMap::default()
            } else {
                let n = (m.width * m.height * m.components) as usize;
                TextureMap { img: unsafe { std::slice::from_raw_parts(m.pixels, n) }.to_vec(), width: m.width as usize, height: m.height as usize, components: m.components as usize }
            }
        };
        let meshes = (0..d.n_meshes)
            .map(|i| {
                let m = unsafe { &*d.meshes.add(i) };
                let t = unsafe { std::slice::from_raw_parts(m.triangles, 9 * m.n_triangles) };
                IndexedMesh {
                    name: name_of(m.name),
                    triangles: t.chunks_exact(9).map(|q| ([q[0] as usize, q[1] as usize, q[2] as usize], [q[3] as usize, q[4] as usize, q[5] as usize], [q[6] as usize, q[7] as usize, q[8] as usize])).collect(),
                    texture_idx: m.material_idx as usize,
                }
            })
            .collect();
        let textures = (0..d.n_materials)
            .map(|i| {
                let m = unsafe { &*d.materials.add(i) };
                Texture {
                    name: name_of(m.name),
                    map_ka: map_of(&m.map_ka),
                    map_kd: map_of(&m.map_kd),
                    ka: Vec3::new(m.ka[0], m.ka[1], m.ka[2]),
                    kd: Vec3::new(m.kd[0], m.kd[1], m.kd[2]),
                    ks: Vec3::new(m.ks[0], m.ks[1], m.ks[2]),
                    alpha: m.alpha,
                }
            })
            .collect();
        let obj = Object {
            name: name_of(d.name),
            vertices: vecs(d.positions, d.n_positions),
            normals_vertices: vecs(d.normals, d.n_normals),
            texture_vertices: vecs(d.uvs, d.n_uvs),
            meshes,
            textures,
        };
        unsafe { sys::draw_object_free(h) };
        obj
    }
}

/// scene/object.rs:12-16
pub struct ObjectInfo {
    pub id: usize,
    pub name: String,
}

/// The scene's camera (scene/mod.rs:282-594).  `scene.camera = Camera::new(pos, dir, ratio)` becomes
/// `scene.set_camera(Camera::new(pos, dir, ratio))`; the move_* verbs act on the scene's camera.
#[derive(Clone, Copy)]
pub struct Camera {
    pub position: Vec3,
    pub direction: Vec3,
}
impl Camera {
    /// Camera::new (scene/mod.rs:297).  The ratio is always the scene's width / height (:768).
    pub fn new(position: Vec3, direction: Vec3, _ratio: f32) -> Self {
        Self { position, direction }
    }
}

pub struct Scene {
    h: *mut sys::draw_scene,
    pub width: usize,
    pub height: usize,
}
impl Scene {
    /// Scene::new (scene/mod.rs:760)
    pub fn new(width: usize, height: usize) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::draw_scene_create(width, height, &mut h) });
        Self { h, width, height }
    }
    /// Scene::add_obj (scene/mod.rs:788).  The object is copied to the device; `obj` is consumed like in the reference.
    pub fn add_obj(&mut self, obj: Object) -> ObjectInfo {
        let name = CString::new(obj.name.clone()).unwrap();
        let flat: Vec<Vec<u32>> = obj
            .meshes
            .iter()
            .map(|m| m.triangles.iter().flat_map(|(v, t, n)| [v[0], v[1], v[2], t[0], t[1], t[2], n[0], n[1], n[2]]).map(|i| i as u32).collect())
            .collect();
        let mesh_names: Vec<CString> = obj.meshes.iter().map(|m| CString::new(m.name.clone()).unwrap()).collect();
        let meshes: Vec<sys::draw_mesh> = obj
            .meshes
            .iter()
            .enumerate()
            .map(|(i, m)| sys::draw_mesh { name: mesh_names[i].as_ptr(), triangles: flat[i].as_ptr(), n_triangles: m.triangles.len(), material_idx: m.texture_idx as u32 })
            .collect();
        let tex_names: Vec<CString> = obj.textures.iter().map(|t| CString::new(t.name.clone()).unwrap()).collect();
        let materials: Vec<sys::draw_material> = obj
            .textures
            .iter()
            .enumerate()
            .map(|(i, t)| sys::draw_material {
                name: tex_names[i].as_ptr(),
                ka: [t.ka.x, t.ka.y, t.ka.z],
                kd: [t.kd.x, t.kd.y, t.kd.z],
                ks: [t.ks.x, t.ks.y, t.ks.z],
                alpha: t.alpha,
                map_ka: t.map_ka.as_sys(),
                map_kd: t.map_kd.as_sys(),
            })
            .collect();
        // Vec3 is #[repr(C)] {f32, f32, f32}: a Vec<Vec3> is a run of 3 floats per element
        let desc = sys::draw_object_desc {
            name: name.as_ptr(),
            positions: obj.vertices.as_ptr().cast(),
            n_positions: obj.vertices.len(),
            normals: obj.normals_vertices.as_ptr().cast(),
            n_normals: obj.normals_vertices.len(),
            uvs: obj.texture_vertices.as_ptr().cast(),
            n_uvs: obj.texture_vertices.len(),
            meshes: meshes.as_ptr(),
            n_meshes: meshes.len(),
            materials: materials.as_ptr(),
            n_materials: materials.len(),
        };
        let mut id = 0u32;
        check(unsafe { sys::draw_scene_add_object(self.h, &desc, &mut id) });
        ObjectInfo { id: id as usize, name: obj.name }
    }
    pub fn set_camera(&mut self, camera: Camera) {
        let (p, d) = ([camera.position.x, camera.position.y, camera.position.z], [camera.direction.x, camera.direction.y, camera.direction.z]);
        check(unsafe { sys::draw_scene_set_camera(self.h, p.as_ptr(), d.as_ptr()) });
    }
    pub fn camera(&self) -> Camera {
        let (mut p, mut d) = ([0f32; 3], [0f32; 3]);
        check(unsafe { sys::draw_scene_get_camera(self.h, p.as_mut_ptr(), d.as_mut_ptr()) });
        Camera { position: Vec3::new(p[0], p[1], p[2]), direction: Vec3::new(d[0], d[1], d[2]) }
    }
    /// Camera::set_pos / move_* (scene/mod.rs:377-405)
    pub fn camera_set_pos(&mut self, pos: Vec3) {
        let p = [pos.x, pos.y, pos.z];
        check(unsafe { sys::draw_scene_set_camera_pos(self.h, p.as_ptr()) });
    }
    pub fn camera_move_up(&mut self, dist: f32) { check(unsafe { sys::draw_scene_camera_move(self.h, sys::DRAW_CAMERA_UP, dist) }) }
    pub fn camera_move_down(&mut self, dist: f32) { check(unsafe { sys::draw_scene_camera_move(self.h, sys::DRAW_CAMERA_DOWN, dist) }) }
    pub fn camera_move_left(&mut self, dist: f32) { check(unsafe { sys::draw_scene_camera_move(self.h, sys::DRAW_CAMERA_LEFT, dist) }) }
    pub fn camera_move_right(&mut self, dist: f32) { check(unsafe { sys::draw_scene_camera_move(self.h, sys::DRAW_CAMERA_RIGHT, dist) }) }
    pub fn camera_move_foward(&mut self, dist: f32) { check(unsafe { sys::draw_scene_camera_move(self.h, sys::DRAW_CAMERA_FOWARD, dist) }) }
    pub fn camera_move_backward(&mut self, dist: f32) { check(unsafe { sys::draw_scene_camera_move(self.h, sys::DRAW_CAMERA_BACKWARD, dist) }) }
    /// Scene::move_camera_direction (scene/mod.rs:803)
    pub fn move_camera_direction(&mut self, dx: i32, dy: i32) {
        check(unsafe { sys::draw_scene_move_camera_direction(self.h, dx, dy) })
    }
    /// Scene::render (scene/mod.rs:901): enqueues the frame; `Canvas::as_bytes_slice` waits for it.
    pub fn render(&mut self, canvas: &mut Canvas) {
        check(unsafe { sys::draw_scene_render(self.h, canvas.h) })
    }
}
impl Drop for Scene {
    fn drop(&mut self) {
        unsafe { sys::draw_scene_destroy(self.h) }
    }
}

/// VertexSimpleAttributes (canvas.rs:185-191); `color` is the payload of Color::Custom.
#[derive(Clone, Copy, Debug)]
pub struct VertexSimpleAttributes {
    pub screen_coord: [f32; 2],
    pub texture_coord: [f32; 2],
    pub color: [u8; 3],
    pub alpha: f32,
}
/// Rectangle::from_coords' arguments (canvas.rs:315-330).
#[derive(Clone, Copy, Debug)]
pub struct Rectangle {
    pub x0: usize,
    pub y0: usize,
    pub x1: usize,
    pub y1: usize,
}
impl Rectangle {
    pub fn from_coords(x0: usize, y0: usize, x1: usize, y1: usize) -> Self {
        Self { x0, y0, x1, y1 }
    }
}
/// The texture Canvas::draw_triangle samples (Texture::map_kd through get_rgba_slice, scene/mod.rs:137-152), on the device.
pub struct DeviceTexture {
    h: *mut sys::draw_texture,
}
impl DeviceTexture {
    pub fn new(texture: &Texture) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::draw_texture_create(&texture.map_kd.as_sys(), &mut h) });
        Self { h }
    }
}
impl Drop for DeviceTexture {
    fn drop(&mut self) {
        unsafe { sys::draw_texture_destroy(self.h) }
    }
}

pub struct Canvas {
    h: *mut sys::draw_canvas,
    pub width: usize,
    pub height: usize,
}
impl Canvas {
    /// Canvas::new (canvas.rs:366)
    pub fn new(width: usize, height: usize) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::draw_canvas_create(width, height, &mut h) });
        Self { h, width, height }
    }
    pub fn init_depth(&mut self, depth: f32) { check(unsafe { sys::draw_canvas_init_depth(self.h, depth) }) } // :403
    pub fn apply_offset(&mut self, x: i32, y: i32) { check(unsafe { sys::draw_canvas_apply_offset(self.h, x, y) }) } // :382
    pub fn resize(&mut self, width: usize, height: usize) {
        check(unsafe { sys::draw_canvas_resize(self.h, width, height) }); // :387
        self.width = width;
        self.height = height;
    }
    pub fn clear(&mut self) { check(unsafe { sys::draw_canvas_clear(self.h) }) } // :425
    pub fn enable_depth_update(&mut self) { check(unsafe { sys::draw_canvas_enable_depth_update(self.h) }) }
    pub fn disable_depth_update(&mut self) { check(unsafe { sys::draw_canvas_disable_depth_update(self.h) }) }
    pub fn pixel_bytes() -> usize { 4 } // :966
    pub fn size_bytes(&self) -> usize { self.width * self.height * 4 } // :970
    /// Canvas::as_bytes_slice (canvas.rs:974): B,G,R,pad per pixel, row 0 = top.  Waits for the frame and copies it
    /// to the library's pinned host mirror; the slice is valid until the next render / resize, like the borrow.
    pub fn as_bytes_slice(&self) -> &[u8] {
        let (mut p, mut n) = (std::ptr::null(), 0usize);
        check(unsafe { sys::draw_canvas_map_host(self.h, &mut p, &mut n) });
        unsafe { std::slice::from_raw_parts(p, n) }
    }
    pub fn as_ptr(&self) -> *const u8 { self.as_bytes_slice().as_ptr() } // :980
    /// get_pixel_depth (canvas.rs:413) for the whole buffer: row = canvas y (not flipped).
    pub fn depth(&self) -> Vec<f32> {
        let mut out = vec![0f32; self.width * self.height];
        check(unsafe { sys::draw_canvas_read_depth(self.h, out.as_mut_ptr(), out.len()) });
        out
    }
    /// Canvas::draw_triangle (canvas.rs:435-575).  The reference takes `Option<&Texture>` and asserts it is Some.
    pub fn draw_triangle(&mut self, a_vertex: VertexSimpleAttributes, b_vertex: VertexSimpleAttributes, c_vertex: VertexSimpleAttributes,
                         texture: Option<&DeviceTexture>, clipping_rect: Option<Rectangle>) {
        self.draw_triangles(&[a_vertex, b_vertex, c_vertex], texture.expect("draw_triangle needs a texture"), clipping_rect)
    }
    /// One draw command of Gui::render (src/app/gui.rs:382-485): `vertices.len() / 3` draw_triangle calls in order
    /// with one texture and one clipping rectangle, in two kernel launches.
    pub fn draw_triangles(&mut self, vertices: &[VertexSimpleAttributes], texture: &DeviceTexture, clipping_rect: Option<Rectangle>) {
        assert!(vertices.len() % 3 == 0);
        let v: Vec<sys::draw_vertex2d> = vertices
            .iter()
            .map(|a| sys::draw_vertex2d {
                x: a.screen_coord[0], y: a.screen_coord[1], u: a.texture_coord[0], v: a.texture_coord[1],
                r: a.color[0], g: a.color[1], b: a.color[2], pad: 0, alpha: a.alpha,
            })
            .collect();
        let rect = clipping_rect.map(|r| sys::draw_rect { x0: r.x0 as u64, y0: r.y0 as u64, x1: r.x1 as u64, y1: r.y1 as u64 });
        let rect_ptr = rect.as_ref().map_or(std::ptr::null(), |r| r as *const sys::draw_rect);
        check(unsafe { sys::draw_canvas_draw_triangles(self.h, v.as_ptr(), v.len() / 3, texture.h, rect_ptr) })
    }
    /// A whole GUI frame in one submission (src/app/gui.rs:389-485): `commands` are consecutive runs of `vertices`'
    /// triangles, each (number of triangles, clipping rectangle), drawn in order with one texture.
    pub fn draw_commands(&mut self, vertices: &[VertexSimpleAttributes], commands: &[(usize, Option<Rectangle>)], texture: &DeviceTexture) {
        let v: Vec<sys::draw_vertex2d> = vertices
            .iter()
            .map(|a| sys::draw_vertex2d {
                x: a.screen_coord[0], y: a.screen_coord[1], u: a.texture_coord[0], v: a.texture_coord[1],
                r: a.color[0], g: a.color[1], b: a.color[2], pad: 0, alpha: a.alpha,
            })
            .collect();
        let table: Vec<sys::draw_command2d> = commands
            .iter()
            .map(|(n, r)| sys::draw_command2d {
                n_triangles: *n,
                has_clip: r.is_some() as c_int,
                clip: r.map_or(sys::draw_rect::default(), |r| sys::draw_rect { x0: r.x0 as u64, y0: r.y0 as u64, x1: r.x1 as u64, y1: r.y1 as u64 }),
            })
            .collect();
        check(unsafe { sys::draw_canvas_draw_commands(self.h, v.as_ptr(), v.len() / 3, table.as_ptr(), table.len(), texture.h) })
    }
    /// Application::export_frame_as(Png) (app/mod.rs:316-360)
    pub fn export_png(&self, path: &str) {
        let c = CString::new(path).unwrap();
        check(unsafe { sys::draw_canvas_export_png(self.h, c.as_ptr()) })
    }
}
impl Canvas {
    /// Application::export_frame_as(Jpeg) (app/mod.rs:316-378)
    pub fn export_jpeg(&self, path: &str) {
        let c = CString::new(path).unwrap();
        check(unsafe { sys::draw_canvas_export_jpeg(self.h, c.as_ptr()) })
    }
}
impl Drop for Canvas {
    fn drop(&mut self) {
        unsafe { sys::draw_canvas_destroy(self.h) }
    }
}
