"""Scene / Camera / Canvas with the reference's names and verbs, over the C ABI.

Reads like code written against mororo18/draw's `renderer` module:

    scene = Scene(width, height)            # Scene::new            scene/mod.rs:760
    canvas = Canvas(width, height)          # Canvas::new           canvas.rs:366
    canvas.init_depth(100000.0)             #                       canvas.rs:403
    canvas.apply_offset(0, 0)               #                       canvas.rs:382
    scene.add_obj(obj)                      # Scene::add_obj        scene/mod.rs:788
    scene.camera = Camera.new(pos, dir)     # Camera::new           scene/mod.rs:297
    scene.render(canvas)                    # Scene::render         scene/mod.rs:901
    frame = canvas.as_bytes_slice()         # Canvas::as_bytes_slice canvas.rs:974

Everything executes in libdraw_b200.so on the GPU; this file only marshals arguments.
"""
import ctypes as C

import numpy as np

from . import _native as N
from .model import Object

DEPTH_MAX_DEFAULT = 100000.0  # Application::run, src/app/mod.rs:80


def _vec3(v):
    a = np.asarray(v, dtype=np.float32).reshape(3)
    return (C.c_float * 3)(float(a[0]), float(a[1]), float(a[2]))


class Camera:
    """Camera handle of a scene (scene/mod.rs:282-594).  `Camera.new(pos, dir)` builds a value that
    can be assigned to `scene.camera`, like `scene.camera = Camera::new(pos, dir, ratio)`; the
    aspect ratio is always the scene's width/height (scene/mod.rs:768)."""

    def __init__(self, scene):
        self._scene = scene

    @staticmethod
    def new(pos, direction):
        """Camera::new(pos, dir, ratio) as a value for `scene.camera = ...` (marshalled once, here)."""
        return ("camera", _vec3(pos), _vec3(direction))

    def _h(self):
        return self._scene._h

    def get_pos(self):
        p, d = (C.c_float * 3)(), (C.c_float * 3)()
        N.check(N.lib().draw_scene_get_camera(self._h(), p, d))
        return np.array(p, np.float32)

    def get_direction(self):
        p, d = (C.c_float * 3)(), (C.c_float * 3)()
        N.check(N.lib().draw_scene_get_camera(self._h(), p, d))
        return np.array(d, np.float32)

    def set_pos(self, pos):
        N.check(N.lib().draw_scene_set_camera_pos(self._h(), _vec3(pos)))

    def _move(self, which, dist):
        N.check(N.lib().draw_scene_camera_move(self._h(), which, float(dist)))

    def move_up(self, dist): self._move(0, dist)
    def move_down(self, dist): self._move(1, dist)
    def move_left(self, dist): self._move(2, dist)
    def move_right(self, dist): self._move(3, dist)
    def move_foward(self, dist): self._move(4, dist)
    def move_backward(self, dist): self._move(5, dist)


# VertexSimpleAttributes (canvas.rs:185-191) as draw_vertex2d lays it out
VERTEX2D = np.dtype([("x", "<f4"), ("y", "<f4"), ("u", "<f4"), ("v", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("pad", "u1"),
                     ("alpha", "<f4")])


class DeviceTexture:
    """The `texture: Option<&Texture>` argument of Canvas::draw_triangle (canvas.rs:435): map_kd as an RGBA8 array
    [height, width, 4], row 0 = top, copied to the device once (draw_texture_create)."""

    def __init__(self, map_kd):
        a = np.ascontiguousarray(map_kd, np.uint8)
        if a.ndim != 3 or a.shape[2] != 4:
            raise ValueError("get_rgba_slice needs a four-component map: uint8 [height, width, 4]")
        tm = N.TextureMap(a.ctypes.data, a.shape[1], a.shape[0], 4)
        h = C.c_void_p()
        N.check(N.lib().draw_texture_create(C.byref(tm), C.byref(h)))
        self._h = h
        self.width, self.height = int(a.shape[1]), int(a.shape[0])

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and N is not None:
            N.lib().draw_texture_destroy(h)


class Canvas:
    """canvas.rs:353-983."""

    def __init__(self, width, height):
        h = C.c_void_p()
        N.check(N.lib().draw_canvas_create(width, height, C.byref(h)))
        self._h = h
        self.width, self.height = int(width), int(height)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and N is not None:  # module globals are torn down at interpreter exit
            N.lib().draw_canvas_destroy(h)

    def init_depth(self, depth=DEPTH_MAX_DEFAULT):
        N.check(N.lib().draw_canvas_init_depth(self._h, float(depth)))

    def apply_offset(self, x, y):
        N.check(N.lib().draw_canvas_apply_offset(self._h, int(x), int(y)))

    def resize(self, width, height):
        N.check(N.lib().draw_canvas_resize(self._h, width, height))
        self.width, self.height = int(width), int(height)

    def clear(self):
        N.check(N.lib().draw_canvas_clear(self._h))

    def enable_depth_update(self):
        N.check(N.lib().draw_canvas_enable_depth_update(self._h))

    def disable_depth_update(self):
        N.check(N.lib().draw_canvas_disable_depth_update(self._h))

    def draw_triangles(self, vertices, texture, clipping_rect=None):
        """Canvas::draw_triangle (canvas.rs:435-575) for every three entries of `vertices` (a VERTEX2D array) in order,
        with one texture (a DeviceTexture, or an RGBA8 array uploaded for this call) and one clipping rectangle
        (x0, y0, x1, y1 as given to Rectangle::from_coords, or None) — one draw command of src/app/gui.rs:382-485."""
        v = np.ascontiguousarray(vertices, VERTEX2D)
        if v.ndim != 1 or v.size % 3:
            raise ValueError("vertices holds three VERTEX2D entries per triangle")
        tex = texture if isinstance(texture, DeviceTexture) else DeviceTexture(texture)
        rect = None if clipping_rect is None else C.byref(N.Rect(*[int(c) for c in clipping_rect]))
        N.check(N.lib().draw_canvas_draw_triangles(self._h, v.ctypes.data, v.size // 3, tex._h, rect))

    def draw_commands(self, commands, texture):
        """A whole GUI frame (src/app/gui.rs:389-485) in one submission: `commands` is a sequence of (clipping_rect or None,
        VERTEX2D array) pairs, drawn in order with one texture — the result of one draw_triangles call per pair."""
        commands = [(clip, np.ascontiguousarray(v, VERTEX2D)) for clip, v in commands]
        if any(v.ndim != 1 or v.size % 3 for _, v in commands):
            raise ValueError("vertices hold three VERTEX2D entries per triangle")
        verts = np.concatenate([v for _, v in commands]) if commands else np.zeros(0, VERTEX2D)
        table = (N.Command2D * max(1, len(commands)))()
        for k, (clip, v) in enumerate(commands):
            table[k].n_triangles = v.size // 3
            table[k].has_clip = 0 if clip is None else 1
            if clip is not None:
                table[k].clip = N.Rect(*[int(c) for c in clip])
        tex = texture if isinstance(texture, DeviceTexture) else DeviceTexture(texture)
        N.check(N.lib().draw_canvas_draw_commands(self._h, verts.ctypes.data, verts.size // 3, table, len(commands), tex._h))

    def draw_triangle(self, a_vertex, b_vertex, c_vertex, texture, clipping_rect=None):
        """One Canvas::draw_triangle call; each vertex is (x, y, u, v, (r, g, b), alpha)."""
        v = np.zeros(3, VERTEX2D)
        for k, (x, y, tu, tv, rgb, alpha) in enumerate((a_vertex, b_vertex, c_vertex)):
            v[k] = (x, y, tu, tv, rgb[0], rgb[1], rgb[2], 0, alpha)
        self.draw_triangles(v, texture, clipping_rect)

    @staticmethod
    def pixel_bytes():
        return 4  # size_of::<Pixel>(), canvas.rs:966

    def size_bytes(self):
        return self.width * self.height * 4

    def as_bytes_slice(self, copy=True):
        """uint8 [H, W, 4], B,G,R,pad per pixel, row 0 = top (canvas.rs:974).  Waits for the frame.
        copy=False returns a view of the library's pinned mirror, valid until the next render."""
        p, n = C.c_void_p(), C.c_size_t()
        N.check(N.lib().draw_canvas_map_host(self._h, C.byref(p), C.byref(n)))
        buf = (C.c_uint8 * n.value).from_address(p.value)
        a = np.frombuffer(buf, dtype=np.uint8).reshape(self.height, self.width, 4)
        return a.copy() if copy else a

    def enable_host_mirror(self, enabled=True):
        """Every render also copies the frame to the pinned host mirror (the reference's frame lives in
        host memory); as_bytes_slice then only waits."""
        N.check(N.lib().draw_canvas_enable_host_mirror(self._h, 1 if enabled else 0))

    def depth(self):
        """float32 [H, W]; row = canvas y, not flipped (get_pixel_depth, canvas.rs:413)."""
        out = np.empty((self.height, self.width), np.float32)
        N.check(N.lib().draw_canvas_read_depth(self._h, out.ctypes.data, out.size))
        return out

    def get_pixel_depth(self, x, y):
        return float(self.depth()[y, x])

    def sync(self):
        N.check(N.lib().draw_canvas_sync(self._h))

    def last_frame_stats(self):
        s = N.FrameStats()
        N.check(N.lib().draw_canvas_last_frame_stats(self._h, C.byref(s)))
        return s.as_dict()

    # ---- device-side plumbing (multi-GPU drivers, torch interop)
    def device_ptrs(self):
        c, d = C.c_void_p(), C.c_void_p()
        N.check(N.lib().draw_canvas_device_ptrs(self._h, C.byref(c), C.byref(d)))
        return c.value, d.value

    def bind_external(self, color_ptr, depth_ptr):
        N.check(N.lib().draw_canvas_bind_external(self._h, color_ptr, depth_ptr))

    def set_stream(self, cuda_stream):
        N.check(N.lib().draw_canvas_set_stream(self._h, cuda_stream))

    def export_png(self, path):
        """Application::export_frame_as(Png) (app/mod.rs:316-360): the current frame as an RGBA PNG file."""
        N.check(N.lib().draw_canvas_export_png(self._h, str(path).encode()))

    def export_jpeg(self, path):
        """Application::export_frame_as(Jpeg) (app/mod.rs:316-378): the current frame as a quality-100 baseline JPEG file."""
        N.check(N.lib().draw_canvas_export_jpeg(self._h, str(path).encode()))

    def stream_wait(self, cuda_stream):
        """Make `cuda_stream` wait (on the device) for everything enqueued so far for this canvas."""
        N.check(N.lib().draw_canvas_stream_wait(self._h, cuda_stream))

    def set_stripe(self, y0, y1):
        N.check(N.lib().draw_canvas_set_stripe(self._h, y0, y1))

    def set_empty_tile_color(self, enabled=True):
        """False: renders leave the colour of tiles without geometry untouched (the buffer was cleared by its owner)."""
        N.check(N.lib().draw_canvas_set_empty_tile_color(self._h, 1 if enabled else 0))

    def set_tile_rows(self, phase, step):
        """Sort-first, interleaved: render only the tile rows ty with ty % step == phase."""
        N.check(N.lib().draw_canvas_set_tile_rows(self._h, int(phase), int(step)))


class ObjectInfo:
    """scene/object.rs:12-16."""

    def __init__(self, id_, name, mesh_info_list):
        self.id, self.name, self.mesh_info_list = id_, name, mesh_info_list


class Scene:
    """scene/mod.rs:749-1252."""

    def __init__(self, width, height):
        h = C.c_void_p()
        N.check(N.lib().draw_scene_create(width, height, C.byref(h)))
        self._h = h
        self.width, self.height = int(width), int(height)
        self._camera = Camera(self)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and N is not None:
            N.lib().draw_scene_destroy(h)

    @property
    def camera(self):
        return self._camera

    @camera.setter
    def camera(self, value):
        if not (isinstance(value, tuple) and value and value[0] == "camera"):
            raise TypeError("assign Camera.new(pos, direction)")
        N.check(N.lib().draw_scene_set_camera(self._h, value[1], value[2]))

    def set_light(self, pos):
        N.check(N.lib().draw_scene_set_light(self._h, _vec3(pos)))

    def add_obj(self, obj: Object):
        pos = np.ascontiguousarray(obj.vertices, np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(obj.normals_vertices, np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(obj.texture_vertices, np.float32).reshape(-1, 3)
        keep = [pos, nrm, uv]
        mats = (N.Material * max(1, len(obj.textures)))()
        for i, t in enumerate(obj.textures):
            m = mats[i]
            m.name = t.name.encode()
            for j in range(3):
                m.ka[j], m.kd[j], m.ks[j] = float(np.float32(t.ka[j])), float(np.float32(t.kd[j])), float(np.float32(t.ks[j]))
            m.alpha = float(np.float32(t.alpha))
            for field, img in (("map_ka", t.map_ka), ("map_kd", t.map_kd)):
                tm = getattr(m, field)
                if img is None:
                    tm.pixels = None
                    continue
                img = np.ascontiguousarray(img, np.uint8)
                if img.ndim != 3:
                    raise ValueError("texture maps are uint8 [height, width, components]")
                same = [k for k in keep if k is img or (k.dtype == np.uint8 and k.shape == img.shape
                                                        and k.ctypes.data == img.ctypes.data)]
                if not same:
                    keep.append(img)
                tm.pixels = img.ctypes.data
                tm.height, tm.width, tm.components = img.shape
        meshes = (N.Mesh * max(1, len(obj.meshes)))()
        for i, me in enumerate(obj.meshes):
            tris = np.ascontiguousarray(me.triangles, np.uint32).reshape(-1, 9)
            keep.append(tris)
            meshes[i].name = me.name.encode()
            meshes[i].triangles = tris.ctypes.data
            meshes[i].n_triangles = tris.shape[0]
            meshes[i].material_idx = int(me.texture_idx)
        desc = N.ObjectDesc(obj.name.encode(), pos.ctypes.data, pos.shape[0], nrm.ctypes.data, nrm.shape[0],
                            uv.ctypes.data, uv.shape[0], C.addressof(meshes), len(obj.meshes),
                            C.addressof(mats), len(obj.textures))
        out = C.c_uint32()
        N.check(N.lib().draw_scene_add_object(self._h, C.byref(desc), C.byref(out)))
        del keep
        infos = [(m.name, int(np.asarray(m.triangles).reshape(-1, 9).shape[0]),
                  obj.textures[m.texture_idx].name) for m in obj.meshes]
        return ObjectInfo(int(out.value), obj.name, infos)

    def move_camera_direction(self, dx, dy):
        N.check(N.lib().draw_scene_move_camera_direction(self._h, int(dx), int(dy)))

    def render(self, canvas: Canvas):
        N.check(N.lib().draw_scene_render(self._h, canvas._h))

    def prepare(self, canvas: Canvas):
        """Set up every frame-in-flight slot for this canvas' geometry now and wait (draw_scene_prepare)."""
        N.check(N.lib().draw_scene_prepare(self._h, canvas._h))

    # ---- parity / measurement taps
    def uniforms(self):
        m, p = (C.c_float * 16)(), (C.c_float * 24)()
        N.check(N.lib().draw_scene_get_uniforms(self._h, m, p))
        return np.array(m, np.float32).reshape(4, 4), np.array(p, np.float32).reshape(6, 4)

    def vertex_visual(self, canvas, first, count):
        out = np.empty((count, 7), np.float32)
        N.check(N.lib().draw_scene_read_vertex_visual(self._h, canvas._h, first, count, out.ctypes.data))
        return out

    KERNELS = ("k_sort_transparent", "k_front", "k_raster", "k_tile")
    OPTIONAL_KERNELS = ("k_sort_transparent",)

    def set_kernel_timing(self, enabled):
        N.check(N.lib().draw_scene_set_kernel_timing(self._h, 1 if enabled else 0))

    def last_kernel_times(self, canvas):
        """Device time (ms) of each kernel of the last frame, keyed by kernel name."""
        ms = (C.c_float * 4)()
        N.check(N.lib().draw_scene_last_kernel_times(self._h, canvas._h, ms))
        return dict(zip(self.KERNELS, (float(x) for x in ms)))

    def debug_list_counts(self, canvas):
        """Per tile of the last frame: large references, medium / small weight, transparent references (uint32 arrays)."""
        nc = C.c_size_t()
        N.check(N.lib().draw_scene_debug_list_counts(self._h, canvas._h, None, 0, C.byref(nc)))
        n = nc.value * 3
        out = np.empty(n, np.uint32)
        N.check(N.lib().draw_scene_debug_list_counts(self._h, canvas._h, out.ctypes.data, n, C.byref(nc)))
        return out[:nc.value], out[nc.value:2 * nc.value], out[2 * nc.value:]

    def debug_trace(self, enable=True, cap=1 << 18):
        """Returns the CTA records gathered so far as uint32 [n, 4] (kernel | SM << 8 | work set << 24,
        CTA, start ns, end ns), then switches recording on or off and empties the buffer."""
        out = np.zeros((cap, 4), np.uint32)
        n = C.c_size_t(0)
        N.check(N.lib().draw_scene_debug_trace(self._h, 1 if enable else 0, out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    def debug_tile_cycles(self, canvas=None, enable=True):
        """Toggle per-tile cycle recording; with a canvas, return the last frame's cycles per coarse tile."""
        if canvas is None:
            N.check(N.lib().draw_scene_debug_tile_cycles(self._h, None, 1 if enable else 0, None, 0))
            return None
        nc = C.c_size_t()
        N.check(N.lib().draw_scene_debug_list_counts(self._h, canvas._h, None, 0, C.byref(nc)))
        out = np.zeros(nc.value * 4, np.uint32)
        N.check(N.lib().draw_scene_debug_tile_cycles(self._h, canvas._h, 1 if enable else 0, out.ctypes.data, nc.value * 4))
        return out.reshape(4, nc.value)  # rows: whole item, end of phase A, of phase C, of phase D (cycles since the item started)

    def counts(self):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        N.check(N.lib().draw_scene_counts(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"objects": a.value, "triangles": b.value, "vertices": c.value}

    def launch_count(self):
        n = C.c_uint64()
        N.check(N.lib().draw_scene_launch_count(self._h, C.byref(n)))
        return int(n.value)


def device_count():
    n = C.c_int()
    N.check(N.lib().draw_device_count(C.byref(n)))
    return n.value


def set_device(index):
    N.check(N.lib().draw_set_device(int(index)))


def tile_size():
    return N.lib().draw_tile_size()


def load_image(path):
    """TextureMap::load_from_file (scene/mod.rs:174-202) through the library's own decoders (PNG; JPEG with stb_image's arithmetic):
    uint8 [height, width, components], components 3 or 4, row 0 = top."""
    px, w, h, c = C.c_void_p(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    N.check(N.lib().draw_image_load(str(path).encode(), C.byref(px), C.byref(w), C.byref(h), C.byref(c)))
    try:
        n = w.value * h.value * c.value
        return np.ctypeslib.as_array(C.cast(px, C.POINTER(C.c_uint8)), shape=(n,)).copy().reshape(h.value, w.value, c.value)
    finally:
        N.lib().draw_image_free(px)


def write_png(path, pixels):
    """stbi_write_png (app/mod.rs:362-378): uint8 [height, width, 3 or 4] -> PNG file."""
    a = np.ascontiguousarray(pixels, dtype=np.uint8)
    N.check(N.lib().draw_image_write_png(str(path).encode(), a.ctypes.data, a.shape[1], a.shape[0], a.shape[2]))


def write_jpg(path, pixels, quality=100):
    """stbi_write_jpg (app/mod.rs:362-378): uint8 [height, width, 3 or 4] -> baseline JPEG file; quality as
    stb_image_write reads it (0 = 90, clamped to 1..100; only qualities above 90 are written)."""
    a = np.ascontiguousarray(pixels, dtype=np.uint8)
    N.check(N.lib().draw_image_write_jpg(str(path).encode(), a.ctypes.data, a.shape[1], a.shape[0], a.shape[2], int(quality)))


def load_obj(path, decode_images=True):
    """Object::load_from_file (scene/object.rs:106) through the library's C++ loader
    (draw_object_load_obj).  Texture files named by the MTL are decoded by the library (PNG, JPEG) when
    decode_images is true — PIL only stands in for the formats the library does not decode — else maps stay at
    the 1x1 default."""
    from .model import IndexedMesh, Texture
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]

    def _decode(cpath, _user, out_pixels, out_w, out_h, out_comp):
        # PNG / JPEG: the library's own decoders; anything else: PIL stands in for stb_image
        if N.lib().draw_image_loader_builtin(cpath, None, out_pixels, out_w, out_h, out_comp) == 0:
            return 0
        try:
            from PIL import Image
            im = Image.open(cpath.decode())
            if im.mode != "RGBA":
                im = im.convert("RGB")
            a = np.ascontiguousarray(np.asarray(im, dtype=np.uint8))
            buf = libc.malloc(a.size)
            C.memmove(buf, a.ctypes.data, a.size)
            out_pixels[0] = buf
            out_w[0], out_h[0], out_comp[0] = a.shape[1], a.shape[0], a.shape[2]
            return 0
        except Exception:
            return 1

    cb = N.IMAGE_LOADER(_decode) if decode_images else None
    handle = C.c_void_p()
    N.check(N.lib().draw_object_load_obj(str(path).encode(), cb, None, C.byref(handle)))
    try:
        d = N.ObjectDesc()
        N.check(N.lib().draw_object_desc_of(handle, C.byref(d)))

        def arr(ptr, n, ctype, dtype, cols):
            if n == 0:
                return np.zeros((0, cols), dtype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n * cols,)).astype(dtype, copy=True).reshape(n, cols)

        meshes_c = C.cast(d.meshes, C.POINTER(N.Mesh))
        mats_c = C.cast(d.materials, C.POINTER(N.Material))
        meshes = [IndexedMesh((meshes_c[i].name or b"").decode(), arr(meshes_c[i].triangles, meshes_c[i].n_triangles, C.c_uint32, np.uint32, 9),
                              int(meshes_c[i].material_idx)) for i in range(d.n_meshes)]
        images = {}

        def img(tm):
            if not tm.pixels:
                return None
            if tm.pixels not in images:
                n = tm.width * tm.height * tm.components
                images[tm.pixels] = np.ctypeslib.as_array(C.cast(tm.pixels, C.POINTER(C.c_uint8)), shape=(n,)).copy().reshape(tm.height, tm.width, tm.components)
            return images[tm.pixels]

        textures = [Texture((mats_c[i].name or b"").decode(), np.array(mats_c[i].ka, np.float32), np.array(mats_c[i].kd, np.float32),
                            np.array(mats_c[i].ks, np.float32), float(mats_c[i].alpha), img(mats_c[i].map_ka), img(mats_c[i].map_kd))
                    for i in range(d.n_materials)]
        return Object((d.name or b"").decode(), arr(d.positions, d.n_positions, C.c_float, np.float32, 3),
                      arr(d.normals, d.n_normals, C.c_float, np.float32, 3), arr(d.uvs, d.n_uvs, C.c_float, np.float32, 3),
                      meshes, textures)
    finally:
        N.lib().draw_object_free(handle)
