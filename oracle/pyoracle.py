"""ctypes wrapper of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  The product package (draw_b200/) never imports this.

The classes mirror the reference's Scene / Canvas verbs (src/renderer/scene/mod.rs:749-1252,
src/renderer/canvas.rs:353-983) so a test reads like code written against the reference.
Objects are duck-typed: anything with .vertices, .normals_vertices, .texture_vertices
(float32 [N,3]), .meshes (each .triangles uint32 [T,9] as v0 v1 v2 t0 t1 t2 n0 n1 n2 and
.texture_idx) and .textures (each .ka .kd .ks .alpha .map_ka .map_kd, maps = None or
uint8 [h,w,comp]) is accepted.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    """Compile oracle.cpp with the committed Makefile (g++, -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


class _Material(C.Structure):
    _fields_ = [("ka", C.c_float * 3), ("kd", C.c_float * 3), ("ks", C.c_float * 3),
                ("alpha", C.c_float),
                ("map_ka", C.c_void_p), ("map_ka_w", C.c_uint32), ("map_ka_h", C.c_uint32),
                ("map_ka_comp", C.c_uint32),
                ("map_kd", C.c_void_p), ("map_kd_w", C.c_uint32), ("map_kd_h", C.c_uint32),
                ("map_kd_comp", C.c_uint32)]


class _Mesh(C.Structure):
    _fields_ = [("tris", C.c_void_p), ("n_tris", C.c_size_t), ("texture_idx", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("input_tris", "culled_tris", "emitted_tris",
                                          "bbox_pixels", "covered_frags", "written_frags")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_scene_new.restype = C.c_void_p
        L.orc_scene_new.argtypes = [C.c_size_t, C.c_size_t]
        L.orc_scene_free.argtypes = [C.c_void_p]
        L.orc_scene_add_object.restype = C.c_int
        L.orc_scene_add_object.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                           C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                           C.c_void_p, C.c_size_t]
        L.orc_scene_set_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_scene_set_light.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_scene_camera_move.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.orc_scene_move_camera_direction.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_scene_get_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_scene_uniforms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_scene_vertex_visual.restype = C.c_int
        L.orc_scene_vertex_visual.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_canvas_new.restype = C.c_void_p
        L.orc_canvas_new.argtypes = [C.c_size_t, C.c_size_t]
        L.orc_canvas_free.argtypes = [C.c_void_p]
        L.orc_canvas_init_depth.argtypes = [C.c_void_p, C.c_float]
        L.orc_canvas_apply_offset.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_canvas_resize.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.orc_canvas_clear.argtypes = [C.c_void_p]
        L.orc_canvas_bytes.restype = C.c_void_p
        L.orc_canvas_bytes.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_canvas_depth.restype = C.c_void_p
        L.orc_canvas_depth.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_canvas_winner.restype = C.c_void_p
        L.orc_canvas_winner.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_canvas_draw_triangle.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_canvas_set_depth_update.argtypes = [C.c_void_p, C.c_int]
        L.orc_scene_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_scene_stats.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _f32(a, cols=3):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return a.reshape(-1, cols) if a.size else a.reshape(0, cols)


def _vec3(v):
    return (C.c_float * 3)(*[float(np.float32(x)) for x in v])


# VertexSimpleAttributes (canvas.rs:185-191) as a 24-byte record: the layout of draw_vertex2d / orc_vertex2d
VERTEX2D = np.dtype([("x", "<f4"), ("y", "<f4"), ("u", "<f4"), ("v", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("pad", "u1"),
                     ("alpha", "<f4")])


class Canvas:
    """canvas.rs:353-433 — new / init_depth / apply_offset / resize / clear; draw_triangle (:435-575)."""

    def __init__(self, width, height):
        self._h = lib().orc_canvas_new(width, height)
        self.width, self.height = width, height

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_canvas_free(self._h)
            self._h = None

    def init_depth(self, depth):
        lib().orc_canvas_init_depth(self._h, depth)

    def apply_offset(self, x, y):
        lib().orc_canvas_apply_offset(self._h, x, y)

    def resize(self, width, height):
        lib().orc_canvas_resize(self._h, width, height)
        self.width, self.height = width, height

    def clear(self):
        lib().orc_canvas_clear(self._h)

    def enable_depth_update(self):
        lib().orc_canvas_set_depth_update(self._h, 1)

    def disable_depth_update(self):
        lib().orc_canvas_set_depth_update(self._h, 0)

    def draw_triangles(self, vertices, texture, clipping_rect=None):
        """Canvas::draw_triangle (canvas.rs:435-575) for each consecutive triple of `vertices` (VERTEX2D records:
        x, y, u, v, r, g, b, pad, alpha), in order.  texture: uint8 [h, w, 4]; clipping_rect: (x0, y0, x1, y1) or None."""
        v = np.ascontiguousarray(vertices, dtype=VERTEX2D)
        tex = np.ascontiguousarray(texture, np.uint8)
        assert tex.ndim == 3 and tex.shape[2] == 4 and v.size % 3 == 0
        clip = (C.c_uint64 * 4)(*[int(c) for c in clipping_rect]) if clipping_rect is not None else None
        for t in range(v.size // 3):
            lib().orc_canvas_draw_triangle(self._h, v[3 * t:3 * t + 3].ctypes.data, tex.ctypes.data, tex.shape[1], tex.shape[0], clip)

    def _view(self, fn, ctype, dtype):
        n = C.c_size_t(0)
        p = fn(self._h, C.byref(n))
        if not p or n.value == 0:
            return np.zeros(0, dtype)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(n.value,)).astype(dtype, copy=True)

    def as_bytes(self):
        """as_bytes_slice (canvas.rs:974): uint8 [H, W, 4] in B,G,R,pad order, row 0 = top."""
        return self._view(lib().orc_canvas_bytes, C.c_uint8, np.uint8).reshape(self.height, self.width, 4)

    def depth(self):
        """depth_frame, float32 [H, W]; row index = canvas y (NOT flipped, canvas.rs:413-423)."""
        return self._view(lib().orc_canvas_depth, C.c_float, np.float32).reshape(self.height, self.width)

    def winner(self):
        return self._view(lib().orc_canvas_winner, C.c_uint32, np.uint32).reshape(self.height, self.width)


class Scene:
    """scene/mod.rs:749-1252 — new / add_obj / render / camera verbs."""

    def __init__(self, width, height):
        self._h = lib().orc_scene_new(width, height)
        self.width, self.height = width, height
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_scene_free(self._h)
            self._h = None

    def add_obj(self, obj):
        pos, nrm, uv = _f32(obj.vertices), _f32(obj.normals_vertices), _f32(obj.texture_vertices)
        mats = (_Material * max(1, len(obj.textures)))()
        keep = [pos, nrm, uv]
        for i, t in enumerate(obj.textures):
            m = mats[i]
            for j in range(3):
                m.ka[j], m.kd[j], m.ks[j] = np.float32(t.ka[j]), np.float32(t.kd[j]), np.float32(t.ks[j])
            m.alpha = np.float32(t.alpha)
            for name in ("map_ka", "map_kd"):
                img = getattr(t, name)
                if img is None:
                    setattr(m, name, None)
                    continue
                img = np.ascontiguousarray(img, dtype=np.uint8)
                assert img.ndim == 3
                keep.append(img)
                setattr(m, name, img.ctypes.data)
                setattr(m, name + "_w", img.shape[1])
                setattr(m, name + "_h", img.shape[0])
                setattr(m, name + "_comp", img.shape[2])
        meshes = (_Mesh * max(1, len(obj.meshes)))()
        for i, me in enumerate(obj.meshes):
            tris = np.ascontiguousarray(me.triangles, dtype=np.uint32).reshape(-1, 9)
            keep.append(tris)
            meshes[i].tris = tris.ctypes.data
            meshes[i].n_tris = tris.shape[0]
            meshes[i].texture_idx = me.texture_idx
        rc = lib().orc_scene_add_object(self._h, pos.ctypes.data, pos.shape[0], nrm.ctypes.data, nrm.shape[0],
                                        uv.ctypes.data, uv.shape[0], C.addressof(meshes), len(obj.meshes),
                                        C.addressof(mats), len(obj.textures))
        if rc < 0:
            raise ValueError(f"oracle add_object failed ({rc})")
        return rc

    def set_camera(self, pos, direction):
        """scene.camera = Camera::new(pos, dir, W/H) (scene/mod.rs:297)."""
        lib().orc_scene_set_camera(self._h, _vec3(pos), _vec3(direction))

    def set_light(self, pos):
        lib().orc_scene_set_light(self._h, _vec3(pos))

    def camera_move(self, which, dist):
        names = {"up": 0, "down": 1, "left": 2, "right": 3, "foward": 4, "backward": 5}
        lib().orc_scene_camera_move(self._h, names[which] if isinstance(which, str) else which, dist)

    def move_camera_direction(self, dx, dy):
        lib().orc_scene_move_camera_direction(self._h, dx, dy)

    def get_camera(self):
        p, d = (C.c_float * 3)(), (C.c_float * 3)()
        lib().orc_scene_get_camera(self._h, p, d)
        return np.array(p, np.float32), np.array(d, np.float32)

    def uniforms(self):
        m, pl = (C.c_float * 16)(), (C.c_float * 24)()
        lib().orc_scene_uniforms(self._h, m, pl)
        return np.array(m, np.float32).reshape(4, 4), np.array(pl, np.float32).reshape(6, 4)

    def vertex_visual(self, obj_idx, n_vertices):
        out = np.zeros((n_vertices, 10), np.float32)
        rc = lib().orc_scene_vertex_visual(self._h, obj_idx, out.ctypes.data, n_vertices)
        if rc:
            raise ValueError(f"vertex_visual failed ({rc})")
        return out

    def render(self, canvas, stats=False):
        lib().orc_scene_render(self._h, canvas._h, 1 if stats else 0)

    def stats(self):
        s = Stats()
        lib().orc_scene_stats(self._h, C.byref(s))
        return s.as_dict()
