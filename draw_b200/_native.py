"""ctypes binding of libdraw_b200.so (the C ABI in include/draw_b200.h).

The library is the product; there is no Python or CPU fallback.  Importing this module fails
loudly if the shared library has not been built (python -m draw_b200.build), and every compute
call fails with DrawError if no CUDA device is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DRAW_B200_LIB selects an experimental build variant (python -m draw_b200.build -D... --out=...)
LIB_PATH = os.environ.get("DRAW_B200_LIB") or os.path.join(_HERE, "libdraw_b200.so")


class DrawError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"draw_b200 error {code}: {message}")
        self.code = code


class TextureMap(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("components", C.c_uint32)]


class Material(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ka", C.c_float * 3), ("kd", C.c_float * 3), ("ks", C.c_float * 3),
                ("alpha", C.c_float), ("map_ka", TextureMap), ("map_kd", TextureMap)]


class Mesh(C.Structure):
    _fields_ = [("name", C.c_char_p), ("triangles", C.c_void_p), ("n_triangles", C.c_size_t),
                ("material_idx", C.c_uint32)]


class ObjectDesc(C.Structure):
    _fields_ = [("name", C.c_char_p),
                ("positions", C.c_void_p), ("n_positions", C.c_size_t),
                ("normals", C.c_void_p), ("n_normals", C.c_size_t),
                ("uvs", C.c_void_p), ("n_uvs", C.c_size_t),
                ("meshes", C.c_void_p), ("n_meshes", C.c_size_t),
                ("materials", C.c_void_p), ("n_materials", C.c_size_t)]


class Rect(C.Structure):  # draw_rect
    _fields_ = [("x0", C.c_uint64), ("y0", C.c_uint64), ("x1", C.c_uint64), ("y1", C.c_uint64)]


class Command2D(C.Structure):  # draw_command2d
    _fields_ = [("n_triangles", C.c_size_t), ("has_clip", C.c_int), ("clip", Rect)]


class FrameStats(C.Structure):
    _NAMES = ("input_triangles", "setup_records", "tile_refs", "large_refs", "medium_refs", "small_refs", "transparent_refs",
              "overflow", "empty_tiles", "work_items", "mirror_kbytes")
    _fields_ = [(n, C.c_uint32) for n in _NAMES] + [("front_phase_ns", C.c_uint32 * 7), ("front_block_ns", C.c_uint32 * 5)]

    def as_dict(self):
        d = {n: int(getattr(self, n)) for n in self._NAMES}
        d["front_phase_ns"] = [int(x) for x in self.front_phase_ns]
        d["front_block_ns"] = [int(x) for x in self.front_block_ns]
        return d


IMAGE_LOADER = C.CFUNCTYPE(C.c_int, C.c_char_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32),
                           C.POINTER(C.c_uint32), C.POINTER(C.c_uint32))

# name -> (restype, argtypes); exactly the symbols declared in include/draw_b200.h
SIGNATURES = {
    "draw_version": (C.c_int, []),
    "draw_last_error": (C.c_char_p, []),
    "draw_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "draw_set_device": (C.c_int, [C.c_int]),
    "draw_scene_create": (C.c_int, [C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p)]),
    "draw_scene_destroy": (None, [C.c_void_p]),
    "draw_scene_add_object": (C.c_int, [C.c_void_p, C.POINTER(ObjectDesc), C.POINTER(C.c_uint32)]),
    "draw_scene_set_camera": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "draw_scene_get_camera": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "draw_scene_set_camera_pos": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "draw_scene_camera_move": (C.c_int, [C.c_void_p, C.c_int, C.c_float]),
    "draw_scene_move_camera_direction": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "draw_scene_set_light": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "draw_scene_render": (C.c_int, [C.c_void_p, C.c_void_p]),
    "draw_scene_prepare": (C.c_int, [C.c_void_p, C.c_void_p]),
    "draw_scene_get_uniforms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "draw_scene_read_vertex_visual": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "draw_scene_counts": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "draw_scene_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "draw_scene_set_kernel_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "draw_scene_last_kernel_times": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "draw_scene_debug_list_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "draw_scene_debug_tile_cycles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "draw_scene_debug_trace": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "draw_canvas_create": (C.c_int, [C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p)]),
    "draw_canvas_destroy": (None, [C.c_void_p]),
    "draw_canvas_init_depth": (C.c_int, [C.c_void_p, C.c_float]),
    "draw_canvas_apply_offset": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "draw_canvas_resize": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "draw_canvas_clear": (C.c_int, [C.c_void_p]),
    "draw_canvas_enable_depth_update": (C.c_int, [C.c_void_p]),
    "draw_canvas_disable_depth_update": (C.c_int, [C.c_void_p]),
    "draw_canvas_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "draw_canvas_map_host": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "draw_canvas_enable_host_mirror": (C.c_int, [C.c_void_p, C.c_int]),
    "draw_canvas_read_depth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "draw_canvas_sync": (C.c_int, [C.c_void_p]),
    "draw_canvas_last_frame_stats": (C.c_int, [C.c_void_p, C.POINTER(FrameStats)]),
    "draw_canvas_device_ptrs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "draw_canvas_bind_external": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "draw_canvas_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "draw_canvas_stream_wait": (C.c_int, [C.c_void_p, C.c_void_p]),
    "draw_canvas_set_stripe": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "draw_canvas_set_tile_rows": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "draw_canvas_set_empty_tile_color": (C.c_int, [C.c_void_p, C.c_int]),
    "draw_texture_create": (C.c_int, [C.POINTER(TextureMap), C.POINTER(C.c_void_p)]),
    "draw_texture_destroy": (None, [C.c_void_p]),
    "draw_canvas_draw_triangles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(Rect)]),
    "draw_canvas_draw_commands": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Command2D), C.c_size_t, C.c_void_p]),
    "draw_device_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "draw_device_free": (C.c_int, [C.c_void_p]),
    "draw_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "draw_flag_signal": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "draw_flags_wait": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "draw_tile_size": (C.c_int, []),
    "draw_canvas_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "draw_ipc_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "draw_ipc_close": (C.c_int, [C.c_void_p]),
    "draw_object_load_obj": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "draw_image_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "draw_image_free": (None, [C.c_void_p]),
    "draw_image_write_png": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "draw_canvas_export_png": (C.c_int, [C.c_void_p, C.c_char_p]),
    "draw_canvas_export_jpeg": (C.c_int, [C.c_void_p, C.c_char_p]),
    "draw_image_write_jpg": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
    "draw_image_loader_builtin": (C.c_int, [C.c_char_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "draw_object_free": (None, [C.c_void_p]),
    "draw_object_desc_of": (C.c_int, [C.c_void_p, C.POINTER(ObjectDesc)]),
}

_lib = None


def lib():
    """Load libdraw_b200.so (once).  Raises if it is missing: build it, there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: run `python -m draw_b200.build` (needs nvcc). "
                              "draw_b200 has no Python or CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.draw_version() != 200:
            raise ImportError(f"libdraw_b200.so version {L.draw_version()} does not match this binding (200)")
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise DrawError(rc, lib().draw_last_error().decode("utf-8", "replace"))
