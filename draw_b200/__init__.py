"""draw_b200 — B200-native implementation of mororo18/draw's software rasterization path.

The renderer itself is libdraw_b200.so (hand-written sm_100a CUDA kernels behind the C ABI of
include/draw_b200.h).  This package is the thin host-side mirror of the reference's
Scene / Camera / Canvas / Object interface plus data helpers.
"""
from .model import IndexedMesh, Object, Texture  # noqa: F401

__all__ = ["Scene", "Canvas", "Camera", "Object", "IndexedMesh", "Texture", "DrawError"]


def __getattr__(name):
    # api needs the native library; keep `import draw_b200.model` usable for tools that only
    # handle data, but anything that renders goes through the library or fails loudly.
    if name in ("Scene", "Canvas", "Camera", "ObjectInfo", "device_count", "set_device", "tile_size", "load_obj", "load_image", "write_png", "write_jpg", "DeviceTexture", "VERTEX2D"):
        from . import api
        return getattr(api, name)
    if name == "DrawError":
        from ._native import DrawError
        return DrawError
    raise AttributeError(name)
