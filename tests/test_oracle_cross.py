"""Pin the C++ oracle against the independent numpy-float32 restatement, bit for bit.

The reference publishes no golden vectors, so the pin is manufactured:
oracle/oracle.cpp and oracle/np_oracle.py were written separately from the Rust source and
must agree exactly on colour bytes and depth bits, over scenes that exercise every branch of
the path: plain raster, texture fetch, near/far clipping (0/1/2 outputs), the transparent
pass (painter sort + blend + depth-write off), a canvas offset, and a scene larger than the
canvas.
"""
import numpy as np
import pytest

from conftest import load_scene
from draw_b200 import synthetic
from draw_b200.model import Texture
from oracle import np_oracle, pyoracle

F = np.float32


def _render_both(objs, W, H, cam=None, offset=(0, 0), scene_wh=None, frames=1):
    sw, sh = scene_wh or (W, H)
    s1, c1 = pyoracle.Scene(sw, sh), pyoracle.Canvas(W, H)
    s2, c2 = np_oracle.Scene(sw, sh), np_oracle.Canvas(W, H)
    c1.init_depth(100000.0)
    c2.init_depth(100000.0)
    c1.apply_offset(*offset)
    c2.apply_offset(*offset)
    for o in objs:
        s1.add_obj(o)
        s2.add_obj(o)
    if cam is not None:
        s1.set_camera(cam[:3], cam[3:])
        s2.set_camera(cam[:3], cam[3:])
    for _ in range(frames):
        s1.render(c1)
        s2.render(c2)
    return (c1.as_bytes(), c1.depth()), (c2.frame, c2.depth_frame), s1, s2


def _assert_same(a, b):
    assert np.array_equal(a[0], b[0]), "colour bytes differ"
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), "depth bits differ"
    assert (a[1] < 100000).any(), "nothing was drawn"


def test_uniforms_agree():
    for (w, h) in ((800, 600), (1920, 1080), (64, 48)):
        s1, s2 = pyoracle.Scene(w, h), np_oracle.Scene(w, h)
        cam = np.array([30.5, 10.0, 120.25, 0.3, -0.05, -1.0], F)
        s1.set_camera(cam[:3], cam[3:])
        s2.set_camera(cam[:3], cam[3:])
        m1, p1 = s1.uniforms()
        assert np.array_equal(m1.view(np.uint32), s2.transformation_matrix().view(np.uint32))
        depth, lateral = s2.camera.gen_view_planes()
        p2 = np.array([[*n, k] for n, k in depth + lateral], F)
        assert np.array_equal(p1.view(np.uint32), p2.view(np.uint32))


def test_donut_default_camera():
    a, b, _, _ = _render_both(load_scene("c2_donut"), 96, 64)
    _assert_same(a, b)


def test_textured_and_transparent_two_frames():
    """lemur (RGBA texture) + airplane stand-in (RGB texture, glass alpha 0.7); two frames so the
    persistent painter sort of the transparent mesh (scene/mod.rs:1100-1115) is covered."""
    a, b, _, _ = _render_both(load_scene("c1_lemur_airplane"), 120, 90, frames=2)
    _assert_same(a, b)
    assert (a[0][..., 3] == 0).any(), "no blended (pad=0) pixel: transparent pass not exercised"


@pytest.mark.parametrize("cam", [
    [0.0, 0.0, 60.0, 0.0, 0.0, -1.0],        # inside the torus tube's reach: near clipping
    [80.0, 5.0, 20.0, -1.0, 0.0, -0.2],      # grazing view from the side, lateral rejects
    [0.0, 0.0, 450.0, 0.0, 0.0, -1.0],       # far plane cuts the model (far = 510 from camera)
])
def test_clipping_cameras(cam):
    torus = synthetic.torus(12, 8)
    a, b, s1, _ = _render_both([torus], 64, 48, cam=np.array(cam, F))
    _assert_same(a, b)


def test_near_clip_actually_clips():
    torus = synthetic.torus(12, 8)
    s, c = pyoracle.Scene(64, 48), pyoracle.Canvas(64, 48)
    c.init_depth(100000.0)
    s.add_obj(torus)
    s.set_camera([0.0, 0.0, 60.0], [0.0, 0.0, -1.0])
    s.render(c, stats=True)
    st = s.stats()
    assert st["emitted_tris"] != st["input_tris"] - st["culled_tris"]


def test_offset_and_scene_larger_than_canvas():
    """The app renders a monitor-sized Scene into a window-sized Canvas with an offset
    (src/app/mod.rs:67-84); canvas.rs:585-587 subtracts it before snapping."""
    a, b, _, _ = _render_both(load_scene("c2_donut"), 64, 48, offset=(48, 36), scene_wh=(160, 120))
    _assert_same(a, b)


def test_checker_torus_texture_fetch():
    tex = Texture("small", np.ones(3, F), np.ones(3, F), np.full(3, 0.5, F), 1.0,
                  synthetic.checker_texture(64, 8), synthetic.checker_texture(64, 8, seed=7))
    a, b, _, _ = _render_both([synthetic.torus(16, 12, texture=tex)], 96, 72)
    _assert_same(a, b)


def test_painter_sort_ties_moving_camera():
    """Two transparent meshes cut from a symmetric torus (many equal centroid distances: only a stable sort keeps
    their list order), a camera that moves between frames so the persistent in-place order matters
    (scene/mod.rs:1100-1115).  The C++ oracle sorts with std::stable_sort, the numpy one with Python's sort:
    two independent stable sorts that must agree with each other — and the GPU's radix sort with both
    (tests/test_gpu_parity.py::test_painter_sort_on_device_large_meshes_moving_camera)."""
    from draw_b200.model import IndexedMesh, Object
    t = synthetic.torus(10, 8)
    tris = t.meshes[0].triangles
    half = tris.shape[0] // 2 + 3
    glass = Object("glass", t.vertices, t.normals_vertices, t.texture_vertices,
                   [IndexedMesh("a", tris[:half].copy(), 1), IndexedMesh("b", tris[half:].copy(), 2)],
                   [Texture(), Texture(name="ga", alpha=0.6, kd=np.array([0.1, 0.8, 0.3], F)),
                    Texture(name="gb", alpha=0.35, kd=np.array([0.9, 0.2, 0.1], F))])
    W, H = 64, 48
    s1, c1 = pyoracle.Scene(W, H), pyoracle.Canvas(W, H)
    s2, c2 = np_oracle.Scene(W, H), np_oracle.Canvas(W, H)
    for c in (c1, c2):
        c.init_depth(100000.0)
    s1.add_obj(glass)
    s2.add_obj(glass)
    for cam in ([0, 0, 150, 0, 0, -150], [130, 30, 80, -1, -0.2, -0.6], [-90, -60, 110, 0.7, 0.5, -1], [0, 0, 150, 0, 0, -150]):
        cam = np.array(cam, F)
        s1.set_camera(cam[:3], cam[3:])
        s2.set_camera(cam[:3], cam[3:])
        s1.render(c1)
        s2.render(c2)
        # (transparent triangles do not write depth, so only the colour carries the result)
        assert np.array_equal(c1.as_bytes(), c2.frame), "colour bytes differ"
        assert np.array_equal(c1.depth().view(np.uint32), c2.depth_frame.view(np.uint32)), "depth bits differ"
        assert (c1.as_bytes()[..., 3] == 0).any(), "transparent pass not exercised"


def test_draw_triangle_2d_gui_path_cross():
    """Canvas::draw_triangle (canvas.rs:435-575): the C++ oracle and the numpy restatement agree bit for bit on textured,
    vertex-coloured, alpha-blended 2-D triangles drawn in order over a cleared canvas, with and without a clipping rectangle
    and with depth updates on (the first triangle then shuts the others out of its pixels)."""
    from oracle import np_oracle, pyoracle
    rng = np.random.default_rng(7)
    W, H = 40, 30
    tex = rng.integers(0, 256, (8, 16, 4), dtype=np.uint8)
    tex[:2, :, 3] = 255  # some fully opaque texels
    for case in range(6):
        n = 5
        v = np.zeros(3 * n, pyoracle.VERTEX2D)
        v["x"] = rng.uniform(-5, W + 5, 3 * n).astype(np.float32)
        v["y"] = rng.uniform(-5, H + 5, 3 * n).astype(np.float32)
        v["u"] = rng.uniform(0, 0.999, 3 * n).astype(np.float32)
        v["v"] = rng.uniform(0, 0.999, 3 * n).astype(np.float32)
        for ch in "rgb":
            v[ch] = rng.integers(0, 256, 3 * n)
        v["alpha"] = np.where(rng.random(3 * n) < 0.3, 1.0, rng.uniform(0, 1, 3 * n)).astype(np.float32)
        clip = None if case % 2 == 0 else (3, 4, 30, 22)
        c, pc = pyoracle.Canvas(W, H), np_oracle.Canvas(W, H)
        c.init_depth(50.0)
        pc.init_depth(50.0)
        c.clear()
        pc.clear()
        if case >= 4:
            c.enable_depth_update()
            pc.depth_update = True
        c.draw_triangles(v, tex, clip)
        for t in range(n):
            tri = [(float(q["x"]), float(q["y"]), float(q["u"]), float(q["v"]), (int(q["r"]), int(q["g"]), int(q["b"])), float(q["alpha"]))
                   for q in v[3 * t:3 * t + 3]]
            pc.draw_triangle_2d(*tri, tex, clip)
        assert np.array_equal(c.as_bytes(), pc.frame), f"case {case}"
        assert np.array_equal(c.depth().view(np.uint32), pc.depth_frame.view(np.uint32))
        assert (c.as_bytes()[..., 3] == 0).any(), "nothing was drawn"


def test_draw_triangle_2d_golden_gui_frame():
    """The committed GUI frame (tests/golden/make_overlay_golden.py, drawn by the numpy restatement) out of the C++ oracle."""
    import os
    from conftest import GOLDEN
    from draw_b200 import synthetic
    from oracle import pyoracle
    want = np.load(os.path.join(GOLDEN, "overlay_gui_96x64.npz"))["frame"]
    atlas = synthetic.font_atlas(64, 32)
    c = pyoracle.Canvas(96, 64)
    c.init_depth(10.0)
    c.clear()
    for clip, v in synthetic.gui_command_list(96, 64, n_commands=4, quads_per_command=8, seed=2):
        c.draw_triangles(v, atlas, clip)
    assert np.array_equal(c.as_bytes(), want)
