"""Known-answer tests for the CPU oracle, derived from the reference source.

The reference has no tests of its own; these values were derived by hand from
scene/mod.rs:297-357 (camera constants), :817-899 (matrix chain) and canvas.rs:51-59,131-133.
"""
import numpy as np
import pytest

from oracle import pyoracle
from draw_b200.model import IndexedMesh, Object, Texture

F = np.float32


def test_default_matrix_800x600():
    s = pyoracle.Scene(800, 600)
    m, planes = s.uniforms()
    expect = np.array([[-165.68542, 0, 399.5, -59925], [0, -165.68542, 299.5, -44925],
                       [0, 0, -1.0400002, 135.60004], [0, 0, 1, -150]], F)
    assert np.array_equal(m, expect)


def _project(m, p):
    v = m @ np.array([*p, 1.0], np.float64)
    return v[0] / v[3], v[1] / v[3]


def test_projected_points_800x600():
    m, _ = pyoracle.Scene(800, 600).uniforms()
    assert np.allclose(_project(m, (0, 0, 0)), (399.5, 299.5), atol=1e-3)
    assert np.allclose(_project(m, (100, 0, 0)), (509.95697, 299.5), atol=1e-3)
    assert np.allclose(_project(m, (0, 100, 0)), (399.5, 409.95694), atol=1e-3)


def test_planes_orientation_and_order():
    """near, far, right, left, top, bottom; inside evaluates positive (scene/mod.rs:603-632)."""
    _, pl = pyoracle.Scene(800, 600).uniforms()
    f = lambda i, p: float(np.dot(pl[i, :3].astype(np.float64), p) + pl[i, 3])
    inside = np.array([0.0, 0.0, 0.0])
    assert all(f(i, inside) > 0 for i in range(6))
    assert f(0, np.array([0, 0, 145.0])) < 0      # nearer than the near plane (camera z=150, near 10)
    assert f(1, np.array([0, 0, -400.0])) < 0     # beyond the far plane (150-510)
    assert f(2, np.array([1000.0, 0, 0])) < 0     # right
    assert f(3, np.array([-1000.0, 0, 0])) < 0    # left
    assert f(4, np.array([0, 1000.0, 0])) < 0     # top
    assert f(5, np.array([0, -1000.0, 0])) < 0    # bottom


def test_clear_colour_and_depth():
    c = pyoracle.Canvas(5, 4)
    c.init_depth(100000.0)
    c.clear()
    b = c.as_bytes()
    assert b.shape == (4, 5, 4)
    assert (b == np.array([255, 186, 155, 255], np.uint8)).all()   # B,G,R,pad (canvas.rs:51-59,131-133)
    assert (c.depth() == F(100000.0)).all()


def _one_triangle(verts, alpha=1.0, uv=None):
    v = np.array(verts, F)
    n = np.tile(np.array([[0, 0, 1]], F), (3, 1))
    t = np.zeros((3, 3), F) if uv is None else np.array(uv, F)
    tris = np.array([[0, 1, 2, 0, 1, 2, 0, 1, 2]], np.uint32)
    return Object("tri", v, n, t, [IndexedMesh("m", tris, 0)], [Texture(alpha=alpha)])


def test_back_face_is_culled_and_front_face_drawn():
    front = _one_triangle([[-50, -50, 0], [50, -50, 0], [0, 50, 0]])     # CCW seen from +z
    back = _one_triangle([[-50, -50, 0], [0, 50, 0], [50, -50, 0]])
    for obj, expect in ((front, True), (back, False)):
        s, c = pyoracle.Scene(80, 60), pyoracle.Canvas(80, 60)
        c.init_depth(100000.0)
        s.add_obj(obj)
        s.render(c, stats=True)
        assert (s.stats()["covered_frags"] > 0) == expect


def test_y_flip_colour_but_not_depth():
    """canvas.rs:941-956 vs :413-423: colour rows are flipped, depth rows are not."""
    tri = _one_triangle([[-20, 20, 0], [20, 20, 0], [0, 60, 0]])         # upper half (world y up)
    s, c = pyoracle.Scene(80, 60), pyoracle.Canvas(80, 60)
    c.init_depth(100000.0)
    s.add_obj(tri)
    s.render(c)
    drawn_depth_rows = np.where((c.depth() < 100000).any(axis=1))[0]
    clear = np.array([255, 186, 155, 255], np.uint8)
    drawn_colour_rows = np.where((c.as_bytes() != clear).any(axis=(1, 2)))[0]
    assert drawn_depth_rows.min() > 30                                   # canvas y grows upward
    assert drawn_colour_rows.max() < 30                                  # image row 0 is the top
    assert np.array_equal(np.sort(59 - drawn_depth_rows), drawn_colour_rows)


def test_shared_edge_drawn_once():
    """The (-1,-1) tie-break (canvas.rs:677-680): two triangles sharing an edge never both
    cover a pixel, and together cover the quad without holes."""
    quad = np.array([[-40, -30, 0], [40, -30, 0], [40, 30, 0], [-40, 30, 0]], F)
    n = np.tile(np.array([[0, 0, 1]], F), (4, 1))
    t = np.zeros((1, 3), F)
    tris = np.array([[0, 1, 2, 0, 0, 0, 0, 1, 2], [0, 2, 3, 0, 0, 0, 0, 2, 3]], np.uint32)
    obj = Object("quad", quad, n, t, [IndexedMesh("m", tris, 0)], [Texture()])
    s, c = pyoracle.Scene(160, 120), pyoracle.Canvas(160, 120)
    c.init_depth(100000.0)
    s.add_obj(obj)
    s.render(c, stats=True)
    st = s.stats()
    covered = int((c.depth() < 100000).sum())
    assert st["covered_frags"] == covered            # no pixel covered twice
    ys, xs = np.where(c.depth() < 100000)
    assert covered == (xs.max() - xs.min() + 1) * (ys.max() - ys.min() + 1)   # no holes


def test_equal_depth_first_drawn_wins():
    """Strict `<` (canvas.rs:923): a later coplanar triangle does not overwrite."""
    a = _one_triangle([[-50, -50, 0], [50, -50, 0], [0, 50, 0]])
    b = _one_triangle([[-50, -50, 0], [50, -50, 0], [0, 50, 0]])
    b.textures = [Texture(kd=np.array([1.0, 0.0, 0.0], F))]
    s, c = pyoracle.Scene(80, 60), pyoracle.Canvas(80, 60)
    c.init_depth(100000.0)
    s.add_obj(a)
    s.add_obj(b)
    s.render(c)
    w = c.winner()
    assert set(np.unique(w)) == {0, 0xFFFFFFFF}      # draw id 0 = first triangle, never id 4


def test_degenerate_triangle_draws_nothing():
    tri = _one_triangle([[-50, 0, 0], [0, 0, 0], [50, 0, 0]])
    s, c = pyoracle.Scene(80, 60), pyoracle.Canvas(80, 60)
    c.init_depth(100000.0)
    s.add_obj(tri)
    s.render(c, stats=True)
    assert s.stats()["written_frags"] == 0


@pytest.mark.parametrize("cam_z,expect_emitted", [(150.0, 1), (5.0, 2), (-700.0, 0)])
def test_near_clip_emits_0_1_2(cam_z, expect_emitted):
    """ViewPlane::clip emits 0/1/2 triangles (scene/mod.rs:662-746)."""
    # front-facing for a camera on +z looking down -z; one vertex pokes through the near plane
    tri = _one_triangle([[-30, -10, -20], [30, -10, -20], [0, 10, 0]])
    s, c = pyoracle.Scene(80, 60), pyoracle.Canvas(80, 60)
    c.init_depth(100000.0)
    s.add_obj(tri)
    s.set_camera([0, 0, cam_z], [0, 0, -1])
    s.render(c, stats=True)
    st = s.stats()
    assert st["emitted_tris"] == expect_emitted
