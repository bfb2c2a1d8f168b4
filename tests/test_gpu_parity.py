"""GPU parity: libdraw_b200.so (through the C ABI) against the CPU oracle, bit for bit.

Covers BASELINE.json's configs C1-C4 at their full sizes (C5 has its own file), the branches of
the path (clipping, transparency, texture fetch, offsets, stripes) and the edge cases the
reference's arithmetic implies (degenerate and off-screen triangles, coplanar ties, odd canvas
sizes, empty scenes).  Run on the B200 box:  python -m pytest tests -m gpu
"""
import numpy as np
import pytest

from conftest import GOLDEN, load_scene
from parity_util import DEPTH_MAX, assert_frames_equal, render_gpu, render_oracle

pytestmark = pytest.mark.gpu
F = np.float32


def _both(objs, W, H, **kw):
    return render_gpu(objs, W, H, **kw), render_oracle(objs, W, H, **kw)


def test_library_loaded_and_device_present():
    import draw_b200
    assert draw_b200.device_count() >= 1
    assert draw_b200.tile_size() == 32


def test_uniforms_match_oracle():
    import draw_b200
    from oracle import pyoracle
    for (w, h), cam in (((800, 600), None), ((3840, 2160), [30.5, 10.0, 120.25, 0.3, -0.05, -1.0])):
        s, o = draw_b200.Scene(w, h), pyoracle.Scene(w, h)
        if cam:
            s.camera = draw_b200.Camera.new(cam[:3], cam[3:])
            o.set_camera(cam[:3], cam[3:])
        (m1, p1), (m2, p2) = s.uniforms(), o.uniforms()
        assert np.array_equal(m1.view(np.uint32), m2.view(np.uint32))
        assert np.array_equal(p1.view(np.uint32), p2.view(np.uint32))


def test_vertex_visual_matches_oracle():
    import draw_b200
    from oracle import pyoracle
    objs = load_scene("c3_trio")
    s, c = draw_b200.Scene(640, 360), draw_b200.Canvas(640, 360)
    o, oc = pyoracle.Scene(640, 360), pyoracle.Canvas(640, 360)
    c.init_depth(DEPTH_MAX)
    oc.init_depth(DEPTH_MAX)
    for ob in objs:
        s.add_obj(ob)
        o.add_obj(ob)
    s.render(c)
    o.render(oc)
    first = 0
    for i, ob in enumerate(objs):
        n = ob.vertices.shape[0]
        got = s.vertex_visual(c, first, n)            # light3 halfway3 depth
        want = o.vertex_visual(i, n)                  # light3 eye3 halfway3 depth
        want = np.concatenate([want[:, 0:3], want[:, 6:9], want[:, 9:10]], 1)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"object {i}"
        first += n


def test_c2_donut_1080p():
    got, want = _both(load_scene("c2_donut"), 1920, 1080)
    assert_frames_equal(got, want, "C2")


def test_c3_trio_4k():
    got, want = _both(load_scene("c3_trio"), 3840, 2160)
    assert_frames_equal(got, want, "C3")


def test_c1_textured_transparent_800x600_two_frames():
    got, want = _both(load_scene("c1_lemur_airplane"), 800, 600, frames=2)
    assert (want[0][..., 3] == 0).any(), "transparent pass not exercised"
    assert_frames_equal(got, want, "C1")


def test_c1_transparent_moving_camera():
    """The painter sort is persistent and re-done per frame (scene/mod.rs:1100-1115)."""
    cams = [[0, 0, 150, 0, 0, -150], [120, 40, 90, -1, -0.3, -0.8], [-100, -20, 120, 0.8, 0.1, -1], [60, 10, 40, 0, 0, -1]]
    cams = [np.array(c, F) for c in cams]
    got, want = _both(load_scene("c1_lemur_airplane"), 640, 480, cam=cams)
    assert_frames_equal(got, want, "C1 moving")


@pytest.mark.parametrize("k", [0, 45, 100])
def test_c4_dungeon_flythrough_4k(k):
    path = np.load(GOLDEN + "/c4_camera_path.npy")
    got, want = _both(load_scene("c4_dungeon"), 3840, 2160, cam=path[k])
    assert_frames_equal(got, want, f"C4 frame {k}")


def test_c4_dungeon_flythrough_all_frames_small():
    """Every frame of the 120-frame path at 480x270 (heavy near-plane clipping in the middle)."""
    import draw_b200
    from oracle import pyoracle
    path = np.load(GOLDEN + "/c4_camera_path.npy")
    objs = load_scene("c4_dungeon")
    W, H = 480, 270
    s, c = draw_b200.Scene(W, H), draw_b200.Canvas(W, H)
    o, oc = pyoracle.Scene(W, H), pyoracle.Canvas(W, H)
    c.init_depth(DEPTH_MAX)
    oc.init_depth(DEPTH_MAX)
    for ob in objs:
        s.add_obj(ob)
        o.add_obj(ob)
    for k in range(path.shape[0]):
        s.camera = draw_b200.Camera.new(path[k, :3], path[k, 3:])
        o.set_camera(path[k, :3], path[k, 3:])
        s.render(c)
        o.render(oc)
        assert_frames_equal((c.as_bytes_slice(), c.depth()), (oc.as_bytes(), oc.depth()), f"C4 small frame {k}")


@pytest.mark.parametrize("cam", [
    [0.0, 0.0, 60.0, 0.0, 0.0, -1.0],
    [80.0, 5.0, 20.0, -1.0, 0.0, -0.2],
    [0.0, 0.0, 450.0, 0.0, 0.0, -1.0],
    [0.0, 0.0, 0.0, 1.0, 0.2, 0.1],          # camera inside the model
])
def test_clipping_cameras(cam):
    from draw_b200 import synthetic
    got, want = _both([synthetic.torus(48, 32)], 640, 480, cam=np.array(cam, F))
    assert_frames_equal(got, want, f"clip {cam}")


@pytest.mark.parametrize("wh", [(63, 47), (65, 129), (130, 70), (1, 1), (257, 3), (1000, 7)])
def test_odd_canvas_sizes(wh):
    got, want = _both(load_scene("c2_donut"), *wh)
    assert_frames_equal(got, want, f"size {wh}")


def test_offset_and_scene_larger_than_canvas():
    got, want = _both(load_scene("c3_trio"), 800, 600, offset=(560, 240), scene_wh=(1920, 1080))
    assert_frames_equal(got, want, "offset")
    got, want = _both(load_scene("c2_donut"), 300, 200, offset=(-40, 33), scene_wh=(320, 240))
    assert_frames_equal(got, want, "negative offset")


def test_checker_torus_texture_fetch():
    from draw_b200 import synthetic
    tex = synthetic.checker_material()
    got, want = _both([synthetic.torus(64, 48, texture=tex)], 1280, 720)
    assert_frames_equal(got, want, "checker torus")


def test_empty_scene_is_clear_colour():
    got = render_gpu([], 200, 100)
    assert (got[0] == np.array([255, 186, 155, 255], np.uint8)).all()
    assert (got[1] == F(DEPTH_MAX)).all()


def _tri_object(verts, alpha=1.0, kd=(0.4, 0.4, 0.4)):
    from draw_b200.model import IndexedMesh, Object, Texture
    v = np.array(verts, F)
    n = np.tile(np.array([[0, 0, 1]], F), (len(verts), 1))
    t = np.zeros((1, 3), F)
    tris = np.array([[0, 1, 2, 0, 0, 0, 0, 1, 2]], np.uint32)
    return Object("tri", v, n, t, [IndexedMesh("m", tris, 0)], [Texture(alpha=alpha, kd=np.array(kd, F))])


def test_coplanar_ties_first_drawn_wins():
    a = _tri_object([[-50, -50, 0], [50, -50, 0], [0, 50, 0]])
    b = _tri_object([[-50, -50, 0], [50, -50, 0], [0, 50, 0]], kd=(1.0, 0.0, 0.0))
    c = _tri_object([[-60, -40, 0], [40, -60, 0], [10, 60, 0]], kd=(0.0, 1.0, 0.0))
    for order in ([a, b, c], [c, b, a], [b, a, c]):
        got, want = _both(order, 320, 240)
        assert_frames_equal(got, want, "ties")


def test_degenerate_offscreen_and_huge_triangles():
    objs = [
        _tri_object([[-50, 0, 0], [0, 0, 0], [50, 0, 0]]),                 # zero area
        _tri_object([[400, -10, 0], [500, -10, 0], [450, 30, 0]]),         # off-screen right: bbox falls back to column 0
        _tri_object([[-10, 400, 0], [10, 400, 0], [0, 500, 0]]),           # off-screen top: row 0 fallback
        _tri_object([[-1e6, -1e6, -50], [1e6, -1e6, -50], [0, 1e6, -50]]), # huge, covers the screen
        _tri_object([[-30, -30, 10], [30, -30, 10], [0, 30, 10]]),
        _tri_object([[-1e18, -1e18, -50], [1e18, -1e18, -50], [0, 1e18, -50]]),  # overflows f32 products
        _tri_object([[0, 0, 0], [np.nan, 0, 0], [0, 10, 0]]),              # NaN vertex
        _tri_object([[0, 0, 0], [np.inf, 0, 0], [0, 10, 0]]),              # inf vertex
    ]
    for o in objs:
        got, want = _both([o], 200, 150)
        assert_frames_equal(got, want, "edge case")
    got, want = _both(objs, 200, 150)
    assert_frames_equal(got, want, "edge cases together")


def test_transparent_over_opaque_order():
    """Transparent triangles of object i are blended before object i+1's opaque ones are drawn."""
    glass = _tri_object([[-60, -50, 20], [60, -50, 20], [0, 60, 20]], alpha=0.5, kd=(0.0, 0.0, 1.0))
    back = _tri_object([[-80, -60, 0], [80, -60, 0], [0, 80, 0]], kd=(1.0, 1.0, 0.0))
    front = _tri_object([[-20, -20, 40], [20, -20, 40], [0, 20, 40]], kd=(1.0, 0.0, 0.0))
    for order in ([back, glass, front], [glass, back, front], [front, glass, back], [glass, glass, back]):
        got, want = _both(order, 320, 240)
        assert_frames_equal(got, want, "transparent order")


def test_stripes_compose_to_the_full_frame():
    import draw_b200
    objs = load_scene("c3_trio")
    W, H = 1280, 720
    full = render_gpu(objs, W, H)
    scene, canvas = draw_b200.Scene(W, H), draw_b200.Canvas(W, H)
    canvas.init_depth(DEPTH_MAX)
    for o in objs:
        scene.add_obj(o)
    bounds = [0, 192, 448, 640, H]
    for y0, y1 in zip(bounds[:-1], bounds[1:]):
        canvas.set_stripe(y0, y1)
        scene.render(canvas)
    assert_frames_equal((canvas.as_bytes_slice(), canvas.depth()), full, "stripes")


def test_interleaved_tile_rows_compose_to_the_full_frame():
    """draw_canvas_set_tile_rows: the sort-first partition by tile rows ty % step == phase; the passes of all phases
    (here on one GPU, one after the other) compose the full frame, also inside a stripe and with a transparent mesh."""
    import draw_b200
    for name, (W, H), steps in (("c3_trio", (1280, 720), (2, 3, 8)), ("c1_lemur_airplane", (800, 600), (4,)), ("c4_dungeon", (960, 540), (5,))):
        objs = load_scene(name)
        full = render_gpu(objs, W, H)
        scene, canvas = draw_b200.Scene(W, H), draw_b200.Canvas(W, H)
        canvas.init_depth(DEPTH_MAX)
        for o in objs:
            scene.add_obj(o)
        for step in steps:
            canvas.clear()
            for phase in range(step):
                canvas.set_tile_rows(phase, step)
                scene.render(canvas)
            assert_frames_equal((canvas.as_bytes_slice(), canvas.depth()), full, f"{name} rows % {step}")
        # interleaved rows inside two stripes
        canvas.clear()
        for y0, y1 in ((0, 256), (256, H)):
            canvas.set_stripe(y0, y1)
            for phase in range(3):
                canvas.set_tile_rows(phase, 3)
                scene.render(canvas)
        assert_frames_equal((canvas.as_bytes_slice(), canvas.depth()), full, f"{name} rows % 3 in stripes")
    with pytest.raises(draw_b200.DrawError):
        canvas.set_tile_rows(3, 3)


def test_resize_and_clear_semantics():
    import draw_b200
    from oracle import pyoracle
    c, o = draw_b200.Canvas(40, 30), pyoracle.Canvas(40, 30)
    assert (c.as_bytes_slice() == np.array([0, 0, 0, 255], np.uint8)).all()     # Pixel::black()
    c.init_depth(5.0)
    o.init_depth(5.0)
    c.clear()
    o.clear()
    c.resize(50, 40)
    o.resize(50, 40)
    assert np.array_equal(c.as_bytes_slice(), o.as_bytes())
    assert np.array_equal(c.depth(), o.depth())
    c.resize(20, 10)
    o.resize(20, 10)
    assert np.array_equal(c.as_bytes_slice(), o.as_bytes())


def test_errors_do_not_cross_the_boundary():
    import draw_b200
    s, c = draw_b200.Scene(64, 48), draw_b200.Canvas(64, 48)
    with pytest.raises(draw_b200.DrawError, match="Depth not initialized"):
        s.render(c)
    bad = _tri_object([[0, 0, 0], [1, 0, 0], [0, 1, 0]])
    bad.meshes[0].triangles = np.array([[0, 1, 7, 0, 0, 0, 0, 1, 2]], np.uint32)
    with pytest.raises(draw_b200.DrawError, match="out of range"):
        s.add_obj(bad)
    bad2 = _tri_object([[0, 0, 0], [1, 0, 0], [0, 1, 0]])
    bad2.meshes[0].texture_idx = 3
    with pytest.raises(draw_b200.DrawError, match="material index"):
        s.add_obj(bad2)
    with pytest.raises(draw_b200.DrawError):
        c.set_stripe(10, 20)


def test_small_buffers_grow_and_rerender():
    """A clipping-heavy frame that overflows the initial work buffers is re-rendered internally."""
    path = np.load(GOLDEN + "/c4_camera_path.npy")
    objs = load_scene("c4_dungeon")
    (got, scene, canvas) = render_gpu(objs, 3840, 2160, cam=path[60], return_handles=True)
    st = canvas.last_frame_stats()
    assert st["setup_records"] > 0 and st["tile_refs"] > 0 and st["work_items"] > 0
    assert st["tile_refs"] == st["large_refs"] + st["medium_refs"] + st["small_refs"] + st["transparent_refs"]
    want = render_oracle(objs, 3840, 2160, cam=path[60])
    assert_frames_equal(got, want, "C4 frame 60")


def _glass_torus(n_theta=48, n_phi=40):
    """A torus cut into two transparent meshes (3 840 triangles: several 1024-element chunks of the
    device radix sort, and plenty of equal centroid distances from the symmetry, i.e. ties that only a
    stable sort keeps in list order)."""
    from draw_b200 import synthetic
    from draw_b200.model import IndexedMesh, Object, Texture
    t = synthetic.torus(n_theta, n_phi)
    tris = t.meshes[0].triangles
    half = tris.shape[0] // 2 + 17
    meshes = [IndexedMesh("a", tris[:half].copy(), 1), IndexedMesh("b", tris[half:].copy(), 2)]
    tex = [Texture(), Texture(name="ga", alpha=0.6, kd=np.array([0.1, 0.8, 0.3], F)),
           Texture(name="gb", alpha=0.35, kd=np.array([0.9, 0.2, 0.1], F))]
    return Object("glass_torus", t.vertices, t.normals_vertices, t.texture_vertices, meshes, tex)


def test_painter_sort_on_device_large_meshes_moving_camera():
    """scene/mod.rs:1100-1115: per-frame, in-place, stable sort of every transparent mesh (k_sort.cu)."""
    back = _tri_object([[-120, -90, -40], [120, -90, -40], [0, 110, -40]], kd=(1.0, 1.0, 0.0))
    cams = [[0, 0, 150, 0, 0, -150], [0, 0, 150, 0, 0, -150], [130, 30, 80, -1, -0.2, -0.6], [-90, -60, 110, 0.7, 0.5, -1],
            [10, 140, 60, 0, -1, -0.4], [0, 0, 150, 0, 0, -150]]
    cams = [np.array(c, F) for c in cams]
    got, want = _both([back, _glass_torus()], 320, 240, cam=cams)
    assert (want[0][..., 3] == 0).any(), "transparent pass not exercised"
    assert_frames_equal(got, want, "device painter sort")


def test_painter_order_survives_adding_an_object():
    """The sorted order of a transparent mesh persists in the reference's list; a geometry re-upload
    (add_obj between frames) must not reset it."""
    import draw_b200
    from oracle import pyoracle
    W, H = 256, 192
    s, c = draw_b200.Scene(W, H), draw_b200.Canvas(W, H)
    o, oc = pyoracle.Scene(W, H), pyoracle.Canvas(W, H)
    c.init_depth(DEPTH_MAX)
    oc.init_depth(DEPTH_MAX)
    glass = _glass_torus(24, 20)
    s.add_obj(glass)
    o.add_obj(glass)
    for cam in ([120, 40, 90, -1, -0.3, -0.8], [-100, -20, 120, 0.8, 0.1, -1]):
        cam = np.array(cam, F)
        s.camera = draw_b200.Camera.new(cam[:3], cam[3:])
        o.set_camera(cam[:3], cam[3:])
        s.render(c)
        o.render(oc)
    extra = _tri_object([[-80, -60, 0], [80, -60, 0], [0, 80, 0]], kd=(0.2, 0.2, 1.0))
    s.add_obj(extra)
    o.add_obj(extra)
    for _ in range(2):
        s.render(c)
        o.render(oc)
    assert_frames_equal((c.as_bytes_slice(), c.depth()), (oc.as_bytes(), oc.depth()), "order after add_obj")


def test_glass_mesh_of_ten_thousand_triangles():
    """The transparent pass is binned per tile and put in draw order on chip (k_tile phase D): a 10 240-triangle
    glass torus (40 960 ordered slots: forty 1024-slot ordering windows) over an opaque backdrop, small on screen so
    that single tiles hold hundreds of references, and again filling the screen."""
    back = _tri_object([[-120, -90, -40], [120, -90, -40], [0, 110, -40]], kd=(1.0, 1.0, 0.0))
    glass = _glass_torus(80, 64)
    assert glass.triangle_count() == 10240
    cams = [np.array(c, F) for c in ([0, 0, 150, 0, 0, -150], [60, 20, 60, -1, -0.3, -1], [0, 0, 150, 0, 0, -150])]
    got, want = _both([back, glass], 480, 360, cam=cams)
    assert (want[0][..., 3] == 0).any(), "transparent pass not exercised"
    assert_frames_equal(got, want, "10k glass torus")
    got, want = _both([glass, back], 1920, 1080, cam=cams[1:2])
    assert_frames_equal(got, want, "10k glass torus, 1080p")


def test_opaque_scene_then_a_transparent_object_is_added():
    """an opaque-only scene renders (graph replay), a transparent object is added, the next frames take
    the path with the painter sort; the cross-frame ordering event must be usable in both."""
    import draw_b200
    from oracle import pyoracle
    W, H = 320, 240
    s, c = draw_b200.Scene(W, H), draw_b200.Canvas(W, H)
    o, oc = pyoracle.Scene(W, H), pyoracle.Canvas(W, H)
    c.init_depth(DEPTH_MAX)
    oc.init_depth(DEPTH_MAX)
    back = _tri_object([[-120, -90, -40], [120, -90, -40], [0, 110, -40]], kd=(1.0, 1.0, 0.0))
    s.add_obj(back)
    o.add_obj(back)
    for _ in range(3):
        s.render(c)
        o.render(oc)
    assert_frames_equal((c.as_bytes_slice(), c.depth()), (oc.as_bytes(), oc.depth()), "opaque only")
    glass = _glass_torus(24, 20)
    s.add_obj(glass)
    o.add_obj(glass)
    for cam in ([0, 0, 150, 0, 0, -150], [120, 40, 90, -1, -0.3, -0.8], [120, 40, 90, -1, -0.3, -0.8]):
        cam = np.array(cam, F)
        s.camera = draw_b200.Camera.new(cam[:3], cam[3:])
        o.set_camera(cam[:3], cam[3:])
        s.render(c)
        o.render(oc)
    assert_frames_equal((c.as_bytes_slice(), c.depth()), (oc.as_bytes(), oc.depth()), "after adding glass")
