"""GPU parity of Canvas::draw_triangle (canvas.rs:435-575, the GUI's 2-D path): draw_canvas_draw_triangles through
the C ABI against the CPU oracle, colour bytes and depth bits exact.  Run on the B200 box: python -m pytest tests -m gpu"""
import numpy as np
import pytest

from conftest import load_scene
from parity_util import DEPTH_MAX, assert_frames_equal

pytestmark = pytest.mark.gpu


def _pair(W, H, depth=DEPTH_MAX):
    import draw_b200
    from oracle import pyoracle
    c, oc = draw_b200.Canvas(W, H), pyoracle.Canvas(W, H)
    c.init_depth(depth)
    oc.init_depth(depth)
    return c, oc


def _draw(c, oc, commands, atlas, dev_tex=None):
    for clip, v in commands:
        c.draw_triangles(v, dev_tex if dev_tex is not None else atlas, clip)
        oc.draw_triangles(v, atlas, clip)


def _check(c, oc, what):
    assert_frames_equal((c.as_bytes_slice(), c.depth()), (oc.as_bytes(), oc.depth()), what)


@pytest.mark.parametrize("W,H", [(640, 360), (333, 217), (1920, 1080), (61, 7)])
def test_gui_command_lists_match_oracle(W, H):
    """Window-like quads, glyph quads, free triangles, degenerate / off-screen / non-finite vertices, clipping
    rectangles (None, inside, swapped corners, beyond the screen), on canvas sizes that are and are not multiples of
    the kernel's tile and bin sizes."""
    import draw_b200
    from draw_b200 import synthetic
    atlas = synthetic.font_atlas()
    tex = draw_b200.DeviceTexture(atlas)
    c, oc = _pair(W, H)
    c.clear()
    oc.clear()
    cmds = synthetic.gui_command_list(W, H, n_commands=6, quads_per_command=30 if W > 1000 else 60, seed=W)
    _draw(c, oc, cmds, atlas, tex)
    _check(c, oc, f"gui {W}x{H}")
    assert (oc.as_bytes()[..., 3] == 0).mean() > 0.05, "the command list drew next to nothing"


def test_order_and_blending_of_overlapping_triangles():
    """Sixty translucent triangles stacked on the same pixels: the blend is order dependent and truncates to u8 at
    every step, so any reordering shows."""
    from draw_b200 import synthetic
    W, H = 200, 120
    rng = np.random.default_rng(5)
    atlas = synthetic.font_atlas(64, 32)
    v = np.zeros(180, synthetic._VERTEX2D)
    v["x"] = rng.uniform(20, 180, 180)
    v["y"] = rng.uniform(10, 110, 180)
    v["u"] = rng.uniform(0, 0.99, 180)
    v["v"] = rng.uniform(0, 0.99, 180)
    for ch in "rgb":
        v[ch] = rng.integers(0, 256, 180)
    v["alpha"] = rng.uniform(0.2, 0.9, 180)
    c, oc = _pair(W, H)
    c.clear()
    oc.clear()
    _draw(c, oc, [(None, v)], atlas)
    _check(c, oc, "stacked")
    # the same triangles one call at a time: stream order between calls is submission order too
    c2, oc2 = _pair(W, H)
    c2.clear()
    for t in range(60):
        c2.draw_triangles(v[3 * t:3 * t + 3], atlas)
    assert np.array_equal(c2.as_bytes_slice(), oc.as_bytes())


def test_depth_update_switch():
    """With depth updates on (enable_depth_update, canvas.rs:399) the first triangle to reach a pixel sets its depth
    to 0.0 and shuts every later one out; off (what Gui::render sets, gui.rs:383) the depth buffer is left alone."""
    from draw_b200 import synthetic
    W, H = 160, 96
    atlas = synthetic.font_atlas(32, 32)
    cmds = synthetic.gui_command_list(W, H, n_commands=3, quads_per_command=40, seed=9)
    for enabled in (True, False):
        c, oc = _pair(W, H, depth=50.0)
        c.clear()
        oc.clear()
        if enabled:
            c.enable_depth_update()
            oc.enable_depth_update()
        else:
            c.disable_depth_update()
            oc.disable_depth_update()
        _draw(c, oc, cmds, atlas)
        _check(c, oc, f"depth update {enabled}")
        assert ((c.depth() == 0.0).any()) == enabled


def test_overlay_on_top_of_a_rendered_frame():
    """The application's frame (src/app/mod.rs): Scene::render, then Gui::render's triangles over it with depth
    updates off; a second render + overlay on the same canvas starts from a clean frame again."""
    import draw_b200
    from draw_b200 import synthetic
    from oracle import pyoracle
    objs = load_scene("c3_trio")
    W, H = 800, 600
    atlas = synthetic.font_atlas()
    c, oc = _pair(W, H)
    s, o = draw_b200.Scene(W, H), pyoracle.Scene(W, H)
    for ob in objs:
        s.add_obj(ob)
        o.add_obj(ob)
    cmds = synthetic.gui_command_list(W, H, n_commands=4, quads_per_command=50, seed=3)
    for frame in range(2):
        s.render(c)
        o.render(oc)
        c.disable_depth_update()
        oc.disable_depth_update()
        _draw(c, oc, cmds, atlas)
        _check(c, oc, f"render + gui, frame {frame}")


def test_more_triangles_than_one_batch():
    """70 000 small triangles in one call (the library cuts batches of 32 768; order holds across the cut) and a
    full-screen pair drawn last that blends over all of them."""
    from draw_b200 import synthetic
    W, H = 512, 256
    rng = np.random.default_rng(11)
    n = 70000
    atlas = synthetic.font_atlas(64, 64)
    v = np.zeros(3 * n + 6, synthetic._VERTEX2D)
    cx, cy = rng.uniform(0, W, n), rng.uniform(0, H, n)
    for k in range(3):
        v["x"][k:3 * n:3] = cx + rng.uniform(-3, 3, n)
        v["y"][k:3 * n:3] = cy + rng.uniform(-3, 3, n)
    v["u"] = rng.uniform(0, 0.99, 3 * n + 6)
    v["v"] = rng.uniform(0, 0.99, 3 * n + 6)
    for ch in "rgb":
        v[ch] = rng.integers(0, 256, 3 * n + 6)
    v["alpha"] = rng.uniform(0.3, 1.0, 3 * n + 6)
    v["x"][3 * n:] = (0, W, W, 0, W, 0)
    v["y"][3 * n:] = (0, 0, H, 0, H, H)
    v["alpha"][3 * n:] = 0.5
    c, oc = _pair(W, H)
    c.clear()
    oc.clear()
    _draw(c, oc, [(None, v)], atlas)
    _check(c, oc, "70k triangles")


def test_errors():
    import draw_b200
    from draw_b200 import synthetic
    atlas = synthetic.font_atlas(16, 16)
    c = draw_b200.Canvas(32, 32)
    v = np.zeros(3, synthetic._VERTEX2D)
    with pytest.raises(draw_b200.DrawError, match="Depth not initialized"):
        c.draw_triangles(v, atlas)
    with pytest.raises(ValueError):
        draw_b200.DeviceTexture(atlas[..., :3])
    c.init_depth(1.0)
    with pytest.raises(ValueError):
        c.draw_triangles(v[:2], atlas)
    c.draw_triangles(v[:0], atlas)  # nothing to draw is not an error


def test_whole_gui_frame_in_one_submission():
    """draw_canvas_draw_commands: the commands of a GUI frame, each with its own clipping rectangle (None, inside, swapped
    corners), in one submission — the frame one draw_triangles call per command gives, and the oracle's."""
    import draw_b200
    from draw_b200 import synthetic
    W, H = 1280, 720
    atlas = synthetic.font_atlas()
    tex = draw_b200.DeviceTexture(atlas)
    cmds = synthetic.gui_command_list(W, H, n_commands=9, quads_per_command=40, seed=21)
    c, oc = _pair(W, H)
    c.clear()
    oc.clear()
    c.draw_commands(cmds, tex)
    for clip, v in cmds:
        oc.draw_triangles(v, atlas, clip)
    _check(c, oc, "one submission")
    c2, _ = _pair(W, H)
    c2.clear()
    for clip, v in cmds:
        c2.draw_triangles(v, tex, clip)
    assert np.array_equal(c2.as_bytes_slice(), c.as_bytes_slice())
    c.draw_commands([], tex)  # nothing to draw
    with pytest.raises(ValueError):
        c.draw_commands([(None, cmds[0][1][:4])], tex)
