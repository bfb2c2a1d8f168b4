"""The alternative code paths of the library, each bit-exact against the oracle.

The defaults (CUDA-graph replay of the frame, k_front with one CTA per SM for small scenes, tile windows) are
what every other GPU test runs.  The knobs that select the other paths are read once when the library is loaded,
so each variant runs a subset of the parity tests in a fresh interpreter:

  DRAW_B200_GRAPH=0        direct kernel launches instead of graph replay
  DRAW_B200_SPLIT_MAX=1    no tile windows
  DRAW_B200_FRONT_CPS=4    k_front with four CTAs per SM (what large scenes get) on the small test scenes
  DRAW_B200_TILE_CTAS / DRAW_B200_RASTER_CTAS / DRAW_B200_SETS   small grids, two frames in flight
  DRAW_B200_REC_CAP / DRAW_B200_REFS_CAP   tiny initial work buffers: every first frame overflows and is re-rendered
"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SUBSET = "c2 or c3 or c1_textured or clipping or odd or ties or degenerate or stripes or checker or small_buffers or glass"


@pytest.mark.parametrize("env", [{"DRAW_B200_GRAPH": "0"}, {"DRAW_B200_SPLIT_MAX": "1"}, {"DRAW_B200_FRONT_CPS": "4"},
                                 {"DRAW_B200_TILE_CTAS": "37", "DRAW_B200_RASTER_CTAS": "19", "DRAW_B200_SETS": "2"},
                                 {"DRAW_B200_REC_CAP": "1500", "DRAW_B200_REFS_CAP": "3000"}],
                         ids=["no_graph", "no_windows", "front_4_per_sm", "few_ctas_two_sets", "tiny_buffers"])
def test_alternative_paths_are_bit_exact(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q",
                        "-k", SUBSET], env=e, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_forced_overflow_rerenders_from_the_frame_s_own_inputs():
    """Tiny work buffers: the overflow flags are raised (asserted), the frame is re-rendered with grown buffers from
    the camera / stripe it was rendered with — also when those have changed before the frame is read."""
    e = dict(os.environ)
    e.update({"DRAW_B200_REC_CAP": "600", "DRAW_B200_REFS_CAP": "900"})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "overflow_worker.py")], env=e, capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "overflow worker ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_host_mirror_follows_every_render():
    """draw_canvas_enable_host_mirror: the pinned frame after a render equals an explicit read-back."""
    import numpy as np
    import draw_b200
    from conftest import load_scene
    objs = load_scene("c2_donut")
    s = draw_b200.Scene(640, 360)
    for o in objs:
        s.add_obj(o)
    a, b = draw_b200.Canvas(640, 360), draw_b200.Canvas(640, 360)
    for c in (a, b):
        c.init_depth(100000.0)
    a.enable_host_mirror(True)
    for cam in ([0.0, 0.0, 150.0, 0.0, 0.0, -150.0], [20.0, 5.0, 120.0, -0.2, 0.0, -1.0]):
        s.camera = draw_b200.Camera.new(cam[:3], cam[3:])
        s.render(a)
        s.render(b)
        assert np.array_equal(a.as_bytes_slice(), b.as_bytes_slice())


def test_export_png_is_the_frame_with_b_and_r_swapped(tmp_path):
    """draw_canvas_export_png = Application::export_frame_as(Png) (app/mod.rs:316-360): RGBA file of the BGRA frame."""
    import numpy as np
    import draw_b200
    from conftest import load_scene
    s, c = draw_b200.Scene(400, 300), draw_b200.Canvas(400, 300)
    c.init_depth(100000.0)
    for o in load_scene("c1_lemur_airplane"):
        s.add_obj(o)
    s.render(c)
    p = str(tmp_path / "frame.png")
    c.export_png(p)
    frame = c.as_bytes_slice()
    assert np.array_equal(draw_b200.load_image(p), frame[..., [2, 1, 0, 3]])


def test_export_jpeg_decodes_to_the_frame(tmp_path):
    """draw_canvas_export_jpeg = Application::export_frame_as(Jpeg) (app/mod.rs:316-378): the quality the reference ends
    up with is 100 (it passes width * 4, clamped), so the file decodes — with the library's own decoder — to within the
    rounding of a quantiser-1 DCT round trip of the frame's R, G, B."""
    import numpy as np
    import draw_b200
    from conftest import load_scene
    s, c = draw_b200.Scene(401, 299), draw_b200.Canvas(401, 299)
    c.init_depth(100000.0)
    for o in load_scene("c1_lemur_airplane"):
        s.add_obj(o)
    s.render(c)
    p = str(tmp_path / "frame.jpg")
    c.export_jpeg(p)
    frame = c.as_bytes_slice()
    got = draw_b200.load_image(p)
    assert got.shape == (299, 401, 3)
    d = np.abs(got.astype(int) - frame[..., [2, 1, 0]].astype(int))
    assert d.max() <= 4 and d.mean() < 0.6
