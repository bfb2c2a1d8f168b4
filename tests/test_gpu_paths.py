"""The alternative code paths of the library, each bit-exact against the oracle.

The defaults (CUDA-graph replay, k_raster key pages, in-tile shading) are what every other GPU test
runs.  The knobs that select the other paths are read once when the library is loaded, so each
variant runs a subset of the parity tests in a fresh interpreter:

  DRAW_B200_PAGES=0        no key pages: k_tile rasterises the medium / small lists itself (phases B1, B2)
  DRAW_B200_DEFER_MAX=16   k_shade resolves the key pages of tiles with few large triangles
  DRAW_B200_GRAPH=0        direct kernel launches instead of graph replay (PDL between the kernels)
  DRAW_B200_SPLIT_MAX=1    no tile windows
  DRAW_B200_CLEAR_IN_TILE=0  the empty tiles are written by k_clear_empty on its own stream (default: by k_tile's
                             CTAs after each raster item, mode 2; 1 = before the item, 3 = alternating)
  DRAW_B200_BIN_RPW=0      k_bin always in thread-per-record mode (default: warp per record for small scenes)
"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SUBSET = "c2 or c3 or c1_textured or clipping or odd or ties or degenerate or stripes or checker or small_buffers"


@pytest.mark.parametrize("env", [{"DRAW_B200_PAGES": "0"}, {"DRAW_B200_DEFER_MAX": "16"}, {"DRAW_B200_GRAPH": "0"},
                                 {"DRAW_B200_SPLIT_MAX": "1", "DRAW_B200_DEFER_MAX": "1"},
                                 {"DRAW_B200_CLEAR_IN_TILE": "0", "DRAW_B200_BIN_RPW": "0"},
                                 {"DRAW_B200_CLEAR_IN_TILE": "3", "DRAW_B200_TILE_CTAS": "37", "DRAW_B200_SETS": "2"}],
                         ids=["no_pages", "k_shade", "no_graph", "no_windows_raster_only_deferred", "k_clear_empty_thread_bin",
                              "clear_alternating_few_ctas"])
def test_alternative_paths_are_bit_exact(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q",
                        "-k", SUBSET], env=e, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


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
