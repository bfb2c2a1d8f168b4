#!/usr/bin/env python
"""Renders a few small frames through libdraw_b200.so and compares them with the oracle; run it under
compute-sanitizer (tools/sanitize.sh).  Scenes: C1 (textured + transparent, 800x600), C3 at 1280x720 (thousands of small
triangles: key pages, k_raster atomics), C4 frame 60 at 960x544 (near-plane clipping, records covering hundreds of
tiles: k_front's huge-record phase, tile windows), each rendered three times on two canvases with the host mirror on (frames in flight, k_mirror); then a GUI
command list over a rendered frame (k_overlay: bin masks, ordered blend)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import draw_b200  # noqa: E402
from conftest import GOLDEN, load_scene  # noqa: E402
from parity_util import assert_frames_equal, render_oracle  # noqa: E402

path = np.load(GOLDEN + "/c4_camera_path.npy")
for name, (W, H), cam in (("c1_lemur_airplane", (800, 600), None), ("c3_trio", (1280, 720), None), ("c4_dungeon", (960, 544), path[60])):
    objs = load_scene(name)
    s = draw_b200.Scene(W, H)
    for o in objs:
        s.add_obj(o)
    if cam is not None:
        s.camera = draw_b200.Camera.new(cam[:3], cam[3:])
    cs = []
    for _ in range(2):
        c = draw_b200.Canvas(W, H)
        c.init_depth(100000.0)
        c.enable_host_mirror(True)  # the frame's way to the host: whole-frame copy first, k_mirror's changed tiles after
        cs.append(c)
    for k in range(6):
        s.render(cs[k % 2])
        if k >= 2:
            cs[k % 2].as_bytes_slice(copy=False)
    want = render_oracle(objs, W, H, cam=cam, frames=3)
    for c in cs:
        assert_frames_equal((c.as_bytes_slice(), c.depth()), want, name)
    print(name, W, H, "ok", c.last_frame_stats()["setup_records"], "records", flush=True)
# Canvas::draw_triangle: the overlay kernels on top of the last scene's frame
from draw_b200 import synthetic  # noqa: E402
from oracle import pyoracle  # noqa: E402
W, H = 333, 217
atlas = synthetic.font_atlas(64, 32)
c, oc = draw_b200.Canvas(W, H), pyoracle.Canvas(W, H)
c.init_depth(100000.0)
oc.init_depth(100000.0)
c.clear()
oc.clear()
for clip, v in synthetic.gui_command_list(W, H, n_commands=4, quads_per_command=40, seed=5):
    c.draw_triangles(v, atlas, clip)
    oc.draw_triangles(v, atlas, clip)
assert_frames_equal((c.as_bytes_slice(), c.depth()), (oc.as_bytes(), oc.depth()), "gui overlay")
print("gui overlay", W, H, "ok", flush=True)
print("sanitize frames ok")
