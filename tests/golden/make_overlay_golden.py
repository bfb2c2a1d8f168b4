#!/usr/bin/env python
"""Writes tests/golden/overlay_gui_96x64.npz: a small GUI command list (draw_b200.synthetic.gui_command_list, seed 2)
drawn by the numpy restatement of Canvas::draw_triangle (oracle/np_oracle.py, canvas.rs:435-575) — the frame the C++
oracle and the CUDA path are both held to.  The reference has no tests or golden images for this path;
the two independently written restatements agreeing is what pins it.   python tests/golden/make_overlay_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from draw_b200 import synthetic  # noqa: E402
from oracle import np_oracle  # noqa: E402

W, H = 96, 64
atlas = synthetic.font_atlas(64, 32)
cmds = synthetic.gui_command_list(W, H, n_commands=4, quads_per_command=8, seed=2)
pc = np_oracle.Canvas(W, H)
pc.init_depth(10.0)
pc.clear()
for clip, v in cmds:
    for k in range(0, len(v), 3):
        tri = [(float(q["x"]), float(q["y"]), float(q["u"]), float(q["v"]), (int(q["r"]), int(q["g"]), int(q["b"])), float(q["alpha"]))
               for q in v[k:k + 3]]
        pc.draw_triangle_2d(*tri, atlas, clip)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "overlay_gui_96x64.npz"), frame=pc.frame)
print("drawn pixels:", int((pc.frame[..., 3] == 0).sum()))
