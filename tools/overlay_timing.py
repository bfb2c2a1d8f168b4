#!/usr/bin/env python
"""Times Canvas::draw_triangle batches on the GPU box against the CPU oracle:  python tools/overlay_timing.py [W H]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import draw_b200
from draw_b200 import synthetic
from oracle import pyoracle

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
atlas = synthetic.font_atlas()
tex = draw_b200.DeviceTexture(atlas)
cmds = synthetic.gui_command_list(W, H, n_commands=12, quads_per_command=200, seed=4)
n_tri = sum(len(v) // 3 for _, v in cmds)
c, oc = draw_b200.Canvas(W, H), pyoracle.Canvas(W, H)
c.init_depth(1e5)
oc.init_depth(1e5)
for rep in range(3):
    c.clear()
    c.sync()
    t0 = time.perf_counter()
    for clip, v in cmds:
        c.draw_triangles(v, tex, clip)
    c.sync()
    t_gpu = time.perf_counter() - t0
for rep in range(3):
    c.clear()
    c.sync()
    t0 = time.perf_counter()
    c.draw_commands(cmds, tex)
    c.sync()
    t_one = time.perf_counter() - t0
oc.clear()
t0 = time.perf_counter()
for clip, v in cmds:
    oc.draw_triangles(v, atlas, clip)
t_cpu = time.perf_counter() - t0
same = np.array_equal(c.as_bytes_slice(), oc.as_bytes())
print(f"{W}x{H}: {len(cmds)} commands, {n_tri} triangles: GPU {t_gpu * 1e3:.3f} ms (host-timed, vertices copied from pageable memory), "
      f"as one submission (draw_commands) {t_one * 1e3:.3f} ms, oracle {t_cpu * 1e3:.1f} ms, x{t_cpu / t_one:.0f}, bit-exact {same}")
