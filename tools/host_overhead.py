#!/usr/bin/env python
"""Host cost of the public Python API per frame (run on the GPU box): a tiny canvas and scene, so that the GPU work is
negligible and what is timed is ctypes + numpy + the library's host code."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, draw_b200

cfg = bench.load_workload("c2")
W, H = 128, 64
s = draw_b200.Scene(W, H)
for o in cfg["objects"]:
    s.add_obj(o)
cs = []
for _ in range(3):
    c = draw_b200.Canvas(W, H); c.init_depth(1e5); c.enable_host_mirror(True); cs.append(c)
cam = draw_b200.Camera.new([0.0, 0.0, 150.0], [0.0, 0.0, -150.0])
for k in range(30):
    s.render(cs[k % 3]); cs[k % 3].as_bytes_slice(copy=False)
N = 3000
def t(fn, n=N):
    t0 = time.perf_counter()
    for k in range(n):
        fn(k)
    return (time.perf_counter() - t0) / n * 1e6
print("camera assign           us", round(t(lambda k: setattr(s, "camera", cam)), 2))
print("Camera.new + assign     us", round(t(lambda k: setattr(s, "camera", draw_b200.Camera.new([0.0, 0.0, 150.0], [0.0, 0.0, -150.0]))), 2))
print("render + sync           us", round(t(lambda k: (s.render(cs[0]), cs[0].sync())), 2))
print("render (3 in flight)+as_bytes_slice us", round(t(lambda k: (s.render(cs[k % 3]), cs[(k + 1) % 3].as_bytes_slice(copy=False))), 2))
print("as_bytes_slice (clean)  us", round(t(lambda k: cs[0].as_bytes_slice(copy=False)), 2))
print("last_frame_stats        us", round(t(lambda k: cs[0].last_frame_stats()), 2))
a = cs[0].as_bytes_slice(copy=False)
print("pixel read              us", round(t(lambda k: int(a[H // 2, W // 2, 0])), 2))
