#!/usr/bin/env python
"""Where a lone frame's time goes (run on the GPU box): python tools/trace_lone.py [c3] [n_frames]
One frame at a time (render, wait), CTA trace on: per frame the first CTA start / last CTA end of k_front, k_raster and
k_tile relative to k_front's first CTA, the gaps between the kernels, and the host-timed frame next to it."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, draw_b200

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cfg = bench.load_workload(name)
W, H = cfg["W"], cfg["H"]
s = draw_b200.Scene(W, H)
for o in cfg["objects"]:
    s.add_obj(o)
c = draw_b200.Canvas(W, H)
c.init_depth(100000.0)
for k in range(20):
    s.render(c)
c.sync()
s.debug_trace(True)
for k in range(10):
    s.render(c)
    c.sync()
rows = []
for k in range(n_frames):
    s.debug_trace(True)
    t = time.perf_counter()
    s.render(c)
    c.sync()
    host_us = (time.perf_counter() - t) * 1e6
    rec = s.debug_trace(True)
    kid = rec[:, 0] & 255
    t0 = rec[:, 2].astype(np.int64); t1 = rec[:, 3].astype(np.int64)
    base = t0.min(); t0 -= base; t1 -= base; t1[t1 < t0] += 1 << 32
    out = [host_us]
    for kk in (1, 2, 3):
        m = kid == kk
        out += [t0[m].min() / 1e3, t1[m].max() / 1e3, np.median(t1[m] - t0[m]) / 1e3] if m.any() else [0, 0, 0]
    rows.append(out)
a = np.median(np.array(rows), axis=0)
print(f"{name}: lone frame, medians over {n_frames} frames (us)")
print(f"  host render+sync {a[0]:.1f}")
print(f"  k_front  first CTA start {a[1]:6.1f}  last CTA end {a[2]:6.1f}  median CTA {a[3]:6.1f}")
print(f"  k_raster first CTA start {a[4]:6.1f}  last CTA end {a[5]:6.1f}  median CTA {a[6]:6.1f}   gap after k_front {a[4] - a[2]:.1f}")
print(f"  k_tile   first CTA start {a[7]:6.1f}  last CTA end {a[8]:6.1f}  median CTA {a[9]:6.1f}   gap after k_raster {a[7] - a[5]:.1f}")
print(f"  device span {a[8]:.1f}; host-timed minus span {a[0] - a[8]:.1f} (graph launch, uniform copy, completion)")
