#!/usr/bin/env python
"""Counts, with the CPU oracle, what SURVEY.md §8(d)'s fill roofline needs per config:
F_cov = fragments that pass the inside test (overdraw included), P_vis = pixels whose final colour comes from
a triangle.  Writes tests/golden/fill_counts.json (means per frame; C4 over its 120-frame camera path).
    python tests/golden/make_fill_counts.py [c2 c3 c4 c5]
Test / measurement infrastructure: bench.py only reads the JSON."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from oracle import pyoracle

out_path = os.path.join(ROOT, "tests", "golden", "fill_counts.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for name in sys.argv[1:] or ["c2", "c3", "c4", "c5"]:
    cfg = bench.load_workload(name)
    W, H = cfg["W"], cfg["H"]
    s, c = pyoracle.Scene(W, H), pyoracle.Canvas(W, H)
    c.init_depth(bench.DEPTH_MAX)
    c.apply_offset(0, 0)
    for o in cfg["objects"]:
        s.add_obj(o)
    cams = cfg["cameras"]
    frames = range(len(cams)) if cams is not None else [None]
    f_cov, p_vis, t0 = [], [], time.time()
    for k in frames:
        if k is not None:
            s.set_camera(cams[k][:3], cams[k][3:])
        s.render(c, stats=True)
        st = s.stats()
        f_cov.append(st["covered_frags"])
        p_vis.append(int((c.depth() < bench.DEPTH_MAX).sum()))
    out[name] = {"workload": cfg["label"], "frames": len(f_cov), "f_cov_mean": float(np.mean(f_cov)), "p_vis_mean": float(np.mean(p_vis)),
                 "f_cov_max": int(max(f_cov)), "bbox_pixels_last": st["bbox_pixels"]}
    print(name, out[name], f"{time.time() - t0:.1f}s", flush=True)
    json.dump(out, open(out_path, "w"), indent=1)
