#!/usr/bin/env python
"""Print tile-list statistics of one frame (run on the GPU box): python tools/list_stats.py c3"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, draw_b200

for name in sys.argv[1:] or ["c3"]:
    cfg = bench.load_workload(name)
    s, c = draw_b200.Scene(cfg["W"], cfg["H"]), draw_b200.Canvas(cfg["W"], cfg["H"])
    c.init_depth(100000.0)
    for o in cfg["objects"]:
        s.add_obj(o)
    frames = [None] if cfg["cameras"] is None else [cfg["cameras"][k] for k in (0, 30, 60, 90)]
    for cam in frames:
        if cam is not None:
            s.camera = draw_b200.Camera.new(cam[:3], cam[3:])
        s.debug_tile_cycles(enable=True)
        s.render(c)
        s.render(c)
        cyc3 = s.debug_tile_cycles(c)
        cyc = cyc3[0]
        coarse, medium, fine = s.debug_list_counts(c)
        order = np.argsort(cyc)[::-1][:8]
        print("  slowest tiles (cycles, large n, medium n, small n):", [(int(cyc[t]), int(coarse[t]), int(medium[t]), int(fine[t]), "A/B/rest", int(cyc3[1][t]), int(cyc3[2][t]) - int(cyc3[1][t]), int(cyc[t]) - int(cyc3[2][t])) for t in order])
        print("  cycles p50/p90/p99/max", [int(np.percentile(cyc, p)) for p in (50, 90, 99, 100)], "sum", int(cyc.sum()),
              "empty-tile median", int(np.median(cyc[(coarse == 0)])))
        st = c.last_frame_stats()
        q = lambda a: [int(np.percentile(a, p)) for p in (50, 90, 99, 100)]
        print(name, "records", st["setup_records"], "refs", st["tile_refs"],
              "| medium lists: nonempty", int((medium > 0).sum()), "sum", int(medium.sum()), "max", int(medium.max()),
              "| coarse lists: nonempty", int((coarse > 0).sum()), "of", coarse.size, "sum", int(coarse.sum()), "p50/90/99/max", q(coarse[coarse > 0]) if (coarse > 0).any() else None,
              "| fine lists: nonempty", int((fine > 0).sum()), "of", fine.size, "sum", int(fine.sum()), "p50/90/99/max", q(fine[fine > 0]) if (fine > 0).any() else None)
