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
    frames = [None] if cfg["cameras"] is None else [cfg["cameras"][k % len(cfg["cameras"])] for k in (0, 30, 60, 90)]
    for cam in frames:
        if cam is not None:
            s.camera = draw_b200.Camera.new(cam[:3], cam[3:])
        s.debug_tile_cycles(enable=True)
        s.render(c)
        s.render(c)
        cyc4 = s.debug_tile_cycles(c)  # rows: whole item, end of phase A, of phase C, of phase D
        cyc = cyc4[0]
        large, ms, transparent = s.debug_list_counts(c)
        order = np.argsort(cyc)[::-1][:8]
        print("  slowest tiles (cycles | large refs, medium/small weight, transparent refs | cycles in A, C, D, E):")
        for t in order:
            print("   ", int(cyc[t]), "|", int(large[t]), int(ms[t]), int(transparent[t]), "|", int(cyc4[1][t]), int(cyc4[2][t]) - int(cyc4[1][t]),
                  int(cyc4[3][t]) - int(cyc4[2][t]), int(cyc[t]) - int(cyc4[3][t]))
        busy = cyc > 0
        print("  item cycles p50/p90/p99/max", [int(np.percentile(cyc[busy], p)) for p in (50, 90, 99, 100)], "sum", int(cyc.sum()), "items", int(busy.sum()))
        for nm, row0, row1 in (("A", None, 1), ("C", 1, 2), ("D", 2, 3), ("E", 3, 0)):
            d = cyc4[row1].astype(np.int64) - (cyc4[row0].astype(np.int64) if row0 is not None else 0)
            print(f"    phase {nm}: p50/p90/max", [int(np.percentile(d[busy], p)) for p in (50, 90, 100)], "share of item cycles", round(float(d[busy].sum()) / max(1, int(cyc.sum())), 3))
        st = c.last_frame_stats()
        print(name, {k: st[k] for k in ("setup_records", "large_refs", "medium_refs", "small_refs", "transparent_refs", "empty_tiles", "work_items")},
              "k_front phases us", [round(x / 1e3, 1) for x in st["front_phase_ns"]],
              "first block (set-up, scan, records, binning, clip) us", [round(x / 1e3, 1) for x in st["front_block_ns"]])
