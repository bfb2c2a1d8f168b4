import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import draw_b200, bench
for name in ("c2", "c3"):
    cfg = bench.load_workload(name)
    s, c = draw_b200.Scene(cfg["W"], cfg["H"]), draw_b200.Canvas(cfg["W"], cfg["H"])
    c.init_depth(100000.0)
    for o in cfg["objects"]: s.add_obj(o)
    s.debug_tile_cycles(enable=True)
    s.render(c); s.render(c)
    cyc = s.debug_tile_cycles(c)
    flat = np.concatenate([np.asarray(a).ravel() for a in cyc])
    print(name, "per CTA [prologue, first fetch, shade total, n units, whole, smid, t0]:")
    for b in (0, 1, 2, 30, 63): print("   ", flat[8*b:8*b+7].tolist())
