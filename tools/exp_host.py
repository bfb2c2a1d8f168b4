import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import draw_b200, bench
for name in ("c2", "c3"):
    cfg = bench.load_workload(name)
    W, H = cfg["W"], cfg["H"]
    s = draw_b200.Scene(W, H)
    for o in cfg["objects"]: s.add_obj(o)
    cs = []
    for _ in range(4):
        c = draw_b200.Canvas(W, H); c.init_depth(100000.0); cs.append(c)
    for k in range(20): s.render(cs[k % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(200): s.render(cs[k % 4])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(name, "host enqueue us/frame", 1e6 * (t1 - t0) / 200, "total us/frame", 1e6 * (t2 - t0) / 200)
