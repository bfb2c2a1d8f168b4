"""Host enqueue cost vs GPU time per frame (is the timed region host-bound?)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import draw_b200, bench
for name in ("empty1080", "empty4k", "c2", "c3"):
    if name.startswith("empty"):
        W, H = (1920, 1080) if name == "empty1080" else (3840, 2160)
        objs = []
    else:
        cfg = bench.load_workload(name)
        W, H, objs = cfg["W"], cfg["H"], cfg["objects"]
    s = draw_b200.Scene(W, H)
    for o in objs: s.add_obj(o)
    cs = []
    for _ in range(5):
        c = draw_b200.Canvas(W, H); c.init_depth(100000.0); cs.append(c)
    for k in range(20): s.render(cs[k % 5])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(400): s.render(cs[k % 5])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(name, "host enqueue us/frame", round(1e6 * (t1 - t0) / 400, 2), "total us/frame", round(1e6 * (t2 - t0) / 400, 2), flush=True)
