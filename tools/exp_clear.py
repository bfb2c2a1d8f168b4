import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import draw_b200, bench
W, H = 3840, 2160
s = draw_b200.Scene(W, H)
cs = []
for _ in range(4):
    c = draw_b200.Canvas(W, H); c.init_depth(100000.0); cs.append(c)
s.set_kernel_timing(True)
for name in ("empty", "c3"):
    if name == "c3":
        for o in bench.load_workload("c3")["objects"]: s.add_obj(o)
    t = []
    for k in range(12):
        s.render(cs[k % 4]); t.append(s.last_kernel_times(cs[k % 4])["k_tile"])
    print(name, "k_tile us", [round(1e3 * x, 1) for x in t[4:]])
buf = torch.empty(8 * W * H, dtype=torch.uint8, device="cuda")
for _ in range(3): buf.fill_(1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10): buf.fill_(7)
b.record(); torch.cuda.synchronize()
print("torch fill 66MB us", 1e3 * a.elapsed_time(b) / 10)
