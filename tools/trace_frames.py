#!/usr/bin/env python
"""Timeline of the frames in flight from the CTA trace (run on the GPU box):
    python tools/trace_frames.py [c3] [n_frames] [own|shared]
Prints, per kernel launch (kernel, work set, order), first CTA start / last CTA end relative to the
window, CTAs, SMs touched, busy SM-time; then per kernel totals and the GPU-wide busy fraction."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, draw_b200

NAMES = ["-", "front", "raster", "tile"]  # CtaTrace kernel ids (k_front.cu, k_raster.cu, k_tile.cu)
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 24
mode = sys.argv[3] if len(sys.argv) > 3 else "shared"
cfg = bench.load_workload(name)
W, H = cfg["W"], cfg["H"]
stream = torch.cuda.Stream()
s = draw_b200.Scene(W, H)
for o in cfg["objects"]:
    s.add_obj(o)
cs = []
for _ in range(4):
    c = draw_b200.Canvas(W, H); c.init_depth(100000.0)
    if mode == "shared":
        c.set_stream(stream.cuda_stream)
    cs.append(c)
for k in range(40):
    s.render(cs[k % 4])
torch.cuda.synchronize()
s.debug_trace(True)
for k in range(8):  # graphs are re-captured with the trace pointers
    s.render(cs[k % 4])
torch.cuda.synchronize()
s.debug_trace(True)
for k in range(n_frames):
    s.render(cs[k % 4])
torch.cuda.synchronize()
rec = s.debug_trace(False)
kid, sm, tag = rec[:, 0] & 255, (rec[:, 0] >> 8) & 0xFFFF, rec[:, 0] >> 24
t0 = rec[:, 2].astype(np.int64); t1 = rec[:, 3].astype(np.int64)
base = t0.min()
t0 -= base; t1 -= base
t1[t1 < t0] += 1 << 32
span = t1.max()
print(f"{name} {mode}: {len(rec)} CTA records, {n_frames} frames in {span / 1e3:.1f} us = {span / 1e3 / n_frames:.2f} us/frame")
# launches: group by (kernel, tag), split in time where gaps between consecutive CTA starts exceed the frame cadence / 2
launches = []
for k in range(len(NAMES)):
    for g in range(8):
        m = (kid == k) & (tag == g)
        if not m.any():
            continue
        order = np.argsort(t0[m]); a0, a1, asm = t0[m][order], t1[m][order], sm[m][order]
        # a set renders every 4th frame: its launches are separated by >= 2 frame times
        cuts = np.where(np.diff(a0) > span / n_frames * 2.0)[0] + 1
        for seg0, seg1 in zip(np.r_[0, cuts], np.r_[cuts, len(a0)]):
            launches.append((a0[seg0:seg1].min(), a1[seg0:seg1].max(), k, g, seg1 - seg0, len(set(asm[seg0:seg1].tolist())),
                             int((a1[seg0:seg1] - a0[seg0:seg1]).sum())))
launches.sort()
print("start_us  end_us  dur_us kernel      set  ctas  sms  cta-time_us")
lo, hi = span * 0.35, span * 0.35 + 4.2 * span / n_frames
for a, b, k, g, n, nsm, busy in launches:
    if lo <= a <= hi:
        print(f"{(a - lo) / 1e3:8.1f} {(b - lo) / 1e3:7.1f} {(b - a) / 1e3:7.1f} {NAMES[k]:10s} {g:4d} {n:5d} {nsm:4d} {busy / 1e3:10.1f}")
print("\nper kernel, per frame: mean launch duration us | CTA-time us | CTA-time / 148 SMs us")
for k in range(len(NAMES)):
    ls = [l for l in launches if l[2] == k]
    if ls:
        print(f"  {NAMES[k]:10s} {np.mean([(l[1] - l[0]) for l in ls]) / 1e3:8.1f} {np.mean([l[6] for l in ls]) / 1e3:10.1f} {np.mean([l[6] for l in ls]) / 1e3 / 148:8.2f}")
# SM busy fraction: union of CTA intervals per SM
busy_total = 0
for m in np.unique(sm):
    iv = sorted(zip(t0[sm == m].tolist(), t1[sm == m].tolist()))
    cur0, cur1 = iv[0]
    for a, b in iv[1:]:
        if a > cur1:
            busy_total += cur1 - cur0; cur0, cur1 = a, b
        else:
            cur1 = max(cur1, b)
    busy_total += cur1 - cur0
print(f"\nSMs with any CTA resident: {100.0 * busy_total / (span * len(np.unique(sm))):.1f}% of the window ({len(np.unique(sm))} SMs seen)")
