#!/usr/bin/env python
"""Aggregate an ncu report's warp-stall samples per CUDA source line.

    python tools/ncu_hot_lines.py gpurun_out/prof.ncu-rep [draw_b200/libdraw_b200.so] [kernel-substring]

ncu's `--page source --csv` lists samples per SASS instruction; nvdisasm --print-line-info maps
SASS offsets to file:line (the library is built with -lineinfo).  The two are joined on the
instruction's offset inside the function.
"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "draw_b200/libdraw_b200.so"
kern = sys.argv[3] if len(sys.argv) > 3 else "k_tile"

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
ci = {n: i for i, n in enumerate(h)}
body = [r for r in rows[2:] if len(r) > ci["# Samples"] and r[ci["# Samples"]].isdigit()]
base = int(body[0][ci["Address"]], 16)
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, inside = None, None, False
    for ln in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            inside = kern in cur_fn and "$" not in cur_fn.split(kern)[-1][:0]
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m and cur_fn and kern in cur_fn:
            line_of[(cur_fn, int(m.group(1), 16))] = cur_line

# the kernel's own function name is the one whose instruction count matches best; subfunctions
# ($-suffixed) follow it in the same .text section in ncu's listing, so offsets are contiguous
fns = sorted({k[0] for k in line_of})
main = [f for f in fns if "$" not in f]
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
tot = 0
for r in body:
    off = int(r[ci["Address"]], 16) - base
    s = int(r[ci["# Samples"]])
    tot += s
    key = None
    for fn in main:
        key = line_of.get((fn, off))
        if key:
            break
    key = key or ("?", off // 0x400)
    a = agg[key]
    a[0] += s
    a[1] += int(r[ci["Instructions Executed"]] or 0)
    for i in stall_cols:
        if r[i] and r[i] != "0":
            a[2][h[i]] += int(r[i])
print(f"total samples {tot}, instructions {len(body)}")
for key, (s, ie, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    top = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * s / tot:5.1f}%  {s:6d} samples  {ie:9d} inst  {key[0]}:{key[1]}   [{top}]")
