"""Host-side logic of the multi-GPU drivers, on CPU: stripe partition, byte ranges, and a
world_size-2 gloo run of the stripe gather and the frame-parallel assignment."""
import os
import socket
import subprocess
import sys

from conftest import ROOT
from draw_b200 import multi


def test_stripe_bounds_cover_the_canvas():
    for H in (2160, 4320, 600, 33, 32, 1):
        for world in (1, 2, 4, 8):
            b = multi.stripe_bounds(H, world)
            assert len(b) == world
            live = [s for s in b if s[1] > s[0]]
            assert live[0][0] == 0 and live[-1][1] == H
            assert all(x[1] == y[0] for x, y in zip(live, live[1:]))
            assert all(s[0] % 32 == 0 for s in live)
            rows = [(s[1] - s[0] + 31) // 32 for s in live]
            assert max(rows) - min(rows) <= 1


def test_stripe_byte_range_is_flipped_and_contiguous():
    H, W = 2160, 3840
    b = multi.stripe_bounds(H, 8)
    ranges = [multi.stripe_byte_range(H, W, *s) for s in b]
    assert ranges[0][1] == H * W * 4 and ranges[-1][0] == 0          # canvas row 0 is the LAST frame row
    assert all(x[0] == y[1] for x, y in zip(ranges, ranges[1:]))


def test_interleaved_rows_partition_the_canvas():
    for H in (2160, 4320, 600, 33, 32, 1):
        for world in (1, 2, 4, 8):
            rows = sorted(sum((multi.interleaved_rows(H, world, r) for r in range(world)), []))
            assert rows[0][0] == 0 and rows[-1][1] == H and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            counts = [len(multi.interleaved_rows(H, world, r)) for r in range(world)]
            assert max(counts) - min(counts) <= 1
            for r in range(world):
                assert all((y0 // 32) % world == r for y0, _ in multi.interleaved_rows(H, world, r))
            assert multi.rank_blocks(H, world, 0, "stripes") == ([multi.stripe_bounds(H, world)[0]] if multi.stripe_bounds(H, world)[0][1] else [])


def test_frames_of_rank():
    assert multi.frames_of_rank(7, 4, 1) == [1, 5]
    assert sum((multi.frames_of_rank(120, 8, r) for r in range(8)), []).__len__() == 120


def test_gloo_world_size_2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "gloo_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=180, env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
