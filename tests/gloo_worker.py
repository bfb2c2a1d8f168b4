"""Worker of tests/test_multi_cpu.py: world_size-2 checks of the multi-GPU host logic on gloo/CPU.
Run under torch.distributed.run; exits non-zero on failure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from draw_b200 import multi  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H = 96, 200  # 7 tile rows of 32: stripes of 4 and 3 rows, the last one ragged (200 = 6*32 + 8)
    bounds = multi.stripe_bounds(H, world)
    assert bounds[0][0] == 0 and bounds[-1][1] == H and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))

    # every rank "renders" only its stripe: canvas row y of rank r holds the byte (y * 7 + r) & 255
    frame = torch.zeros(H * W * 4, dtype=torch.uint8)
    y0, y1 = bounds[rank]
    img = frame.view(H, W, 4)
    for y in range(y0, y1):
        img[H - 1 - y] = (y * 7 + rank) & 255          # colour rows are y-flipped
    multi.gather_stripes(dist, frame, bounds, H, W, root=0)
    if rank == 0:
        for r, (a, b) in enumerate(bounds):
            for y in range(a, b):
                assert int(img[H - 1 - y, 0, 0]) == (y * 7 + r) & 255, (r, y)
                assert (img[H - 1 - y] == img[H - 1 - y, 0, 0]).all()

    # the same with interleaved tile rows (rank r owns rows ty % world == r): one grouped send/recv of row blocks
    frame2 = torch.zeros(H * W * 4, dtype=torch.uint8)
    img2 = frame2.view(H, W, 4)
    blocks = lambda r: multi.interleaved_rows(H, world, r)
    for y0, y1 in blocks(rank):
        for y in range(y0, y1):
            img2[H - 1 - y] = (y * 5 + rank) & 255
    multi.gather_blocks(dist, frame2, blocks, H, W, root=0)
    if rank == 0:
        for r in range(world):
            for a, b in blocks(r):
                for y in range(a, b):
                    assert int(img2[H - 1 - y, 0, 0]) == (y * 5 + r) & 255, (r, y)
                    assert (img2[H - 1 - y] == img2[H - 1 - y, 0, 0]).all()

    # the packed form of the same gather (one collective: what the NCCL path of the interleaved layout uses)
    frame3 = torch.zeros(H * W * 4, dtype=torch.uint8)
    img3 = frame3.view(H, W, 4)
    for y0, y1 in blocks(rank):
        for y in range(y0, y1):
            img3[H - 1 - y] = (y * 3 + rank) & 255
    rg = multi.RowGather(dist, H, W, blocks, torch.device("cpu"))
    rg.gather(frame3)
    rg.gather(frame3)  # reusable
    if rank == 0:
        for r in range(world):
            for a, b in blocks(r):
                for y in range(a, b):
                    assert int(img3[H - 1 - y, 0, 0]) == (y * 3 + r) & 255, (r, y)
                    assert (img3[H - 1 - y] == img3[H - 1 - y, 0, 0]).all()

    # frame-parallel assignment covers every frame exactly once
    mine = multi.frames_of_rank(10, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    assert sorted(sum(gathered, [])) == list(range(10))

    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
