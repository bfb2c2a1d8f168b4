"""Worker of tests/test_gpu_multi.py: sort-first rendering on WORLD_SIZE GPUs must reproduce the
single-GPU frame bit for bit, through the NCCL gather and through the fused peer-store path, with interleaved
tile rows and with contiguous stripes.  (The single-GPU frame is the one tests/test_gpu_parity.py compares with
the oracle, so equality with it is equality with the oracle.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import draw_b200  # noqa: E402
from draw_b200 import multi, scene_cache  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    draw_b200.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    device = torch.device("cuda", local)
    path = np.load(os.path.join(ROOT, "tests", "golden", "c4_camera_path.npy"))
    for scene_name, (W, H), cam in (("c3_trio", (3840, 2160), None), ("c4_dungeon", (1920, 1080), path[50]),
                                    ("c1_lemur_airplane", (800, 600), None)):
        objs = scene_cache.load(os.path.join(ROOT, "tests", "golden", "scenes", scene_name + ".npz"))
        scene = draw_b200.Scene(W, H)
        for o in objs:
            scene.add_obj(o)
        if cam is not None:
            scene.camera = draw_b200.Camera.new(cam[:3], cam[3:])
        ref = None
        if rank == 0:
            full = draw_b200.Canvas(W, H)
            full.init_depth(100000.0)
            scene.render(full)
            ref = full.as_bytes_slice()
        for layout in ("interleaved", "stripes"):
            for mode in ("nccl", "p2p"):
                sf = multi.SortFirst(scene, W, H, dist, device, mode=mode, layout=layout)
                for _ in range(3):  # consecutive frames into the same framebuffer: the flags / stream ordering hold
                    sf.render()
                if rank == 0:
                    got = sf.canvas.as_bytes_slice()
                    assert np.array_equal(got, ref), f"{scene_name} {mode} {layout}: composed frame differs from the single-GPU frame"
                sf.close()
                del sf
        if rank == 0:
            print(f"{scene_name} {W}x{H}: sort-first over {dist.get_world_size()} GPUs == single GPU (nccl, p2p) x (interleaved, stripes)", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
