"""Multi-GPU drivers: one process per GPU, torch.distributed for the plumbing.

The reference is single-process and single-threaded; it has no multi-device path.  Two
partitions of the render path exist here:

  frame-parallel   a sequence of frames is dealt round-robin to the ranks; every rank holds the
                   whole scene and renders whole frames.  No collective on the data path.
  sort-first       ONE frame is split into horizontal stripes of tile rows; every rank runs the
                   (cheap) vertex/setup stages for the whole scene and rasterises only its stripe.
                   Stripes are disjoint, so compositing needs no depth compare.  Two ways to land
                   the stripes in rank 0's framebuffer:
                     "nccl"  each rank renders into its own canvas, then send/recv of the stripe
                             rows to rank 0 (torch.distributed batch_isend_irecv over NCCL/NVLink);
                     "p2p"   rank 0 exports its colour buffer through CUDA IPC; the other ranks'
                             tile kernels store their pixels straight into it over NVLink (peer
                             stores from k_tile), so the gather is fused into the raster kernel and
                             only a barrier remains.

The colour buffer is y-flipped (canvas.rs:955-956): canvas rows [y0, y1) live in frame rows
[H - y1, H - y0), still one contiguous byte range.
"""
import ctypes as C

import numpy as np

TILE_H = 32  # draw_tile_size(); kept here so the partition logic is testable without the library


def stripe_bounds(height, world, tile_h=TILE_H):
    """Split `height` canvas rows into `world` contiguous stripes of whole tile rows.
    Returns [(y0, y1)] per rank; a rank with no tile row gets (0, 0) — e.g. more ranks than rows."""
    rows = (height + tile_h - 1) // tile_h
    base, extra = divmod(rows, world)
    out, r = [], 0
    for i in range(world):
        n = base + (1 if i < extra else 0)
        y0, y1 = r * tile_h, min((r + n) * tile_h, height)
        out.append((y0, y1) if n > 0 else (0, 0))
        r += n
    return out


def stripe_byte_range(height, width, y0, y1):
    """Byte range of canvas rows [y0, y1) inside the BGRA8 frame (rows are y-flipped)."""
    return (height - y1) * width * 4, (height - y0) * width * 4


def frames_of_rank(n_frames, world, rank):
    """Frame-parallel assignment: frame k goes to rank k mod world."""
    return list(range(rank, n_frames, world))


def gather_stripes(dist, frame, bounds, height, width, root=0):
    """Gather every rank's stripe of `frame` (flat uint8 tensor of H*W*4 bytes, same size on every
    rank) into the root's `frame`.  Works with any backend (NCCL on GPU tensors, gloo on CPU)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = []
    if rank == root:
        for r in range(world):
            y0, y1 = bounds[r]
            if r == root or y1 <= y0:
                continue
            b0, b1 = stripe_byte_range(height, width, y0, y1)
            ops.append(dist.P2POp(dist.irecv, frame[b0:b1], r))
    else:
        y0, y1 = bounds[rank]
        if y1 > y0:
            b0, b1 = stripe_byte_range(height, width, y0, y1)
            ops.append(dist.P2POp(dist.isend, frame[b0:b1], root))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class _DevicePtr:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def canvas_color_tensor(canvas, device):
    import torch
    ptr, _ = canvas.device_ptrs()
    return torch.as_tensor(_DevicePtr(ptr, canvas.width * canvas.height * 4), device=device)


class SortFirst:
    """One frame across all ranks.  `mode` is "nccl" or "p2p" (see module docstring)."""

    def __init__(self, scene, width, height, dist, device, mode="nccl", depth_max=100000.0):
        import draw_b200
        from . import _native as N
        self.scene, self.dist, self.mode = scene, dist, mode
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.width, self.height = width, height
        self.bounds = stripe_bounds(height, self.world, draw_b200.tile_size())
        self.canvas = draw_b200.Canvas(width, height)
        self.canvas.init_depth(depth_max)
        self.canvas.apply_offset(0, 0)
        self.y0, self.y1 = self.bounds[self.rank]
        if self.y1 > self.y0:
            self.canvas.set_stripe(self.y0, self.y1)
        self.frame = canvas_color_tensor(self.canvas, device)
        self._peer = None
        if mode == "p2p":
            import torch
            handle = torch.zeros(64, dtype=torch.uint8, device=device)
            export_error = None
            if self.rank == 0:
                buf = (C.c_uint8 * 64)()
                try:
                    N.check(N.lib().draw_canvas_ipc_export(self.canvas._h, buf))
                    handle.copy_(torch.tensor(list(buf), dtype=torch.uint8))
                except Exception as e:  # every rank must still reach the broadcast: an all-zero handle says "no"
                    export_error = e
            dist.broadcast(handle, 0)
            if export_error is not None:
                raise export_error
            if self.rank != 0:
                if not bool(handle.any().item()):
                    raise RuntimeError("rank 0 could not export its framebuffer through CUDA IPC")
                raw = (C.c_uint8 * 64)(*handle.cpu().tolist())
                peer = C.c_void_p()
                N.check(N.lib().draw_ipc_open(raw, C.byref(peer)))
                self._peer = peer
                _, own_depth = self.canvas.device_ptrs()
                self.canvas.bind_external(peer.value, own_depth)  # colour -> rank 0's framebuffer over NVLink

    def close(self):
        if self._peer is not None:
            from . import _native as N
            self.canvas.bind_external(None, None)
            N.lib().draw_ipc_close(self._peer)
            self._peer = None

    def render(self):
        """Render this rank's stripe and land all stripes in rank 0's frame.  Returns after the
        collective has been enqueued / completed; rank 0 then reads canvas.as_bytes_slice()."""
        if self.y1 > self.y0:
            self.scene.render(self.canvas)
        if self.mode == "nccl":
            self.canvas.sync()  # NCCL runs on torch's stream; the stripe must be complete first
            gather_stripes(self.dist, self.frame, self.bounds, self.height, self.width)
        else:
            self.canvas.sync()  # peer stores are complete when the kernel is
            self.dist.barrier()


def bench_sort_first(scene, cfg, dist, steps=50, warmup=5):
    """Times sort-first rendering of one frame of `cfg` per step in both gather modes."""
    import time

    import torch
    device = torch.device("cuda", torch.cuda.current_device())
    W, H = cfg["W"], cfg["H"]
    out = {"workload": cfg["label"], "stripes": stripe_bounds(H, dist.get_world_size())}
    for mode in ("nccl", "p2p"):
        sf, err = None, ""
        try:
            sf = SortFirst(scene, W, H, dist, device, mode=mode)
        except Exception as e:  # e.g. peer access not available
            err = str(e)[:200]
        # the ranks decide together: a mode is timed only if every rank could set it up (a rank that skipped on
        # its own would leave the others waiting in the gather / barrier)
        ok = torch.tensor([1 if sf is not None else 0], device=device, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if sf is not None:
                sf.close()
            out[mode] = {"error": err or "another rank could not set this mode up"}
            continue
        for _ in range(max(warmup, 3)):
            sf.render()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            sf.render()
        dist.barrier()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ms = float(dt.item()) / steps * 1e3
        out[mode] = {"ms_per_frame": ms, "frames_per_s": 1e3 / ms, "mtri_per_s": cfg["triangles"] / ms / 1e3,
                     "gather_bytes_into_root": 4 * W * (H - (sf.bounds[0][1] - sf.bounds[0][0]))}
        sf.close()
        del sf
    out["note"] = ("one frame split into tile-row stripes; vertex/setup replicated on every rank; host clock "
                   "around `steps` frames with barriers, max over ranks; includes the per-frame stream sync")
    return out
