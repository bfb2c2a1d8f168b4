"""Multi-GPU drivers: one process per GPU, torch.distributed for the plumbing.

The reference is single-process and single-threaded; it has no multi-device path.  Two
partitions of the render path exist here:

  frame-parallel   a sequence of frames is dealt round-robin to the ranks; every rank holds the
                   whole scene and renders whole frames.  No collective on the data path.
  sort-first       ONE frame is split by tile rows (rows of draw_tile_size() = 32 pixels): rank r renders the
                   rows ty with ty % world == r ("interleaved": the scene usually sits mid-screen, contiguous
                   stripes would give the middle ranks all of it) or a contiguous stripe of rows.  Every rank runs
                   the per-vertex / per-triangle stages for the whole scene (k_front: replicated, bit-identical),
                   bins and rasterises only its rows.  Rows are disjoint, so compositing needs no depth compare.
                   Two ways to land the rows in rank 0's framebuffer:
                     "nccl"  each rank renders into its own canvas, then ONE NCCL collective lands the rows on rank 0
                             (stripes: grouped send/recv of the contiguous stripes; interleaved rows: pack, gather,
                             unpack — RowGather), ordered on the canvas' stream: the canvas renders on a torch
                             stream, so the gather follows k_tile and the next frame follows the gather without the
                             host waiting for either;
                     "p2p"   rank 0 exports its colour buffer through CUDA IPC; the other ranks' tile kernels
                             store their pixels straight into it over NVLink (peer stores from k_tile), so the
                             gather is fused into the raster kernel.  Completion is a device-side flag per rank
                             in a small buffer rank 0 owns (draw_flag_signal / draw_flags_wait): no collective,
                             no host synchronisation in the loop.

The colour buffer is y-flipped (canvas.rs:955-956): canvas rows [y0, y1) live in frame rows
[H - y1, H - y0), still one contiguous byte range.
"""
import ctypes as C

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


def interleaved_rows(height, world, rank, tile_h=TILE_H):
    """Pixel-row blocks [(y0, y1)] of rank `rank` under the interleaved partition: tile rows ty % world == rank."""
    rows = (height + tile_h - 1) // tile_h
    return [(ty * tile_h, min((ty + 1) * tile_h, height)) for ty in range(rank, rows, world)]


def rank_blocks(height, world, rank, layout, tile_h=TILE_H):
    """Row blocks [(y0, y1)] a rank renders, for layout "interleaved" or "stripes"."""
    if layout == "interleaved":
        return interleaved_rows(height, world, rank, tile_h)
    y0, y1 = stripe_bounds(height, world, tile_h)[rank]
    return [(y0, y1)] if y1 > y0 else []


def stripe_byte_range(height, width, y0, y1):
    """Byte range of canvas rows [y0, y1) inside the BGRA8 frame (rows are y-flipped)."""
    return (height - y1) * width * 4, (height - y0) * width * 4


def frames_of_rank(n_frames, world, rank):
    """Frame-parallel assignment: frame k goes to rank k mod world."""
    return list(range(rank, n_frames, world))


def gather_blocks(dist, frame, blocks_of_rank, height, width, root=0):
    """Gather every rank's row blocks of `frame` (flat uint8 tensor of H*W*4 bytes, same size on every
    rank) into the root's `frame`: one grouped send/recv.  `blocks_of_rank(r)` lists rank r's (y0, y1)
    blocks.  Works with any backend (NCCL on GPU tensors, gloo on CPU); with NCCL the operations are
    ordered on the current CUDA stream and the host does not wait for the data."""
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = []
    if rank == root:
        for r in range(world):
            if r == root:
                continue
            for y0, y1 in blocks_of_rank(r):
                b0, b1 = stripe_byte_range(height, width, y0, y1)
                ops.append(dist.P2POp(dist.irecv, frame[b0:b1], r))
    else:
        for y0, y1 in blocks_of_rank(rank):
            b0, b1 = stripe_byte_range(height, width, y0, y1)
            ops.append(dist.P2POp(dist.isend, frame[b0:b1], root))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()  # NCCL: makes the current stream wait, not the host


def gather_stripes(dist, frame, bounds, height, width, root=0):
    """gather_blocks for contiguous stripes given as [(y0, y1)] per rank."""
    gather_blocks(dist, frame, lambda r: [bounds[r]] if bounds[r][1] > bounds[r][0] else [], height, width, root)


class RowGather:
    """Gather of many row blocks per rank (the interleaved partition: a rank owns every world-th tile row) as ONE
    collective: every rank packs its rows into a contiguous buffer (one index_select), a single gather lands the
    buffers on the root (NCCL: grouped send / recv over NVLink), and the root scatters the rows of each rank into its
    frame (one index_copy_ per rank).  A send / recv per row block instead would cost the host tens of
    microseconds per block.  Works with any backend (NCCL on GPU tensors, gloo on CPU); stream-ordered with NCCL."""

    def __init__(self, dist, height, width, blocks_of_rank, device, root=0):
        import torch
        self.dist, self.root = dist, root
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.height, self.width = height, width
        self.rows = []
        for r in range(self.world):  # frame rows of rank r (the colour buffer is y-flipped), in block order
            idx = [fr for (y0, y1) in blocks_of_rank(r) for fr in range(height - y1, height - y0)]
            self.rows.append(torch.tensor(idx, dtype=torch.long, device=device))
        self.max_rows = max(1, max(int(t.numel()) for t in self.rows))
        self.send = torch.zeros((self.max_rows, width), dtype=torch.int32, device=device)
        self.recv = torch.zeros((self.world, self.max_rows, width), dtype=torch.int32, device=device) if self.rank == root else None

    def gather(self, frame):
        """frame: flat uint8 tensor of H*W*4 bytes (same size on every rank); the root's frame receives every rank's rows."""
        import torch
        px = frame.view(torch.int32).view(self.height, self.width)  # one BGRA pixel = one 32-bit element
        mine = self.rows[self.rank]
        if self.rank != self.root and mine.numel():
            torch.index_select(px, 0, mine, out=self.send[:mine.numel()])
        self.dist.gather(self.send, [self.recv[r] for r in range(self.world)] if self.rank == self.root else None, dst=self.root)
        if self.rank == self.root:
            for r in range(self.world):
                if r != self.root and self.rows[r].numel():
                    px.index_copy_(0, self.rows[r], self.recv[r, :self.rows[r].numel()])


class _DevicePtr:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def canvas_color_tensor(canvas, device):
    import torch
    ptr, _ = canvas.device_ptrs()
    return torch.as_tensor(_DevicePtr(ptr, canvas.width * canvas.height * 4), device=device)


def _broadcast_handle(dist, device, make):
    """Rank 0 produces a 64-byte IPC handle with make(); every rank gets it (or raises together)."""
    import torch
    handle = torch.zeros(64, dtype=torch.uint8, device=device)
    error = None
    if dist.get_rank() == 0:
        try:
            handle.copy_(torch.tensor(list(make()), dtype=torch.uint8))
        except Exception as e:  # every rank must still reach the broadcast: an all-zero handle says "no"
            error = e
    dist.broadcast(handle, 0)
    if error is not None:
        raise error
    if not bool(handle.any().item()):
        raise RuntimeError("rank 0 could not export its buffer through CUDA IPC")
    return (C.c_uint8 * 64)(*handle.cpu().tolist())


FLAG_CONSUMED = 32  # word of the flag buffer in which rank 0 publishes the last frame it has consumed
FLAG_ERROR = 63     # set by a wait that timed out


class SortFirst:
    """One frame across all ranks.  mode: "nccl" or "p2p"; layout: "interleaved" or "stripes" (module docstring)."""

    def __init__(self, scene, width, height, dist, device, mode="nccl", layout="interleaved", depth_max=100000.0, stream=None):
        import torch
        import draw_b200
        from . import _native as N
        self.scene, self.dist, self.mode, self.layout = scene, dist, mode, layout
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.width, self.height = width, height
        th = draw_b200.tile_size()
        self.blocks = lambda r: rank_blocks(height, self.world, r, layout, th)
        self.mine = self.blocks(self.rank)
        self.canvas = draw_b200.Canvas(width, height)
        self.canvas.init_depth(depth_max)
        self.canvas.apply_offset(0, 0)
        # the canvas renders on a torch stream of its own: the NCCL operations are issued with that stream current
        # and are therefore ordered against the frame (a NULL stream would mean "the canvas' own stream")
        self.stream = stream if stream is not None else torch.cuda.Stream(device=device)
        self.canvas.set_stream(self.stream.cuda_stream)
        if self.mine:
            if layout == "interleaved":
                self.canvas.set_tile_rows(self.rank, self.world)
            else:
                self.canvas.set_stripe(*self.mine[0])
        self.frame = canvas_color_tensor(self.canvas, device)
        self._row_gather = RowGather(dist, height, width, self.blocks, device) if mode == "nccl" and layout == "interleaved" else None
        self.seq = 0
        self._peer = self._flags = self._own_flags = None
        if mode == "p2p":
            lib = N.lib()
            raw = _broadcast_handle(dist, device, lambda: self._export_canvas(lib, N))
            if self.rank == 0:
                own = C.c_void_p()
                N.check(lib.draw_device_alloc(256, C.byref(own)))
                self._own_flags = own
            flags_raw = _broadcast_handle(dist, device, lambda: self._export_ptr(lib, N, self._own_flags))
            if self.rank != 0:
                peer, flags = C.c_void_p(), C.c_void_p()
                N.check(lib.draw_ipc_open(raw, C.byref(peer)))
                N.check(lib.draw_ipc_open(flags_raw, C.byref(flags)))
                self._peer, self._flags = peer, flags
                _, own_depth = self.canvas.device_ptrs()
                self.canvas.bind_external(peer.value, own_depth)  # colour -> rank 0's framebuffer over NVLink
            else:
                self._flags = self._own_flags
            # rank 0 fills its framebuffer with the clear colour before every frame (a local 4 W H-byte fill); nobody then
            # stores the pixels of tiles without geometry — most of a frame — over NVLink or at all
            self.canvas.set_empty_tile_color(False)
            dist.barrier()

    def _export_canvas(self, lib, N):
        buf = (C.c_uint8 * 64)()
        N.check(lib.draw_canvas_ipc_export(self.canvas._h, buf))
        return buf

    @staticmethod
    def _export_ptr(lib, N, ptr):
        buf = (C.c_uint8 * 64)()
        N.check(lib.draw_ipc_export(ptr, buf))
        return buf

    def close(self):
        from . import _native as N
        self.canvas.sync()
        self.dist.barrier()  # nobody closes a mapping a peer may still be writing through
        if self._peer is not None:
            self.canvas.bind_external(None, None)
            N.lib().draw_ipc_close(self._peer)
            N.lib().draw_ipc_close(self._flags)
            self._peer = self._flags = None
        self.dist.barrier()
        if self._own_flags is not None:
            N.lib().draw_device_free(self._own_flags)
            self._own_flags = self._flags = None

    def _flag(self, word):
        return C.c_void_p(self._flags.value + 4 * word)

    def render(self):
        """Enqueue: render this rank's rows and land all rows in rank 0's frame.  Nothing here waits on
        the host; rank 0's canvas stream is ordered after the arrival of every rank's rows (read the
        frame with canvas.as_bytes_slice(), or consume it on that stream)."""
        from . import _native as N
        self.seq += 1
        if self.mode == "nccl":
            import torch
            if self.mine:
                self.scene.render(self.canvas)
            with torch.cuda.stream(self.stream):
                if self._row_gather is not None:
                    self._row_gather.gather(self.frame)  # pack, one collective, unpack
                else:
                    gather_blocks(self.dist, self.frame, self.blocks, self.height, self.width)  # one contiguous stripe per rank
            return
        lib = N.lib()
        if self.rank != 0:
            # rank 0 must be done with the previous frame before its framebuffer is overwritten (k_tile waits for
            # the canvas stream; k_front and k_raster of this frame do not)
            N.check(lib.draw_flags_wait(self._flag(FLAG_CONSUMED), 1, self.seq - 1, self._flag(FLAG_ERROR), self.canvas._h))
            if self.mine:
                self.scene.render(self.canvas)
            N.check(lib.draw_flag_signal(self._flag(self.rank), self.seq, self.canvas._h))
        else:
            # whatever rank 0 has enqueued on the canvas' stream so far (its use of the previous frame) comes first:
            # starting the next frame is what tells the peers that the previous one has been consumed
            self.canvas.clear()  # colour (and depth) of the whole canvas: the tiles nobody draws in keep it
            N.check(lib.draw_flag_signal(self._flag(FLAG_CONSUMED), self.seq - 1, self.canvas._h))
            if self.mine:
                self.scene.render(self.canvas)
            if self.world > 1:
                N.check(lib.draw_flags_wait(self._flag(1), self.world - 1, self.seq, self._flag(FLAG_ERROR), self.canvas._h))


def bench_sort_first(scene, cfg, dist, frames=120, warmup=8):
    """Sort-first rendering of `frames` frames of cfg's camera path, one frame at a time across all ranks, in
    both gather modes and both layouts.  Device-timed (CUDA events on the canvas' stream, max over ranks); the
    composed frame of the first cameras is compared on rank 0 with the single-GPU frame (`bit_exact`)."""
    import torch
    import draw_b200
    device = torch.device("cuda", torch.cuda.current_device())
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H = cfg["W"], cfg["H"]
    cams = cfg["cameras"]
    cam_values = [draw_b200.Camera.new(c[:3], c[3:]) for c in cams] if cams is not None else None
    n_cam = len(cam_values) if cam_values else 1

    def set_cam(k):
        if cam_values is not None:
            scene.camera = cam_values[k % n_cam]

    stream = torch.cuda.Stream(device=device)
    # single-GPU reference on every rank: the time of the same frames rendered whole, one at a time
    full = draw_b200.Canvas(W, H)
    full.init_depth(100000.0)
    full.apply_offset(0, 0)
    full.set_stream(stream.cuda_stream)
    for k in range(n_cam):  # settles work-buffer capacities over the path
        set_cam(k)
        scene.render(full)
        full.sync()
    check = sorted({0, n_cam // 3, (2 * n_cam) // 3}) if rank == 0 else []
    refs = {}
    for k in check:
        set_cam(k)
        scene.render(full)
        with torch.cuda.stream(stream):
            refs[k] = canvas_color_tensor(full, device).clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for k in range(frames):
        set_cam(k)
        scene.render(full)
    e1.record(stream)
    torch.cuda.synchronize()
    single_ms = e0.elapsed_time(e1) / frames
    # ... and one frame at a time (the device idle before each): the latency of a lone frame on one GPU
    lone = []
    for k in range(min(frames, 60)):
        set_cam(k)
        torch.cuda.synchronize()
        e0.record(stream)
        scene.render(full)
        e1.record(stream)
        torch.cuda.synchronize()
        lone.append(e0.elapsed_time(e1))
    lone_ms = sum(lone) / len(lone)
    del full
    out = {"workload": cfg["label"], "frames": frames, "single_gpu_ms_per_frame": single_ms, "single_gpu_lone_frame_ms": lone_ms,
           "gather_bytes_into_root": 4 * W * H * (world - 1) // world}
    for layout in ("interleaved", "stripes"):
        for mode in ("p2p", "nccl"):
            key = f"{mode}_{layout}"
            sf, err = None, ""
            try:
                sf = SortFirst(scene, W, H, dist, device, mode=mode, layout=layout, stream=stream)
            except Exception as e:  # e.g. peer access not available
                err = str(e)[:200]
            # the ranks decide together: a mode is timed only if every rank could set it up (a rank that skipped on
            # its own would leave the others waiting in the gather)
            ok = torch.tensor([1 if sf is not None else 0], device=device, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if sf is not None:
                    sf.close()
                out[key] = {"error": err or "another rank could not set this mode up"}
                continue
            # parity first: composed frame == single-GPU frame, bit for bit (every rank renders, rank 0 compares)
            bit_exact = True
            for k in sorted({0, n_cam // 3, (2 * n_cam) // 3}):
                set_cam(k)
                sf.render()
                if rank == 0:
                    with torch.cuda.stream(stream):
                        bit_exact = bit_exact and bool(torch.equal(sf.frame, refs[k]))
            for k in range(warmup):
                set_cam(k)
                sf.render()
            torch.cuda.synchronize()
            dist.barrier()
            e0.record(stream)
            for k in range(frames):
                set_cam(k)
                sf.render()
            e1.record(stream)
            torch.cuda.synchronize()
            dt = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            ms = float(dt.item()) / frames
            be = torch.tensor([1 if bit_exact else 0], device=device, dtype=torch.int32)
            dist.broadcast(be, 0)
            out[key] = {"ms_per_frame": ms, "frames_per_s": 1e3 / ms, "mtri_per_s": cfg["triangles"] / ms / 1e3,
                        "speedup_vs_single_gpu": single_ms / ms, "speedup_vs_lone_frame": lone_ms / ms,
                        "bit_exact": bool(int(be.item()))}
            sf.close()
            del sf
    out["note"] = ("one frame at a time split by tile rows across the ranks; per-vertex / per-triangle stages replicated on "
                   "every rank; CUDA events on the canvas stream around all frames, max over ranks, no host synchronisation "
                   "inside the loop; single_gpu_ms_per_frame = the same frames rendered whole on one GPU the same way (frames back to "
                   "back: the front kernel of frame k+1 overlaps the tile kernel of frame k); single_gpu_lone_frame_ms = one frame "
                   "at a time on an idle GPU (mean over the first frames of the path); "
                   "bit_exact = composed frame of three cameras of the path == single-GPU frame")
    return out
