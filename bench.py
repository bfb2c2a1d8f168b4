#!/usr/bin/env python
"""bench.py — frames/s and Mtri/s of the draw hot path at 3840x2160 Phong+texture on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c3|c2|c4|c5]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

Workload at N = 1: BASELINE.json configs[2], "C3" (soldier1 + skeleton + lemur, 20 697 triangles, 4K, Phong +
texture).  One "step" is one pass of the hot path over one BATCH of frames: `config.frames_per_step` calls of
Scene::render, each with the next camera of a committed camera path (so that K = 20 steps span > 100 ms and
dozens of clock samples; a single 4K frame of this scene is ~40 us of GPU time).  The metric stays frames/s.
Rank 0 prints ONE JSON line.  Keys beyond the base contract:

  value / ms_per_step   scene resident in HBM, K steps back to back, CUDA events on the timing stream (joined on
                        the device to every canvas' stream), max over ranks.  Frames rotate over a ring of
                        canvases larger than L2 so no frame finds its framebuffer in L2.
  lone_frame            latency of ONE frame on an idle GPU with L2 flushed (256 MB fill) before it.
  e2e                   the same metric through the public API with host buffers: every frame sets the camera
                        (host), renders, and is read back into pinned host memory (Canvas::as_bytes_slice); host
                        clock.  `value` keeps six canvases in flight, `serial_value` is the reference's own loop
                        (src/app/mod.rs:196-202: render, read, repeat — one canvas); `d2h_bytes_per_step` counts the
                        bytes that crossed PCIe (the library refreshes its host mirror whole or, for mostly-empty
                        frames, by the tiles that changed); `d2h_ceiling_gbs` is a plain cudaMemcpyAsync of whole
                        frames, all ranks at once.
  roofline              k_tile (the dominant kernel): algorithmic bytes per launch / its mean device time from
                        CUDA events recorded around it on the launching stream, against MEASURED_PEAKS.json.
  cpu_baseline          the CPU oracle (C++ restatement of the reference renderer, 1 thread) on a bounded number
                        of frames of the same workload, rank 0, N = 1 only.
  sort_first            (N > 1) ONE frame partitioned across the ranks (transform replicated, each rank rasterises
                        its tile rows, stripes land in rank 0's framebuffer over NVLink) on BASELINE.json
                        configs[3] (C4) and, at N = 8, configs[4] (C5); every composed frame is compared with the
                        single-GPU frame (`bit_exact`).

--impl reference times the reference's CPU implementation of the path.  The reference is a Rust program; no
Rust toolchain exists in this image, so the arm runs the C++ oracle port (oracle/oracle.cpp), single-threaded
like the reference (it has no threads anywhere in src/); each of its steps is a bounded sample of the step
above (`cpu_baseline.sample`).  Both arms print the same `config`.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
DEPTH_MAX = 100000.0

# frames_per_step: a step is ~5 ms of GPU work
CONFIGS = {
    "c2": dict(scene="c2_donut", W=1920, H=1080, label="C2 donut 1920x1080 Phong only", path="orbit_camera_path.npy",
               frames_per_step=256),
    "c3": dict(scene="c3_trio", W=3840, H=2160, label="C3 soldier1+skeleton+lemur 3840x2160 Phong+texture",
               path="orbit_camera_path.npy", frames_per_step=128),
    "c4": dict(scene="c4_dungeon", W=3840, H=2160, label="C4 dungeon_set 120-frame fly-through 3840x2160",
               path="c4_camera_path.npy", frames_per_step=16),
    "c5": dict(scene=None, W=7680, H=4320, label="C5 synthetic 10M-triangle torus 7680x4320 Phong+checker texture",
               path=None, frames_per_step=4),
}


def load_workload(name):
    from draw_b200 import scene_cache, synthetic
    cfg = dict(CONFIGS[name])
    cfg["name"] = name
    if name == "c5":
        n_theta, n_phi = int(os.environ.get("DRAW_C5_NTHETA", 2500)), int(os.environ.get("DRAW_C5_NPHI", 2000))
        cfg["objects"] = [synthetic.torus(n_theta, n_phi, texture=synthetic.checker_material())]
    else:
        cfg["objects"] = scene_cache.load(os.path.join(GOLDEN, "scenes", cfg["scene"] + ".npz"))
    cfg["cameras"] = np.load(os.path.join(GOLDEN, cfg["path"])) if cfg.get("path") else None
    objs = cfg["objects"]
    cfg["triangles"] = int(sum(o.triangle_count() for o in objs))
    n_pos = sum(o.vertices.shape[0] for o in objs)
    n_nrm = sum(o.normals_vertices.shape[0] for o in objs)
    n_uv = sum(o.texture_vertices.shape[0] for o in objs)
    tex_bytes = 0
    for o in objs:
        used = {m.texture_idx for m in o.meshes}
        seen = set()
        for i in used:
            for img in (o.textures[i].map_ka, o.textures[i].map_kd):
                if img is None:
                    tex_bytes += 3 if "default" not in seen else 0
                    seen.add("default")
                elif id(img) not in seen:
                    seen.add(id(img))
                    tex_bytes += int(np.asarray(img).size)
    # ALGO_BYTES(frame) = 8 W H + 36 T + 12 (Np + Nn + Nuv) + bound texture bytes
    cfg["tex_bytes"] = tex_bytes
    cfg["algo_bytes_frame"] = 8 * cfg["W"] * cfg["H"] + 36 * cfg["triangles"] + 12 * (n_pos + n_nrm + n_uv) + tex_bytes
    cfg["algo_bytes_tile_kernel"] = 8 * cfg["W"] * cfg["H"] + tex_bytes
    return cfg


def ring_size(cfg):
    return max(2, int(np.ceil(160e6 / (8 * cfg["W"] * cfg["H"]))) + 1)


UNIFORM_BYTES = 296  # sizeof(FrameUniforms), draw_b200/csrc/device_types.h: the per-frame host-to-device copy (camera, light, canvas state)


def workload_config(cfg, world):
    """The `config` object of the JSON line — the same in both arms (--impl ours / reference)."""
    W, H = cfg["W"], cfg["H"]
    n_ring = ring_size(cfg)
    cams = cfg["cameras"]
    return {"workload": cfg["label"], "triangles": cfg["triangles"], "width": W, "height": H,
            "frames_per_step": cfg["frames_per_step"],
            "camera": (f"{len(cams)}-camera path tests/golden/{cfg['path']}, next camera every frame" if cams is not None
                       else "default camera (scene/mod.rs:763-765)"),
            "parallelism": "single GPU" if world == 1 else f"frame-parallel x{world} (no collective)",
            "l2": f"ring of {n_ring} canvases x {8 * W * H / 1e6:.0f} MB (colour+depth) = "
                  f"{n_ring * 8 * W * H / 1e6:.0f} MB > 126 MB L2, rotated every frame"}


def fill_roofline(config, frame_seconds, sm_mhz):
    """SURVEY.md §8(d), secondary bound (binds the overdraw-heavy C4): ALGO_FLOP(frame) = 23 F_cov + 101 P_vis against
    the non-FMA FP32 peak, SMs x 128 lanes x clock.  F_cov / P_vis are counted by the CPU oracle offline
    (tests/golden/make_fill_counts.py -> tests/golden/fill_counts.json); nothing is counted inside the timed region."""
    try:
        fc = json.load(open(os.path.join(GOLDEN, "fill_counts.json")))[config]
    except Exception:
        return None
    flop = 23.0 * fc["f_cov_mean"] + 101.0 * fc["p_vis_mean"]
    peak = 148 * 128 * (sm_mhz or 1965.0) * 1e6 / 1e12
    achieved = flop / frame_seconds / 1e12
    return {"bound": "fp32 without FMA (parity forbids contraction)", "algo_flop_per_frame": flop, "f_cov": fc["f_cov_mean"],
            "p_vis": fc["p_vis_mean"], "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": "148 SMs x 128 lanes x SM clock under load"}


def ncu_traffic(kernel, config):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(kernel, {}).get(config)
    except Exception:
        return None


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md).  NVML is polled
    from a thread every ~2 ms; falls back to one nvidia-smi query if pynvml is unavailable."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
               "hw_power_brake_slowdown": 0x80}

    def __init__(self, gpu_index):
        self.gpu, self.samples, self.reasons, self._stop, self._thread, self.max_mhz = gpu_index, [], set(), False, None, None

    def _run(self, nv, h):
        while not self._stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading

            import pynvml as nv
            nv.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES if it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",") if x.strip().isdigit()]
            h = nv.nvmlDeviceGetHandleByIndex(ids[self.gpu] if self.gpu < len(ids) else self.gpu)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self._thread = threading.Thread(target=self._run, args=(nv, h), daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None

    def stop(self):
        if self._thread is not None:
            self._stop = True
            self._thread.join(timeout=2)
            sm = self.samples
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                    "sm_max_mhz": self.max_mhz, "samples": len(sm), "reasons": sorted(self.reasons), "source": "nvml, 2 ms poll"}
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                  "-i", str(self.gpu)], capture_output=True, text=True, timeout=10).stdout
            a, b2 = [float(x) for x in out.strip().split(",")]
            return {"sm_mhz": a, "sm_max_mhz": b2, "samples": 1, "reasons": [], "source": "nvidia-smi after the run"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock query unavailable"]}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle as cpu_baseline and as the reference arm
# ------------------------------------------------------------------------------------------------
def oracle_scene(cfg):
    from oracle import pyoracle
    s, c = pyoracle.Scene(cfg["W"], cfg["H"]), pyoracle.Canvas(cfg["W"], cfg["H"])
    c.init_depth(DEPTH_MAX)
    c.apply_offset(0, 0)
    for o in cfg["objects"]:
        s.add_obj(o)
    return s, c


def oracle_frame(s, c, cfg, k):
    if cfg["cameras"] is not None:
        cam = cfg["cameras"][k % len(cfg["cameras"])]
        s.set_camera(cam[:3], cam[3:])
    s.render(c)


def cpu_baseline(cfg, budget_s=12.0, max_frames=400):
    s, c = oracle_scene(cfg)
    oracle_frame(s, c, cfg, 0)  # warm-up (page faults, caches)
    t0, n = time.perf_counter(), 0
    while n < max_frames and (n < 3 or time.perf_counter() - t0 < budget_s):
        oracle_frame(s, c, cfg, n)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "port",
            "mtri_per_s": cfg["triangles"] * n / dt / 1e6,
            "sample": f"the first {n} frames of the {cfg['label']} workload's camera path in {dt:.1f} s, C++ oracle port "
                      f"(oracle/oracle.cpp, -O3, -ffp-contract=off), 1 thread of {os.cpu_count()} host cores; "
                      "the reference renderer is single-threaded"}


def run_reference(args, cfg):
    rank, _, world = dist_env()
    if rank != 0:
        return
    s, c = oracle_scene(cfg)
    # A step of this arm is a bounded sample of the GPU arm's step (frames_per_step frames): as many of its frames
    # as fit in ~2 s of CPU time, at least one; the whole run stays within a few minutes.
    oracle_frame(s, c, cfg, 0)
    t0 = time.perf_counter()
    oracle_frame(s, c, cfg, 0)
    per = max(time.perf_counter() - t0, 1e-6)
    steps = max(1, args.steps)
    budget = 120.0 / (steps + args.warmup)  # seconds per step
    fps_step = int(max(1, min(cfg["frames_per_step"], budget / per)))
    k = 0
    for _ in range(args.warmup):
        for _ in range(fps_step):
            oracle_frame(s, c, cfg, k)
            k += 1
    k = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        for _ in range(fps_step):
            oracle_frame(s, c, cfg, k)
            k += 1
    dt = time.perf_counter() - t0
    fps = steps * fps_step / dt
    line = {
        "impl": "reference", "metric": "frames/s at 3840x2160 (Phong+texture)", "value": fps, "unit": "frames/s",
        "mtri_per_s": cfg["triangles"] * fps / 1e6, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "reference model assets (committed scene cache), committed camera path; no published baseline",
        "config": workload_config(cfg, max(1, args.gpus)),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port",
                         "sample": f"each step = the next {fps_step} of the step's {cfg['frames_per_step']} frames of {cfg['label']} "
                                   f"({steps} steps, {steps * fps_step} frames, {dt:.1f} s); C++ oracle port of the "
                                   f"single-threaded Rust reference (no Rust toolchain in this image), 1 of "
                                   f"{os.cpu_count()} host cores"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def new_canvas(draw_b200, W, H):
    c = draw_b200.Canvas(W, H)
    c.init_depth(DEPTH_MAX)
    c.apply_offset(0, 0)
    return c


def run_ours(args, cfg):
    import torch
    import draw_b200

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the draw_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    draw_b200.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None
    # a real (non-default) stream for the timing events.  Every canvas of the ring renders on its own stream
    # (frames on different canvases are independent and overlap, as they would for an application that
    # keeps several frames in flight); before the closing event is recorded the timing stream is made to
    # wait, on the device, for every canvas (Canvas.stream_wait), so ev0 -> ev1 spans all the frames.
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    W, H = cfg["W"], cfg["H"]
    FPS = cfg["frames_per_step"]

    scene = draw_b200.Scene(W, H)
    for o in cfg["objects"]:
        scene.add_obj(o)
    # ring of canvases: working set of framebuffers larger than the 126 MB L2
    n_ring = ring_size(cfg)
    ring = []
    for _ in range(n_ring):
        c = new_canvas(draw_b200, W, H)
        if os.environ.get("DRAW_BENCH_SHARED_STREAM"):  # A/B: all canvases enqueue on the timing stream
            c.set_stream(stream.cuda_stream)
        ring.append(c)
    cams = cfg["cameras"]
    cam_values = [draw_b200.Camera.new(cam[:3], cam[3:]) for cam in cams] if cams is not None else None

    def frame(k, canvas):
        if cam_values is not None:
            scene.camera = cam_values[k % len(cam_values)]  # scene.camera = Camera::new(pos_k, dir_k, ratio)
        scene.render(canvas)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # every frame-in-flight slot of the scene is set up before anything is timed (streams, work buffers, key pages,
    # frame graphs), then work-buffer capacities are settled over the whole camera path, then every (slot, canvas)
    # pairing is run once
    scene.prepare(ring[0])
    for k in range(len(cams) if cams is not None else 1):
        frame(k, ring[0])
        ring[0].sync()
    for k in range(8 * n_ring):
        frame(k, ring[k % n_ring])
    for k in range(args.warmup * FPS):
        frame(k, ring[k % n_ring])
    barrier()

    # ---- timed region: K steps of FPS frames, back to back ---------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = scene.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for k in range(args.steps * FPS):
        frame(k, ring[k % n_ring])
    for c in ring:
        c.stream_wait(stream.cuda_stream)
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = scene.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    for c in ring:
        st = c.last_frame_stats()
        if st["overflow"]:
            raise SystemExit(f"bench.py: a timed frame overflowed a work buffer ({st}); timing invalid")
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    frames_rank = args.steps * FPS
    frames_total = frames_rank * world
    fps = frames_total / (ms_total * 1e-3)

    if args.quick:  # A/B runs: the timed region and the lone-frame latency only
        lone = []
        for k in range(24):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            frame(k * 5, ring[k % n_ring])
            ring[k % n_ring].stream_wait(stream.cuda_stream)
            b.record(stream)
            torch.cuda.synchronize()
            lone.append(a.elapsed_time(b))
        if rank == 0:
            print(json.dumps({"quick": True, "config": args.config, "value": fps, "us_per_frame": 1e3 * ms_total / frames_rank,
                              "lone_frame_us_median": 1e3 * float(np.median(lone)), "lone_frame_us_max": 1e3 * float(np.max(lone)),
                              "gpu_launches": launches, "clocks": clocks}), flush=True)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-kernel times (separate pass, same workload, events around each kernel) -----------
    scene.set_kernel_timing(True)
    ktimes = {k: [] for k in scene.KERNELS}
    n_prof = min(frames_rank, 64)
    for k in range(n_prof):
        frame(k, ring[k % n_ring])
        for name, ms in scene.last_kernel_times(ring[k % n_ring]).items():
            ktimes[name].append(ms)
    scene.set_kernel_timing(False)
    kmean = {k: float(np.mean(v)) for k, v in ktimes.items()}
    not_launched = [k for k, v in kmean.items() if v < 0.004 and k in scene.OPTIONAL_KERNELS]  # an empty event pair

    # ---- lone frame: L2 flushed, idle device -----------------------------------------------------
    flush = torch.empty(int(256e6) // 4, dtype=torch.float32, device="cuda")
    per_frame = []
    for k in range(min(frames_rank, 32)):
        flush.fill_(float(k))
        torch.cuda.synchronize()  # the frame's geometry runs on the scene's own streams: start it on an idle GPU
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        frame(k, ring[0])
        ring[0].stream_wait(stream.cuda_stream)
        b.record(stream)
        torch.cuda.synchronize()
        per_frame.append(a.elapsed_time(b))
    del flush

    # ---- e2e: public API, host in / host out ----------------------------------------------------
    # (1) N_E2E canvases on their own streams with the host mirror enabled: every render is followed by the
    # refresh of its pinned host mirror, frame k renders while the frames before it are on the PCIe link.  Every
    # frame ends up whole in host memory and is looked at there; what crosses the link is counted below.
    N_E2E = max(1, int(os.environ.get("DRAW_BENCH_E2E_DEPTH", "6")))  # canvases in flight (measured on C3: 3 -> 8.3 k, 6 -> 11.1 k, 8 -> 11.9 k frames/s)
    pair = []
    for _ in range(N_E2E):
        c = new_canvas(draw_b200, W, H)
        c.enable_host_mirror(True)
        pair.append(c)
    for k in range(2 * N_E2E):
        frame(k, pair[k % N_E2E])
        pair[k % N_E2E].as_bytes_slice(copy=False)
    n_e2e = frames_rank
    # bytes that actually cross PCIe per frame: the library copies a frame to its pinned host mirror whole (4*W*H), or — when
    # most of it is clear colour and was clear colour in the frame the mirror holds — only the 64x8-pixel strips that differ
    # (draw_frame_stats.mirror_kbytes; the mirror is byte-identical to the device frame either way: tests/test_gpu_mirror.py)
    def mirrored_bytes(canvas):
        return min(4 * W * H, 1024 * canvas.last_frame_stats()["mirror_kbytes"])

    barrier()
    t0 = time.perf_counter()
    checksum = 0
    d2h_bytes = 0
    for k in range(n_e2e + N_E2E - 1):
        if k >= N_E2E - 1:  # the oldest frame in flight: wait for it and look at it on the host
            host = pair[(k - (N_E2E - 1)) % N_E2E].as_bytes_slice(copy=False)
            checksum ^= int(host[H // 2, W // 2, 0])
            d2h_bytes += mirrored_bytes(pair[(k - (N_E2E - 1)) % N_E2E])
        if k < n_e2e:
            if cams is None:  # the per-frame host input: the camera (scene.camera = Camera::new(...))
                scene.camera = draw_b200.Camera.new([0.0, 0.0, 150.0], [0.0, 0.0, -150.0])
            frame(k, pair[k % N_E2E])
    host = pair[(n_e2e - 1) % N_E2E].as_bytes_slice(copy=False)
    checksum ^= int(host[H // 2, W // 2, 0])
    d2h_bytes += mirrored_bytes(pair[(n_e2e - 1) % N_E2E])
    barrier()
    e2e_s = time.perf_counter() - t0
    # (2) the reference's own loop (src/app/mod.rs:196-202): one canvas, render, read the frame, repeat
    serial = pair[0]
    n_serial = min(n_e2e, 256)
    barrier()
    t0 = time.perf_counter()
    for k in range(n_serial):
        if cams is None:
            scene.camera = draw_b200.Camera.new([0.0, 0.0, 150.0], [0.0, 0.0, -150.0])
        frame(k, serial)
        host = serial.as_bytes_slice(copy=False)
        checksum ^= int(host[H // 2, W // 2, 0])
    barrier()
    serial_s = time.perf_counter() - t0
    del pair, serial
    # (3) what the link gives: the same number of bytes per frame as a plain device-to-pinned-host copy, every rank at once
    dev_buf = torch.empty(4 * W * H, dtype=torch.uint8, device="cuda")
    host_buf = torch.empty(4 * W * H, dtype=torch.uint8, pin_memory=True)
    for _ in range(3):
        host_buf.copy_(dev_buf, non_blocking=True)
    n_copy = 64
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_copy):
        host_buf.copy_(dev_buf, non_blocking=True)
    barrier()
    copy_s = time.perf_counter() - t0
    del dev_buf, host_buf
    if dist is not None:
        t = torch.tensor([e2e_s, serial_s, copy_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, serial_s, copy_s = (float(x) for x in t.tolist())
        t = torch.tensor([float(d2h_bytes)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        d2h_bytes = int(t.item())
    e2e_fps = n_e2e * world / e2e_s
    serial_fps = n_serial * world / serial_s
    d2h_ceiling = n_copy * world * 4 * W * H / copy_s / 1e9

    sort_first = None
    if dist is not None and not os.environ.get("DRAW_BENCH_NO_SORT_FIRST"):
        from draw_b200 import multi
        sort_first = {}
        names = ["c4"] + (["c5"] if world >= 8 or os.environ.get("DRAW_BENCH_SORT_FIRST_C5") else [])
        for name in names:
            sf_cfg = cfg if name == args.config else load_workload(name)
            if name == args.config:
                sf_scene = scene
            else:
                sf_scene = draw_b200.Scene(sf_cfg["W"], sf_cfg["H"])
                for o in sf_cfg["objects"]:
                    sf_scene.add_obj(o)
            sort_first[name] = multi.bench_sort_first(sf_scene, sf_cfg, dist, frames=int(os.environ.get("DRAW_BENCH_SF_FRAMES", 120)))
            if sf_scene is not scene:
                del sf_scene

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        # k_tile (the longest kernel of the frame: `roofline`) writes the whole frame once.  Algorithmic bytes of a launch =
        # 8 B per pixel (colour + depth) + the bound textures (read at most once).
        st = ring[0].last_frame_stats()
        tile_bytes = 8 * W * H + cfg["tex_bytes"]
        t_tile = kmean["k_tile"] * 1e-3
        achieved = tile_bytes / t_tile / 1e9
        s_per_frame = ms_total * 1e-3 / frames_rank
        frame_gbs = cfg["algo_bytes_frame"] / s_per_frame / 1e9
        lone_ms = float(np.median(per_frame))
        line = {
            "metric": "frames/s at 3840x2160 (Phong+texture)", "value": fps, "unit": "frames/s",
            "mtri_per_s": cfg["triangles"] * fps / 1e6,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "us_per_frame": 1e6 * s_per_frame,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "reference model assets (committed scene cache), committed camera path; no published baseline",
            "config": workload_config(cfg, world),
            "lone_frame": {"ms": lone_ms, "frames_per_s": 1e3 / lone_ms,
                           "frame_frac": cfg["algo_bytes_frame"] / (lone_ms * 1e-3) / 1e9 / peak,
                           "protocol": "one frame at a time: 256 MB fill (L2 flushed), device idle, CUDA events around the frame, median of 32"},
            "clocks": clocks,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": UNIFORM_BYTES * FPS,
                    "d2h_bytes_per_step": (d2h_bytes // max(1, n_e2e * world) + 128) * FPS,
                    "d2h_frame_bytes": 4 * W * H,
                    "serial_value": serial_fps, "d2h_ceiling_gbs": d2h_ceiling,
                    "d2h_achieved_gbs": d2h_bytes / e2e_s / 1e9,
                    "note": "per frame: camera set on host, Scene::render, Canvas::as_bytes_slice of the frame in pinned host memory, "
                            "one byte of it read on the host.  value: three canvases in flight with the host mirror enabled (the copy "
                            "of a frame follows its render on the canvas stream, frame k renders while frames k-1 / k-2 cross PCIe); "
                            "serial_value: the reference's loop (src/app/mod.rs:196-202), one canvas, render -> read -> next; "
                            "d2h_ceiling_gbs: plain cudaMemcpyAsync of 4*W*H bytes to pinned host memory, all ranks at once "
                            "(what the PCIe / host-memory path gives).  Geometry is uploaded once by add_obj like the reference's "
                            "Scene owns its objects"},
            "gpu_launches": launches,
            "kernel_ms": {k: v for k, v in kmean.items() if k not in not_launched},
            "kernel_ms_note": "CUDA events around each kernel, kernels of a frame run one after the other (no overlap "
                              "between frames); not launched in this configuration: " + (", ".join(not_launched) or "none"),
            "roofline": {"bound": "hbm", "kernel": "k_tile", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic("k_tile", args.config), "peak_source": peak_src,
                         "algo_bytes_per_launch": tile_bytes,
                         "note": "k_tile writes the whole frame once — 8 B/pixel: the rasterised tiles from shared memory, the "
                                 "empty tiles as streaming stores between its raster items — and reads the bound textures; "
                                 "its duration is event-timed with the frame's kernels run one after the other. `frame_*` = "
                                 "SURVEY.md 8(d) ALGO_BYTES(frame) / time per frame of the timed region (frames overlapped)",
                         "empty_tiles": st["empty_tiles"], "work_items": st["work_items"],
                         "refs": {k: st[k] for k in ("large_refs", "medium_refs", "small_refs", "transparent_refs")},
                         "k_front_phase_us": dict(zip(("vertex", "barrier1", "triangle", "barrier2_and_huge", "tile", "triangle_slowest_cta", "huge"),
                                                      (x / 1e3 for x in st["front_phase_ns"]))),
                         "frame_algo_bytes": cfg["algo_bytes_frame"], "frame_achieved": frame_gbs,
                         "frame_frac": frame_gbs / peak,
                         "fill": fill_roofline(args.config, s_per_frame, (clocks or {}).get("sm_mhz"))},
        }
        if sort_first is not None:
            line["sort_first"] = sort_first
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(cfg)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)  # C3: 40 x 128 frames, a ~200 ms timed region
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="A/B runs: time the K steps and print a short line (no e2e, no per-kernel pass)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = load_workload(args.config)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
