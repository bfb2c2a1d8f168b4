"""The pinned host mirror (Canvas::as_bytes_slice, canvas.rs:966-982: the reference's frame lives in host memory) stays
byte-identical to the frame whichever way it is refreshed — whole-frame copies, or the incremental tile copies of k_mirror
when most of the frame is clear colour — across moving cameras, several canvases in flight, and everything else that
writes the frame between renders (clear, the 2-D overlay, resize, partitions).  Run on the B200 box: pytest -m gpu"""
import numpy as np
import pytest

from conftest import GOLDEN, load_scene
from parity_util import DEPTH_MAX, render_oracle

pytestmark = pytest.mark.gpu


def _scene(name, W, H):
    import draw_b200
    s = draw_b200.Scene(W, H)
    objs = load_scene(name)
    for o in objs:
        s.add_obj(o)
    return s, objs


def test_incremental_mirror_follows_a_moving_camera():
    """Sixteen frames of an orbit on three canvases in flight with the mirror enabled: every host frame equals the frame
    a mirror-less canvas reads back, and after the first frames only a fraction of the tiles crosses the bus."""
    import draw_b200
    W, H = 1280, 720
    s, _ = _scene("c3_trio", W, H)
    cams = np.load(GOLDEN + "/orbit_camera_path.npy")
    ring = []
    for _ in range(3):
        c = draw_b200.Canvas(W, H)
        c.init_depth(DEPTH_MAX)
        c.enable_host_mirror(True)
        ring.append(c)
    plain = draw_b200.Canvas(W, H)
    plain.init_depth(DEPTH_MAX)
    whole_kb = (4 * W * H + 1023) // 1024
    copied = []
    for k in range(16):
        cam = cams[(4 * k) % len(cams)]
        s.camera = draw_b200.Camera.new(cam[:3], cam[3:])
        c = ring[k % 3]
        s.render(c)
        if k % 3 == 2 or k == 15:  # look at the canvases a few frames late, like a pipelined reader
            for j in range(max(0, k - 2), k + 1):
                camj = cams[(4 * j) % len(cams)]
                s.camera = draw_b200.Camera.new(camj[:3], camj[3:])
                s.render(plain)
                got = ring[j % 3].as_bytes_slice()
                assert np.array_equal(got, plain.as_bytes_slice()), f"frame {j}"
                copied.append(ring[j % 3].last_frame_stats()["mirror_kbytes"])
    assert copied[0] == whole_kb                       # the first refresh of a mirror is a whole-frame copy
    assert 0 < min(copied) < whole_kb // 4, copied     # later ones are incremental
    assert (np.array(copied) <= whole_kb).all()


def test_mirror_survives_everything_else_that_writes_the_frame():
    """clear, the GUI overlay, a partitioned render and a resize between mirrored renders: the host frame is always the
    device frame (each of them sends the next refresh down the whole-frame path)."""
    import draw_b200
    from draw_b200 import synthetic
    from oracle import pyoracle
    W, H = 640, 352
    s, objs = _scene("c2_donut", W, H)
    o = pyoracle.Scene(W, H)
    for ob in objs:
        o.add_obj(ob)
    c, oc = draw_b200.Canvas(W, H), pyoracle.Canvas(W, H)
    c.init_depth(DEPTH_MAX)
    oc.init_depth(DEPTH_MAX)
    c.enable_host_mirror(True)
    atlas = synthetic.font_atlas(64, 32)
    cmds = synthetic.gui_command_list(W, H, n_commands=2, quads_per_command=20, seed=8)

    def both_render():
        s.render(c)
        o.render(oc)

    def check(what):
        assert np.array_equal(c.as_bytes_slice(), oc.as_bytes()), what

    for k in range(4):  # warm the incremental path
        both_render()
        check(f"render {k}")
    assert c.last_frame_stats()["mirror_kbytes"] < 4 * W * H // 1024 // 2
    c.disable_depth_update()
    oc.disable_depth_update()
    for clip, v in cmds:  # overlay on top of the rendered frame
        c.draw_triangles(v, atlas, clip)
        oc.draw_triangles(v, atlas, clip)
    check("overlay")
    both_render()
    check("render after overlay")   # the overlay's pixels are gone from the mirror too
    both_render()
    check("second render after overlay")
    c.clear()
    oc.clear()
    check("clear")
    both_render()
    check("render after clear")
    # a partitioned render leaves the other rows alone
    c.set_stripe(0, 160)
    s.render(c)
    c.set_stripe(0, H)
    full = c.as_bytes_slice()
    assert np.array_equal(full, oc.as_bytes())  # same scene, same camera: the stripe re-rendered identical rows
    both_render()
    check("render after a stripe")
    c.resize(400, 300)
    oc.resize(400, 300)
    s2, o2 = draw_b200.Scene(400, 300), pyoracle.Scene(400, 300)
    for ob in objs:
        s2.add_obj(ob)
        o2.add_obj(ob)
    for k in range(3):
        s2.render(c)
        o2.render(oc)
        assert np.array_equal(c.as_bytes_slice(), oc.as_bytes()), f"after resize {k}"


def test_mirror_paths_agree(monkeypatch):
    """DRAW_B200_MIRROR_TILES=0 (whole-frame copies only) gives the same host frames as the default."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, sys; sys.path.insert(0, 'tests'); import draw_b200\n"
        "from conftest import load_scene, GOLDEN\n"
        "s = draw_b200.Scene(960, 544); [s.add_obj(o) for o in load_scene('c3_trio')]\n"
        "c = draw_b200.Canvas(960, 544); c.init_depth(100000.0); c.enable_host_mirror(True)\n"
        "cams = np.load(GOLDEN + '/orbit_camera_path.npy'); h = 0\n"
        "for k in range(10):\n"
        "    s.camera = draw_b200.Camera.new(cams[5 * k][:3], cams[5 * k][3:]); s.render(c)\n"
        "    a = c.as_bytes_slice(); h = (h * 1000003 + int(a.astype(np.uint64).sum()) + int(a[::7, ::5].astype(np.uint64).sum())) % (1 << 61)\n"
        "print('HASH', h, c.last_frame_stats()['mirror_kbytes'])\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for mode in ("0", "1"):
        env = dict(os.environ, DRAW_B200_MIRROR_TILES=mode)
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("HASH")][0].split()
        out[mode] = (line[1], int(line[2]))
    assert out["0"][0] == out["1"][0]
    assert out["1"][1] < out["0"][1]  # fewer bytes crossed the bus
