"""Worker of tests/test_gpu_paths.py::test_forced_overflow_rerenders_from_the_frame_s_own_inputs.

Runs with DRAW_B200_REC_CAP / DRAW_B200_REFS_CAP set to tiny values (read when the library is loaded), so that
every first frame overflows its work buffers and is re-rendered internally with grown buffers.  The re-render must
use the inputs the frame was rendered with, whatever the scene and the canvas have become since."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import draw_b200  # noqa: E402
from conftest import GOLDEN, load_scene  # noqa: E402
from parity_util import assert_frames_equal, render_oracle  # noqa: E402

assert os.environ.get("DRAW_B200_REC_CAP") and os.environ.get("DRAW_B200_REFS_CAP")
path = np.load(GOLDEN + "/c4_camera_path.npy")
objs = load_scene("c4_dungeon")
W, H = 960, 544


def canvas():
    c = draw_b200.Canvas(W, H)
    c.init_depth(100000.0)
    return c


def scene():
    s = draw_b200.Scene(W, H)
    for o in objs:
        s.add_obj(o)
    return s


# (1) a lone overflowing frame
s, c = scene(), canvas()
s.camera = draw_b200.Camera.new(path[60, :3], path[60, 3:])
s.render(c)
got = (c.as_bytes_slice(), c.depth())
st = c.last_frame_stats()
assert st["overflow"] != 0, f"the tiny capacities did not overflow: {st}"
assert_frames_equal(got, render_oracle(objs, W, H, cam=path[60]), "overflowed frame")

# (2) pipelined: the camera moves on and another canvas is rendered before the first one is read
s, c0, c1 = scene(), canvas(), canvas()
s.camera = draw_b200.Camera.new(path[60, :3], path[60, 3:])
s.render(c0)
s.camera = draw_b200.Camera.new(path[20, :3], path[20, 3:])
s.render(c1)
got0 = (c0.as_bytes_slice(), c0.depth())
assert c0.last_frame_stats()["overflow"] != 0
got1 = (c1.as_bytes_slice(), c1.depth())
assert_frames_equal(got0, render_oracle(objs, W, H, cam=path[60]), "overflowed frame, camera moved since")
assert_frames_equal(got1, render_oracle(objs, W, H, cam=path[20]), "second canvas")

# (3) the stripe changes between render and read: the overflowed frame is re-rendered with its own stripe
s, c = scene(), canvas()
s.camera = draw_b200.Camera.new(path[60, :3], path[60, 3:])
c.set_stripe(0, 256)
s.render(c)
c.set_stripe(256, H)
got = c.as_bytes_slice()
want = render_oracle(objs, W, H, cam=path[60])[0]
assert c.last_frame_stats()["overflow"] != 0
assert np.array_equal(got[H - 256:], want[H - 256:]), "rows of the first stripe (colour rows are y-flipped)"
print("overflow worker ok")
