"""Shared helpers of the GPU parity tests: render the same scene through libdraw_b200.so and
through the CPU oracle, and compare colour bytes and depth bits exactly."""
import numpy as np

DEPTH_MAX = 100000.0


def render_gpu(objs, W, H, cam=None, offset=(0, 0), scene_wh=None, frames=1, light=None, stripe=None,
               return_handles=False):
    import draw_b200
    sw, sh = scene_wh or (W, H)
    scene, canvas = draw_b200.Scene(sw, sh), draw_b200.Canvas(W, H)
    canvas.init_depth(DEPTH_MAX)
    canvas.apply_offset(*offset)
    for o in objs:
        scene.add_obj(o)
    if light is not None:
        scene.set_light(light)
    if stripe is not None:
        canvas.set_stripe(*stripe)
    cams = cam if isinstance(cam, list) else [cam] * frames
    for cm in cams:
        if cm is not None:
            scene.camera = draw_b200.Camera.new(cm[:3], cm[3:])
        scene.render(canvas)
    out = (canvas.as_bytes_slice(), canvas.depth())
    return (out, scene, canvas) if return_handles else out


def render_oracle(objs, W, H, cam=None, offset=(0, 0), scene_wh=None, frames=1, light=None, stats=False):
    from oracle import pyoracle
    sw, sh = scene_wh or (W, H)
    scene, canvas = pyoracle.Scene(sw, sh), pyoracle.Canvas(W, H)
    canvas.init_depth(DEPTH_MAX)
    canvas.apply_offset(*offset)
    for o in objs:
        scene.add_obj(o)
    if light is not None:
        scene.set_light(light)
    cams = cam if isinstance(cam, list) else [cam] * frames
    for cm in cams:
        if cm is not None:
            scene.set_camera(cm[:3], cm[3:])
        scene.render(canvas, stats=stats)
    out = (canvas.as_bytes(), canvas.depth())
    return (out, scene.stats()) if stats else out


def assert_frames_equal(got, want, what=""):
    """Bit-exact: colour bytes (B,G,R,pad) and depth bits.  The north-star bar is looser
    (coverage/depth exact, RGB within 1 LSB on 99.9% of pixels); we hold the stricter one and
    report the looser figures when it fails."""
    gc, gd = got
    wc, wd = want
    assert gc.shape == wc.shape and gd.shape == wd.shape, f"{what}: shape mismatch {gc.shape} vs {wc.shape}"
    dbits = gd.view(np.uint32) != wd.view(np.uint32)
    cdiff = (gc != wc).any(axis=-1)
    if dbits.any() or cdiff.any():
        H = gd.shape[0]
        cov_g, cov_w = gd < DEPTH_MAX, wd < DEPTH_MAX
        ys, xs = np.where(dbits)
        cy, cx = np.where(cdiff)
        lsb = np.abs(gc[..., :3].astype(np.int32) - wc[..., :3].astype(np.int32)).max(axis=-1)
        msg = (f"{what}: depth bits differ on {int(dbits.sum())} px, coverage differs on "
               f"{int((cov_g != cov_w).sum())} px, colour differs on {int(cdiff.sum())} px "
               f"(>1 LSB on {int((lsb > 1).sum())} px of {lsb.size}); "
               f"first depth diffs (x,y,got,want): "
               f"{[(int(x), int(y), float(gd[y, x]), float(wd[y, x])) for y, x in list(zip(ys, xs))[:5]]}; "
               f"first colour diffs (x,row,got,want): "
               f"{[(int(x), int(y), gc[y, x].tolist(), wc[y, x].tolist()) for y, x in list(zip(cy, cx))[:5]]}")
        raise AssertionError(msg)
