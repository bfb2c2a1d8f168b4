"""include/draw_b200.hpp — the C++ mirror of the reference's Rust host API (Scene / Canvas / Object / Texture with the
reference's method names; errors as exceptions where the reference panics).  The CPU test compiles a program
against it and runs its host-only parts; the GPU test runs the reference's driver sequence through it and compares
the frame with the Python binding's."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "native", "hpp_host.cpp")
LIBDIR = os.path.join(ROOT, "draw_b200")


def _build(tmp_path):
    exe = str(tmp_path / "hpp_host")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                    "-L", LIBDIR, "-ldraw_b200", f"-Wl,-rpath,{LIBDIR}"], check=True, capture_output=True, text=True)
    return exe


def test_hpp_compiles_and_host_paths_work(tmp_path):
    import draw_b200._native as N
    N.lib()  # the library is built
    exe = _build(tmp_path)
    r = subprocess.run([exe, "cpu", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK cpu"), r.stdout + r.stderr


def _fnv(b):
    h = 1469598103934665603
    for x in np.frombuffer(b, np.uint8).tolist():
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.gpu
def test_hpp_renders_the_same_frame_as_the_python_binding(tmp_path):
    import draw_b200
    from draw_b200.model import IndexedMesh, Object, Texture
    exe = _build(tmp_path)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = r.stdout.split()
    F = np.float32

    def tri(z, kd, alpha):
        v = np.array([[-60, -50, z], [60, -50, z], [0, 60, z]], F)
        n = np.tile(np.array([[0, 0, 1]], F), (3, 1))
        return Object("tri", v, n, np.zeros((1, 3), F), [IndexedMesh("m", np.array([[0, 1, 2, 0, 0, 0, 0, 1, 2]], np.uint32), 0)],
                      [Texture(alpha=alpha, kd=np.array(kd, F))])
    s, c = draw_b200.Scene(320, 240), draw_b200.Canvas(320, 240)
    c.init_depth(100000.0)
    c.apply_offset(0, 0)
    s.add_obj(tri(0.0, (1.0, 1.0, 0.0), 1.0))
    s.add_obj(tri(20.0, (0.0, 0.0, 1.0), 0.5))
    s.camera = draw_b200.Camera.new([20.0, 5.0, 120.0], [-0.2, 0.0, -1.0])
    s.camera.move_up(3.0)
    s.render(c)
    frame, depth = c.as_bytes_slice(), c.depth()
    assert got[1] == f"{_fnv(frame.tobytes()):016x}" and got[3] == f"{_fnv(depth.tobytes()):016x}", r.stdout
