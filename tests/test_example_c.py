"""examples/render_obj.c: the reference's frame loop through the C ABI from plain C — compiles against include/draw_b200.h
with -Wall -Wextra, links libdraw_b200.so, fails loudly without a GPU, renders and exports with one."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from test_loader_cpu import SYNTH_MTL, SYNTH_OBJ


def _build(tmp_path):
    from draw_b200 import build
    lib_dir = os.path.dirname(build.build())
    exe = str(tmp_path / "render_obj")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "render_obj.c"), "-L", lib_dir, "-ldraw_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True)
    (tmp_path / "synth.obj").write_text(SYNTH_OBJ)
    (tmp_path / "synth.mtl").write_text(SYNTH_MTL)
    return exe


def test_example_compiles_links_and_has_no_cpu_fallback(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the gpu test runs it")
    r = subprocess.run([exe, str(tmp_path / "synth.obj"), str(tmp_path / "out.png")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr  # the OBJ loaded (host code), the scene could not be created


@pytest.mark.gpu
def test_example_renders_and_exports(tmp_path):
    import draw_b200
    exe = _build(tmp_path)
    out = tmp_path / "out.png"
    r = subprocess.run([exe, str(tmp_path / "synth.obj"), str(out), "640", "352"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "640 x 352, 6 triangles in" in r.stdout
    img = draw_b200.load_image(str(out))
    assert img.shape == (352, 640, 4)
    assert len(np.unique(img.reshape(-1, 4), axis=0)) > 2  # more than the clear colour
