"""The library's C++ OBJ/MTL loader (draw_object_load_obj, object.rs:106-454) against the numpy
restatement in oracle/obj_loader.py, bit for bit, on the reference's own model files when the
reference tree is mounted, and on a synthetic OBJ that exercises quads, missing uvs/normals,
material switches and negative indices everywhere else."""
import os

import numpy as np
import pytest

import draw_b200
from draw_b200 import api
from oracle import obj_loader

REF_MODELS = "/root/reference/models"


def _same(a, b):
    assert a.name == b.name
    for x, y, what in ((a.vertices, b.vertices, "vertices"), (a.normals_vertices, b.normals_vertices, "normals"),
                       (a.texture_vertices, b.texture_vertices, "uvs")):
        assert x.shape == y.shape, what
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), what   # NaNs of unused generated normals included
    assert len(a.meshes) == len(b.meshes)
    for m, n in zip(a.meshes, b.meshes):
        assert m.name == n.name and m.texture_idx == n.texture_idx
        assert np.array_equal(np.asarray(m.triangles, np.uint32), np.asarray(n.triangles, np.uint32))
    assert len(a.textures) == len(b.textures)
    for t, u in zip(a.textures, b.textures):
        assert t.name == u.name
        for k in ("ka", "kd", "ks"):
            assert np.array_equal(np.asarray(getattr(t, k), np.float32), np.asarray(getattr(u, k), np.float32)), k
        assert np.float32(t.alpha) == np.float32(u.alpha)
        for k in ("map_ka", "map_kd"):
            x, y = getattr(t, k), getattr(u, k)
            assert (x is None) == (y is None)
            if x is not None:
                assert np.array_equal(x, y)


SYNTH_OBJ = """# synthetic
mtllib synth.mtl
o first
v 0 0 0
v 2 0 0
v 2 2 0
v 0 2 0
v 1 1 3
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 1
usemtl red
f 1/1/1 2/2/1 3/3/1 4/4/1
usemtl glass
f 1//1 2//1 5//1
f -4 -3 -1
g side
f 2/2 3/3 5/1
s 1
l 1 2
o second
usemtl nosuch
f 3 4 5
"""
SYNTH_MTL = """newmtl red
Ka 1 0 0
Kd 0.8 0.1 0.1
Ks 0.5 0.5 0.5
d 1.0
Ns 10
newmtl glass
Kd 0.2 0.2 0.9
d 0.5
illum 2
"""


def test_synthetic_obj(tmp_path):
    (tmp_path / "synth.obj").write_text(SYNTH_OBJ)
    (tmp_path / "synth.mtl").write_text(SYNTH_MTL)
    got = api.load_obj(str(tmp_path / "synth.obj"))
    want = obj_loader.load_from_file(str(tmp_path / "synth.obj"))
    _same(got, want)
    assert got.triangle_count() == 6            # quad split in two + 4 triangles
    assert [t.name for t in got.textures] == ["default", "red", "glass"]
    assert got.textures[2].alpha == 0.5 and np.allclose(got.textures[2].ka, 0.9)   # missing Ka -> default value
    assert np.isclose(np.linalg.norm(got.vertices, axis=1).max(), 100.0, rtol=1e-6)  # rescaled to radius 100


def test_missing_file_is_an_error_not_a_crash(tmp_path):
    with pytest.raises(draw_b200.DrawError, match="Unable to open"):
        api.load_obj(str(tmp_path / "nope.obj"))


@pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="reference tree not mounted")
@pytest.mark.parametrize("rel", ["donut/donut.obj", "lemur/lemur.obj", "soldier1/soldier1.obj", "skeleton/fgc_skeleton.obj",
                                 "dungeon_set/prop_floor_barrel.obj", "dungeon_set/struct_wall_big_door_base.obj"])
def test_reference_models(rel):
    path = os.path.join(REF_MODELS, rel)
    _same(api.load_obj(path), obj_loader.load_from_file(path))


def test_damaged_obj_and_mtl_text_is_an_error_or_an_object(tmp_path):
    """Tokens dropped, swapped for garbage or for out-of-range indices, lines cut: the C++ loader returns an Object or fails
    with an error (the reference panics on most of these, object.rs:106-454) — no crash."""
    rng = np.random.default_rng(6)
    junk = ["", "x", "-", "1e999", "nan", "99999999999", "-7", "0", "1/", "/1", "1//", "//", "a/b/c", "3/99/1", "3/1/99"]
    n_err = 0
    for trial in range(250):
        obj_lines, mtl_lines = SYNTH_OBJ.splitlines(), SYNTH_MTL.splitlines()
        for lines in (obj_lines, mtl_lines):
            for _ in range(int(rng.integers(1, 4))):
                k = int(rng.integers(0, len(lines)))
                tok = lines[k].split(" ")
                j = int(rng.integers(0, len(tok)))
                what = rng.random()
                if what < 0.5:
                    tok[j] = junk[int(rng.integers(0, len(junk)))]
                elif what < 0.75:
                    del tok[j]
                else:
                    tok = tok[:j]
                lines[k] = " ".join(tok)
        (tmp_path / "synth.obj").write_text("\n".join(obj_lines) + "\n")
        (tmp_path / "synth.mtl").write_text("\n".join(mtl_lines) + "\n")
        try:
            o = api.load_obj(str(tmp_path / "synth.obj"))
            assert o.vertices.ndim == 2 and all(m.triangles.shape[1] == 9 for m in o.meshes)
            for m in o.meshes:  # what comes back is indexable
                assert m.triangles[:, 0:3].max(initial=0) < len(o.vertices)
        except draw_b200.DrawError:
            n_err += 1
    assert n_err > 0
