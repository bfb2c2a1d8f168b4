#!/usr/bin/env python
"""Generate the committed scene fixtures under tests/golden/scenes/ from the reference's assets.

Runs HERE (needs /root/reference/models, PIL for the PNGs' cross-check and the built library for the JPEG); the GPU box only reads the generated files.
Loader = oracle/obj_loader.py, the numpy restatement of object.rs:106-454.  Re-run after any
loader change:  python tests/golden/make_scene_cache.py [--ref /root/reference]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from draw_b200 import scene_cache, synthetic  # noqa: E402
from draw_b200.model import IndexedMesh, Object, Texture  # noqa: E402
from oracle import obj_loader  # noqa: E402


def to_model(o):
    return Object(o.name, o.vertices, o.normals_vertices, o.texture_vertices,
                  [IndexedMesh(m.name, m.triangles, m.texture_idx) for m in o.meshes],
                  [Texture(t.name, t.ka, t.kd, t.ks, t.alpha, t.map_ka, t.map_kd) for t in o.textures])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "scenes"))
    ap.add_argument("--only", default=None, help="write just this scene (e.g. c1_lemur_airplane)")
    args = ap.parse_args()
    models = os.path.join(args.ref, "models")
    os.makedirs(args.out, exist_ok=True)

    def load(rel):
        return to_model(obj_loader.load_from_file(os.path.join(models, rel)))

    donut = load("donut/donut.obj")
    lemur = load("lemur/lemur.obj")
    soldier = load("soldier1/soldier1.obj")
    skeleton = load("skeleton/fgc_skeleton.obj")
    dungeon = [to_model(o) for o in obj_loader.load_from_directory(os.path.join(models, "dungeon_set"))]
    # the one JPEG among the assets: decoded by the library (draw_b200/csrc/jpeg_decode.cpp, stb_image's arithmetic)
    import draw_b200
    jpg = draw_b200.load_image(os.path.join(models, "airplane", "11804_Airplane_diff.jpg"))
    airplane = synthetic.airplane_standin(jpg)

    scenes = {
        # C1: PR1 reference frame (800x600, CPU): real textured asset + the airplane stand-in
        "c1_lemur_airplane": [lemur.translated(-60.0, 0.0, 0.0), airplane.translated(60.0, 0.0, 0.0)],
        "c2_donut": [donut],
        # C3: SURVEY.md §8c deviation 5 — the three models side by side
        "c3_trio": [soldier.translated(-150.0, 0.0, 0.0), skeleton, lemur.translated(150.0, 0.0, 0.0)],
        "c4_dungeon": dungeon,
    }
    for name, objs in scenes.items():
        if args.only and name != args.only:
            continue
        path = os.path.join(args.out, name + ".npz")
        scene_cache.save(path, objs)
        tris = sum(o.triangle_count() for o in objs)
        verts = sum(o.vertices.shape[0] for o in objs)
        print(f"{name}: {len(objs)} objects, {tris} triangles, {verts} vertices, "
              f"{os.path.getsize(path) / 1e6:.2f} MB")
    if not args.only:
        np.save(os.path.join(os.path.dirname(args.out), "c4_camera_path.npy"), synthetic.flythrough_camera(120))
        np.save(os.path.join(os.path.dirname(args.out), "orbit_camera_path.npy"), synthetic.orbit_camera(64))


if __name__ == "__main__":
    main()
