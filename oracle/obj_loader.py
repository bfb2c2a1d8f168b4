"""OBJ/MTL -> Object, restating the reference loader in numpy float32.

TEST INFRASTRUCTURE / FIXTURE GENERATOR.  Follows `Object::load_from_file`
(src/renderer/scene/object.rs:106-454), `Object::new` (object.rs:34-71) and
`Object::load_from_directory` (object.rs:73-103) of mororo18/draw.  The product's own loader
is the C++ one behind `draw_object_load_obj`; tests compare the two bit for bit.

Third-party code restated here: the `obj` crate (kvark/obj, git default branch, UNPINNED in the
reference's Cargo.toml:13; source not under /root/reference).  Its published parsing rules, as
relied on by object.rs:131-267, are restated from the crate's documented behaviour:
  - statements: v, vt (first two floats), vn, f, o, g, usemtl, mtllib, s (ignored), l (ignored)
  - an implicit object and group are both named "default"
  - `o` closes the current group and object; `g` closes the current group
  - `usemtl` on a group that already has a material closes it and opens a new group of the
    same name (a material switch is a new group)
  - face indices are 1-based, negative = relative to the current count
  - MTL: newmtl, Ka, Kd, Ks, d, map_Ka, map_Kd are consumed; other keys are skipped.

Documented deviations from the reference (all needed because the reference
would panic or is non-deterministic there):
  2. a material without Ka/Kd/Ks takes the corresponding Texture::default() value
     (scene/mod.rs:243-245) instead of panicking at object.rs:182-184;
  3. `l` elements and faces with fewer than 3 vertices are ignored;
  4. load_from_directory sorts paths lexicographically (read_dir order is unspecified).
Texture images are decoded with PIL (the reference uses stb_image through the `stb` crate,
scene/mod.rs:174-202); PNG decoding is lossless, JPEG decoding is decoder-specific, so image
bytes are part of the committed scene fixtures rather than re-decoded at test time.
"""
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

f32 = np.float32


@dataclass
class Texture:
    """scene/mod.rs:206-216; maps are None (=> TextureMap::default(), 1x1x3 white) or uint8 [h,w,c]."""
    name: str = "default"
    ka: np.ndarray = field(default_factory=lambda: np.array([0.9, 0.9, 0.9], f32))
    kd: np.ndarray = field(default_factory=lambda: np.array([0.4, 0.4, 0.4], f32))
    ks: np.ndarray = field(default_factory=lambda: np.array([0.5, 0.5, 0.5], f32))
    alpha: float = 1.0
    map_ka: Optional[np.ndarray] = None
    map_kd: Optional[np.ndarray] = None


@dataclass
class IndexedMesh:
    """mesh.rs:31-35; triangles uint32 [T,9] = (v0 v1 v2, t0 t1 t2, n0 n1 n2)."""
    name: str
    triangles: np.ndarray
    texture_idx: int


@dataclass
class Object:
    """object.rs:18-31 (meshes kept in file order; the opaque/transparent split of Object::new,
    object.rs:45-53, is applied by whoever consumes the object)."""
    name: str
    vertices: np.ndarray
    normals_vertices: np.ndarray
    texture_vertices: np.ndarray
    meshes: List[IndexedMesh]
    textures: List[Texture]


# ---------------------------------------------------------------- float32 helpers (linalg.rs)

def _norm_rows(a):
    """Vec3::norm per row (linalg.rs:167-171): sqrt((x*x + y*y) + z*z) in float32."""
    a = a.astype(f32, copy=False)
    s = (a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1]).astype(f32) + a[:, 2] * a[:, 2]
    return np.sqrt(s.astype(f32)).astype(f32)


def _normalize_rows(a):
    """Vec3::normalized (linalg.rs:173-175): three divisions by the norm (0/0 -> NaN kept)."""
    n = _norm_rows(a)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (a / n[:, None]).astype(f32)


# ---------------------------------------------------------------- obj crate restatement

class _Group:
    def __init__(self, name):
        self.name = name
        self.material = None
        self.polys = []  # list of list[(v, vt|None, vn|None)]

    def clone(self):
        g = _Group(self.name)
        g.material = self.material
        g.polys = list(self.polys)
        return g


def _parse_obj(path):
    position, texture, normal = [], [], []
    objects = []  # list of (name, [groups])
    mtllibs = []
    obj_name, obj_groups = "default", []
    group = None

    def fix(idx, count):
        i = int(idx)
        return i - 1 if i > 0 else count + i

    with open(path, "r", errors="replace") as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            words = line.split()
            if not words:
                continue
            key = words[0]
            if key == "v":
                position.append([f32(float(words[1])), f32(float(words[2])), f32(float(words[3]))])
            elif key == "vt":
                texture.append([f32(float(words[1])), f32(float(words[2]))])
            elif key == "vn":
                normal.append([f32(float(words[1])), f32(float(words[2])), f32(float(words[3]))])
            elif key == "f":
                poly = []
                for w in words[1:]:
                    parts = w.split("/")
                    v = fix(parts[0], len(position))
                    vt = fix(parts[1], len(texture)) if len(parts) > 1 and parts[1] != "" else None
                    vn = fix(parts[2], len(normal)) if len(parts) > 2 and parts[2] != "" else None
                    poly.append((v, vt, vn))
                if group is None:
                    group = _Group("default")
                group.polys.append(poly)
            elif key == "o":
                if group is not None:
                    obj_groups.append(group)
                    objects.append((obj_name, obj_groups))
                    group = None
                obj_name = line[1:].strip() if len(line) > 2 else "default"
                obj_groups = []
            elif key == "g":
                if group is not None:
                    obj_groups.append(group)
                    group = None
                if len(line) > 2:
                    group = _Group(line[2:].strip())
            elif key == "mtllib":
                mtllibs.extend(words[1:])
            elif key == "usemtl":
                g = group if group is not None else _Group("default")
                if g.material is not None:
                    obj_groups.append(g.clone())
                    g.polys = []
                g.material = words[1] if len(words) > 1 else None
                group = g
            else:
                continue  # s, l, comments, unknown
    if group is not None:
        obj_groups.append(group)
    objects.append((obj_name, obj_groups))
    return position, texture, normal, objects, mtllibs


def _parse_mtl(path):
    mats, cur = [], None
    with open(path, "r", errors="replace") as fh:
        for line in fh:
            words = line.split()
            if not words or words[0].startswith("#"):
                continue
            key = words[0]
            if key == "newmtl":
                cur = {"name": words[1] if len(words) > 1 else ""}
                mats.append(cur)
            elif cur is None:
                continue
            elif key in ("Ka", "Kd", "Ks"):
                cur[key.lower()] = np.array([float(words[1]), float(words[2]), float(words[3])], f32)
            elif key == "d":
                cur["d"] = f32(float(words[1]))
            elif key in ("map_Ka", "map_Kd"):
                cur[key.lower()] = words[1]
    return mats


def load_image(path):
    """TextureMap::load_from_file (scene/mod.rs:174-202): 3 or 4 components, row 0 = top."""
    from PIL import Image
    im = Image.open(path)
    if im.mode != "RGBA":  # stb keeps the file's own channel count; the reference accepts 3 or 4
        im = im.convert("RGB")
    return np.ascontiguousarray(np.asarray(im, dtype=np.uint8))


def load_from_file(filename, image_loader=load_image):
    """Object::load_from_file (object.rs:106-454)."""
    parent = os.path.dirname(filename)
    position, texture, normal, objects, mtllibs = _parse_obj(filename)

    obj_vertices = np.array(position, f32).reshape(-1, 3)                      # :141-145
    obj_normals = _normalize_rows(np.array(normal, f32).reshape(-1, 3))        # :146-150
    tex = np.array(texture, f32).reshape(-1, 2)
    obj_texture_uv = np.concatenate([tex, np.zeros((tex.shape[0], 1), f32)], 1)  # :151-155
    obj_texture_uv = [row for row in obj_texture_uv]

    # rescale (:159-170): factor = 100.0 / max norm ; v = v * factor
    vertex_max = _norm_rows(obj_vertices).max()
    factor = f32(100.0) / f32(vertex_max)
    obj_vertices = (obj_vertices * factor).astype(f32)

    textures = [Texture()]                                                     # :172
    for lib in mtllibs:                                                        # :177-221
        for m in _parse_mtl(os.path.join(parent, lib)):
            d = Texture()
            t = Texture(name=m["name"],
                        ka=m.get("ka", d.ka), kd=m.get("kd", d.kd), ks=m.get("ks", d.ks),  # deviation 2
                        alpha=float(m.get("d", f32(1.0))),
                        map_ka=image_loader(os.path.join(parent, m["map_ka"])) if "map_ka" in m else None,
                        map_kd=image_loader(os.path.join(parent, m["map_kd"])) if "map_kd" in m else None)
            textures.append(t)

    meshes = []
    normals_list = [obj_normals]
    n_normals = obj_normals.shape[0]
    for _oname, groups in objects:                                             # :230
        for group in groups:
            if not group.polys:                                                # :235
                continue
            tris = []  # (v[3], t[3]|None, n[3]|None)
            material_name = group.material if group.material is not None else "default"  # :248-256
            missing_tex = missing_nrm = False
            for face in group.polys:                                           # :261-364
                if len(face) < 3:
                    continue                                                   # deviation 3
                if len(face) > 4:
                    raise NotImplementedError("faces with more than 4 vertices (object.rs:361-363 todo!())")
                v = [p[0] for p in face]
                t = [p[1] for p in face]
                n = [p[2] for p in face]
                f_mt = any(x is None for x in t)
                f_mn = any(x is None for x in n)
                missing_tex |= f_mt
                missing_nrm |= f_mn
                tris.append(([v[0], v[1], v[2]],
                             None if f_mt else [t[0], t[1], t[2]],
                             None if f_mn else [n[0], n[1], n[2]]))
                if len(face) == 4:                                             # :333-360
                    tris.append(([v[2], v[3], v[0]],
                                 None if f_mt else [t[2], t[3], t[0]],
                                 None if f_mn else [n[2], n[3], n[0]]))
            if missing_tex:                                                    # :368-384
                dummy = len(obj_texture_uv)
                obj_texture_uv.append(np.zeros(3, f32))
                tris = [(v, t if t is not None else [dummy] * 3, n) for v, t, n in tris]
            texture_idx = 0                                                    # :366,387-392
            for i, tx in enumerate(textures):
                if tx.name == material_name:
                    texture_idx = i
                    break
            if missing_nrm:                                                    # :394-429
                gen = np.zeros((obj_vertices.shape[0], 3), f32)
                for v, _t, _n in tris:
                    a, b, c = obj_vertices[v[0]], obj_vertices[v[1]], obj_vertices[v[2]]
                    p, q = (b - a).astype(f32), (c - b).astype(f32)           # mesh.rs:15-28
                    nrm = np.array([f32(p[1] * q[2]) - f32(p[2] * q[1]),
                                    f32(p[2] * q[0]) - f32(p[0] * q[2]),
                                    f32(p[0] * q[1]) - f32(p[1] * q[0])], f32)
                    for idx in v:
                        gen[idx] = gen[idx] + nrm
                gen = _normalize_rows(gen)
                tris = [(v, t, n if n is not None else [v[0] + n_normals, v[1] + n_normals, v[2] + n_normals])
                        for v, t, n in tris]
                normals_list.append(gen)
                n_normals += gen.shape[0]
            arr = np.array([v + t + n for v, t, n in tris], np.uint32).reshape(-1, 9)
            meshes.append(IndexedMesh(group.name, arr, texture_idx))

    return Object(name=os.path.basename(filename),
                  vertices=obj_vertices,
                  normals_vertices=np.concatenate(normals_list, 0).astype(f32),
                  texture_vertices=np.array(obj_texture_uv, f32).reshape(-1, 3),
                  meshes=meshes, textures=textures)


def load_from_directory(dirname, image_loader=load_image):
    """Object::load_from_directory (object.rs:73-103), sorted (deviation 4)."""
    names = sorted(n for n in os.listdir(dirname)
                   if n.endswith(".obj") and os.path.isfile(os.path.join(dirname, n)))
    return [load_from_file(os.path.join(dirname, n), image_loader) for n in names]
