"""Binary scene cache: a list of `Object`s in one compressed .npz.

The reference has no such format (its only input is OBJ/MTL text parsed on every load,
object.rs:106-454); this is the on-disk form the benchmarks and tests load so that no run on
the GPU box needs the reference tree or an image decoder.  Texture images are stored once in
a pool and referenced by index.
"""
import numpy as np

from .model import IndexedMesh, Object, Texture

FORMAT_VERSION = 1


def save(path, objects):
    arrays = {"version": np.array(FORMAT_VERSION), "n_objects": np.array(len(objects))}
    pool = []

    def pool_index(img):
        if img is None:
            return -1
        img = np.ascontiguousarray(img, dtype=np.uint8)
        for i, p in enumerate(pool):
            if p.shape == img.shape and np.array_equal(p, img):
                return i
        pool.append(img)
        return len(pool) - 1

    for i, o in enumerate(objects):
        p = f"o{i}_"
        arrays[p + "name"] = np.array(o.name)
        arrays[p + "vertices"] = np.asarray(o.vertices, np.float32).reshape(-1, 3)
        arrays[p + "normals"] = np.asarray(o.normals_vertices, np.float32).reshape(-1, 3)
        arrays[p + "uvs"] = np.asarray(o.texture_vertices, np.float32).reshape(-1, 3)
        arrays[p + "n_meshes"] = np.array(len(o.meshes))
        arrays[p + "n_textures"] = np.array(len(o.textures))
        for j, m in enumerate(o.meshes):
            arrays[f"{p}m{j}_name"] = np.array(m.name)
            arrays[f"{p}m{j}_tris"] = np.asarray(m.triangles, np.uint32).reshape(-1, 9)
            arrays[f"{p}m{j}_tex"] = np.array(m.texture_idx)
        for k, t in enumerate(o.textures):
            arrays[f"{p}t{k}_name"] = np.array(t.name)
            arrays[f"{p}t{k}_coef"] = np.concatenate([np.asarray(t.ka, np.float32), np.asarray(t.kd, np.float32),
                                                      np.asarray(t.ks, np.float32),
                                                      np.array([t.alpha], np.float32)])
            arrays[f"{p}t{k}_maps"] = np.array([pool_index(t.map_ka), pool_index(t.map_kd)])
    arrays["n_images"] = np.array(len(pool))
    for i, img in enumerate(pool):
        arrays[f"img{i}"] = img
    np.savez_compressed(path, **arrays)


def load(path):
    z = np.load(path, allow_pickle=False)
    if int(z["version"]) != FORMAT_VERSION:
        raise ValueError(f"{path}: unsupported scene cache version {int(z['version'])}")
    pool = [z[f"img{i}"] for i in range(int(z["n_images"]))]
    objects = []
    for i in range(int(z["n_objects"])):
        p = f"o{i}_"
        meshes = [IndexedMesh(str(z[f"{p}m{j}_name"]), z[f"{p}m{j}_tris"], int(z[f"{p}m{j}_tex"]))
                  for j in range(int(z[p + "n_meshes"]))]
        textures = []
        for k in range(int(z[p + "n_textures"])):
            c = z[f"{p}t{k}_coef"]
            ia, id_ = (int(x) for x in z[f"{p}t{k}_maps"])
            textures.append(Texture(str(z[f"{p}t{k}_name"]), c[0:3].copy(), c[3:6].copy(), c[6:9].copy(),
                                    float(c[9]), pool[ia] if ia >= 0 else None, pool[id_] if id_ >= 0 else None))
        objects.append(Object(str(z[p + "name"]), z[p + "vertices"], z[p + "normals"], z[p + "uvs"],
                              meshes, textures))
    return objects
