"""Host-side scene data types, mirroring the reference's input types.

`Object` / `IndexedMesh` / `Texture` carry exactly what `Object::new` takes in the reference
(src/renderer/scene/object.rs:34-41, scene/mesh.rs:31-35, scene/mod.rs:206-216): positions,
normals and uvs as float32 [N,3], index triples per mesh, and one `Texture` (Phong
coefficients + two optional maps) per material.  They are plain numpy containers; the device
copy is made by `Scene.add_obj` through the C ABI (include/draw_b200.h).
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


@dataclass
class Texture:
    """scene/mod.rs:206-216.  A map is None (the reference's 1x1 white TextureMap::default(),
    scene/mod.rs:128-135) or a uint8 array [height, width, components], components 3 or 4,
    row 0 = top of the image (what stb_image returns, scene/mod.rs:193-201)."""
    name: str = "default"
    ka: np.ndarray = field(default_factory=lambda: np.array([0.9, 0.9, 0.9], np.float32))
    kd: np.ndarray = field(default_factory=lambda: np.array([0.4, 0.4, 0.4], np.float32))
    ks: np.ndarray = field(default_factory=lambda: np.array([0.5, 0.5, 0.5], np.float32))
    alpha: float = 1.0
    map_ka: Optional[np.ndarray] = None
    map_kd: Optional[np.ndarray] = None


@dataclass
class IndexedMesh:
    """scene/mesh.rs:31-35.  triangles: uint32 [T, 9] = (v0 v1 v2, t0 t1 t2, n0 n1 n2), the
    reference's (vertex, texture, normal) index triples flattened."""
    name: str
    triangles: np.ndarray
    texture_idx: int = 0


@dataclass
class Object:
    """scene/object.rs:18-31.  Meshes are kept in file order; the split into opaque and
    transparent lists (material alpha < 1, object.rs:45-53) happens inside the library."""
    name: str
    vertices: np.ndarray
    normals_vertices: np.ndarray
    texture_vertices: np.ndarray
    meshes: List[IndexedMesh]
    textures: List[Texture]

    def triangle_count(self):
        return int(sum(m.triangles.shape[0] for m in self.meshes))

    def translated(self, dx, dy, dz):
        """Copy with `vertices` shifted (vertices is a pub field in the reference, object.rs:21;
        the C3 scene places its three models side by side this way, SURVEY.md §8c dev. 5)."""
        v = (self.vertices + np.array([dx, dy, dz], np.float32)).astype(np.float32)
        return Object(self.name, v, self.normals_vertices, self.texture_vertices, self.meshes, self.textures)
