"""Procedural inputs for the configs that have no asset in the reference tree.

  - `torus(n_theta, n_phi)`     C5: the donut.obj generator scaled up: same
                                vertex and face enumeration as models/donut/donut.obj
                                (R=9, r=4), then the loader's post-processing — rescale to
                                radius 100 (object.rs:159-170) and smooth normals as the
                                normalised sum of un-normalised face normals (object.rs:394-411).
  - `checker_texture()`         C5: 1024x1024 RGB checker, per-cell tint from splitmix64.
  - `airplane_standin(...)`     C1: models/airplane/*.obj is missing from the reference tree
                                (.MISSING_LARGE_BLOBS); a two-material ellipsoid + canopy bound
                                to the airplane's MTL coefficients and JPG exercises the same
                                code (textured opaque body + `d 0.7` glass, transparent pass).
  - `flythrough_camera(k)`      C4: the 120-frame clipping-heavy camera path.
Everything is deterministic; nothing here is timed.
"""
import numpy as np

from .model import IndexedMesh, Object, Texture

F = np.float32


def _norm_rows(a):
    s = (a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1]).astype(F) + a[:, 2] * a[:, 2]
    return np.sqrt(s.astype(F)).astype(F)


def _rescale_100(v):
    """object.rs:159-170: v * (100 / max |v|), float32."""
    factor = F(100.0) / F(_norm_rows(v).max())
    return (v * factor).astype(F)


def _smooth_normals(vertices, tri_v):
    """object.rs:394-411 in the reference's accumulation order (face-major, corner-minor)."""
    a, b, c = vertices[tri_v[:, 0]], vertices[tri_v[:, 1]], vertices[tri_v[:, 2]]
    p, q = (b - a).astype(F), (c - b).astype(F)
    n = np.stack([(p[:, 1] * q[:, 2]).astype(F) - (p[:, 2] * q[:, 1]).astype(F),
                  (p[:, 2] * q[:, 0]).astype(F) - (p[:, 0] * q[:, 2]).astype(F),
                  (p[:, 0] * q[:, 1]).astype(F) - (p[:, 1] * q[:, 0]).astype(F)], 1).astype(F)
    gen = np.zeros_like(vertices, dtype=F)
    np.add.at(gen, tri_v.reshape(-1), np.repeat(n, 3, axis=0))  # sequential, in index order
    with np.errstate(invalid="ignore", divide="ignore"):
        return (gen / _norm_rows(gen)[:, None]).astype(F)


def torus(n_theta=32, n_phi=32, R=9.0, r=4.0, texture=None, name="torus"):
    """v[j*n_phi+i] = ((R + r cos phi_i) cos theta_j, (R + r cos phi_i) sin theta_j, r sin phi_i);
    faces (j,i),(j+1,i),(j+1,i+1) and (j,i),(j+1,i+1),(j,i+1), indices wrapping in both directions
    (the enumeration of models/donut/donut.obj)."""
    th = 2.0 * np.pi * np.arange(n_theta, dtype=np.float64) / n_theta
    ph = 2.0 * np.pi * np.arange(n_phi, dtype=np.float64) / n_phi
    ring = R + r * np.cos(ph)
    v = np.stack([np.outer(np.cos(th), ring), np.outer(np.sin(th), ring),
                  np.outer(np.ones_like(th), r * np.sin(ph))], -1).reshape(-1, 3).astype(F)
    j, i = np.meshgrid(np.arange(n_theta), np.arange(n_phi), indexing="ij")
    j1, i1 = (j + 1) % n_theta, (i + 1) % n_phi
    idx = lambda jj, ii: (jj * n_phi + ii).reshape(-1)
    t0 = np.stack([idx(j, i), idx(j1, i), idx(j1, i1)], 1)
    t1 = np.stack([idx(j, i), idx(j1, i1), idx(j, i1)], 1)
    tri_v = np.stack([t0, t1], 1).reshape(-1, 3).astype(np.uint32)

    v = _rescale_100(v)
    normals = _smooth_normals(v, tri_v)
    uv = np.stack([(j / n_theta).reshape(-1), (i / n_phi).reshape(-1), np.zeros(j.size)], 1).astype(F)
    tris = np.concatenate([tri_v, tri_v, tri_v], 1).astype(np.uint32)  # v, t, n share the vertex index
    textures = [Texture()]
    tex_idx = 0
    if texture is not None:
        textures.append(texture)
        tex_idx = 1
    return Object(name, v, normals, uv, [IndexedMesh("default", tris, tex_idx)], textures)


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x, z ^ (z >> 31)


def checker_texture(size=1024, cells=32, seed=0xD0A7):
    """RGB checker: two base colours, each cell XOR'd with a 3x5-bit tint from splitmix64(seed)."""
    base = np.array([[200, 72, 40], [40, 96, 208]], np.uint8)
    img = np.zeros((size, size, 3), np.uint8)
    step = size // cells
    state = seed
    for cy in range(cells):
        for cx in range(cells):
            state, rnd = _splitmix64(state)
            tint = np.array([rnd & 31, (rnd >> 5) & 31, (rnd >> 10) & 31], np.uint8)
            img[cy * step:(cy + 1) * step, cx * step:(cx + 1) * step] = base[(cx + cy) & 1] ^ tint
    return img


def checker_material():
    """C5 material: ka = kd = (1,1,1), ks = (.5,.5,.5), both maps = the checker."""
    img = checker_texture()
    return Texture("checker", np.ones(3, F), np.ones(3, F), np.full(3, 0.5, F), 1.0, img, img)


def _ellipsoid(n_lon, n_lat, radii, center, flip=False):
    lon = 2.0 * np.pi * np.arange(n_lon + 1) / n_lon
    lat = np.pi * (np.arange(n_lat + 1) / n_lat - 0.5)
    LO, LA = np.meshgrid(lon, lat, indexing="ij")
    unit = np.stack([np.cos(LA) * np.cos(LO), np.sin(LA), np.cos(LA) * np.sin(LO)], -1).reshape(-1, 3)
    v = (unit * np.array(radii) + np.array(center)).astype(F)
    n = unit / np.array(radii)
    n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(F)
    uv = np.stack([(LO / (2 * np.pi)).reshape(-1), (LA / np.pi + 0.5).reshape(-1), np.zeros(LO.size)], 1).astype(F)
    a = (np.arange(n_lon)[:, None] * (n_lat + 1) + np.arange(n_lat)[None, :]).reshape(-1)
    b, c, d = a + (n_lat + 1), a + (n_lat + 1) + 1, a + 1
    tri = np.concatenate([np.stack([a, c, b], 1), np.stack([a, d, c], 1)], 0)
    if flip:
        tri = tri[:, ::-1]
    return v, n, uv, tri.astype(np.uint32)


def airplane_standin(diffuse_map, body_coef=None, glass_coef=None):
    """Stand-in for the missing airplane mesh: fuselage ellipsoid (opaque, textured) plus a canopy
    ellipsoid (glass, alpha 0.7) sharing the airplane's JPG, coefficients from its MTL
    (models/airplane/11804_Airplane_v2_l2.mtl: Ka 1 1 1, Kd 1 1 1, Ks .54 .54 .54, d 1.0 / 0.7)."""
    bv, bn, buv, bt = _ellipsoid(32, 16, (3.0, 0.6, 0.6), (0.0, 0.0, 0.0))
    gv, gn, guv, gt = _ellipsoid(20, 10, (0.9, 0.5, 0.45), (1.2, 0.45, 0.0))
    off = bv.shape[0]
    v = _rescale_100(np.concatenate([bv, gv], 0))
    n = np.concatenate([bn, gn], 0)
    uv = np.concatenate([buv, guv], 0)
    body = np.concatenate([bt, bt, bt], 1)
    glass = np.concatenate([gt + off, gt + off, gt + off], 1).astype(np.uint32)
    one, ks = np.ones(3, F), np.full(3, 0.54, F)
    textures = [Texture(),
                Texture("01___11804_Airplane_body", one, one, ks, 1.0, diffuse_map, diffuse_map),
                Texture("02___11804_Airplane_glass", one, one, ks, float(F(0.7)), diffuse_map, diffuse_map)]
    return Object("airplane_standin", v, n, uv,
                  [IndexedMesh("body", body, 1), IndexedMesh("glass", glass, 2)], textures)


def flythrough_camera(n_frames=120):
    """C4 camera path: frame k -> Camera::new(pos_k, dir_k, W/H) with
    pos_k = (30 sin(2 pi k/120), 10, 150 - 2.5 k), dir_k = (0.3 sin(2 pi k/60), -0.05, -1),
    evaluated in float64 and rounded once to float32.  Returns float32 [n_frames, 6]."""
    k = np.arange(n_frames, dtype=np.float64)
    pos = np.stack([30.0 * np.sin(2 * np.pi * k / 120.0), np.full_like(k, 10.0), 150.0 - 2.5 * k], 1)
    dr = np.stack([0.3 * np.sin(2 * np.pi * k / 60.0), np.full_like(k, -0.05), np.full_like(k, -1.0)], 1)
    return np.concatenate([pos, dr], 1).astype(F)


def orbit_camera(n_frames=64):
    """C2 / C3 bench camera path: a small orbit about the default camera so that every step renders a different
    frame of the same scene: pos_k = (15 sin(2 pi k/n), 8 sin(4 pi k/n), 150), dir_k = -pos_k (looking at the
    origin), evaluated in float64 and rounded once to float32.  Frame 0 is the default camera
    (scene/mod.rs:763-765).  Returns float32 [n_frames, 6]."""
    k = np.arange(n_frames, dtype=np.float64)
    pos = np.stack([15.0 * np.sin(2 * np.pi * k / n_frames), 8.0 * np.sin(4 * np.pi * k / n_frames), np.full_like(k, 150.0)], 1)
    return np.concatenate([pos, -pos], 1).astype(F)


# VertexSimpleAttributes (canvas.rs:185-191) in draw_vertex2d's layout (draw_b200.api.VERTEX2D, oracle.pyoracle.VERTEX2D)
_VERTEX2D = np.dtype([("x", "<f4"), ("y", "<f4"), ("u", "<f4"), ("v", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("pad", "u1"),
                      ("alpha", "<f4")])


def font_atlas(width=256, height=64, seed=0xA71A5):
    """A stand-in for the GUI's font atlas (src/app/gui.rs: imgui's RGBA32 font texture): white RGB with an alpha
    channel of glyph-like blobs, plus a fully opaque white block at the top left (what ImGui's solid fills sample)."""
    rng = np.random.default_rng(seed)
    a = np.zeros((height, width, 4), np.uint8)
    a[..., :3] = 255
    yy, xx = np.mgrid[0:height, 0:width]
    alpha = (np.sin(xx * 0.9) * np.cos(yy * 0.7) * 0.5 + 0.5) * 255.0
    alpha[rng.random((height, width)) < 0.35] = 0.0
    a[..., 3] = alpha.astype(np.uint8)
    a[:8, :8, 3] = 255
    a[8:12, :8, :3] = rng.integers(0, 256, (4, 8, 3))  # a few coloured texels: the texture's RGB takes part in the blend
    return a


def gui_command_list(width, height, n_commands=6, quads_per_command=40, seed=1):
    """Draw commands shaped like src/app/gui.rs:382-485 feeds Canvas::draw_triangle: per command a clipping rectangle
    (x0, y0, x1, y1) or None and a VERTEX2D array of window-like quads (two triangles each: solid fills sampling the
    opaque white texel, glyph quads sampling the atlas) with some free triangles, degenerate and off-screen ones."""
    rng = np.random.default_rng(seed)
    out = []
    for c in range(n_commands):
        n = quads_per_command
        v = np.zeros(n * 6, _VERTEX2D)
        for q in range(n):
            kind = rng.random()
            x0, y0 = rng.uniform(-20, width), rng.uniform(-20, height)
            w, h = (rng.uniform(4, max(5.0, width / 3)), rng.uniform(4, max(5.0, height / 3))) if kind < 0.3 else (rng.uniform(3, 24), rng.uniform(5, 24))
            x1, y1 = x0 + w, y0 + h
            if kind < 0.3:
                u0, v0, u1, v1 = 0.01, 0.99, 0.02, 0.98   # the opaque white block (row 0 = top = v near 1)
            else:
                u0, v0 = rng.uniform(0, 0.9), rng.uniform(0, 0.9)
                u1, v1 = u0 + rng.uniform(0.01, 0.09), v0 + rng.uniform(0.01, 0.09)
            col = rng.integers(0, 256, 3)
            alpha = 1.0 if rng.random() < 0.4 else rng.integers(0, 256) / 255.0
            quad = [(x0, y0, u0, v0), (x1, y0, u1, v0), (x1, y1, u1, v1), (x0, y0, u0, v0), (x1, y1, u1, v1), (x0, y1, u0, v1)]
            if kind > 0.9:  # free triangles with per-vertex colours and half-pixel coordinates
                quad = [(rng.uniform(-30, width + 30), rng.uniform(-30, height + 30), rng.uniform(0, 0.99), rng.uniform(0, 0.99)) for _ in range(6)]
            for k, (x, y, tu, tv) in enumerate(quad):
                vc = rng.integers(0, 256, 3) if kind > 0.9 else col
                va = rng.integers(0, 256) / 255.0 if kind > 0.95 else alpha
                v[q * 6 + k] = (x, y, tu, tv, vc[0], vc[1], vc[2], 0, va)
        if c == 1:  # degenerate, far off-screen and non-finite vertices
            v[0:3]["x"] = v[0]["x"]
            v[0:3]["y"] = v[0]["y"]
            v[3:6]["x"] += 1.0e9
            v[6]["x"] = np.nan
            v[9]["y"] = np.inf
            v[12:15]["x"] = -5000.0
        clip = None
        if c % 3 == 1:
            cx0, cy0 = int(rng.integers(0, width // 2)), int(rng.integers(0, height // 2))
            clip = (cx0, cy0, int(rng.integers(cx0, width + 40)), int(rng.integers(cy0, height + 40)))
        elif c % 3 == 2:  # from_coords normalises swapped corners (gui.rs inverts y, so y0 > y1 is the usual case)
            clip = (int(rng.integers(width // 2, width)), int(rng.integers(height // 2, height)), int(rng.integers(0, width // 2)), int(rng.integers(0, height // 2)))
        out.append((clip, v))
    return out
