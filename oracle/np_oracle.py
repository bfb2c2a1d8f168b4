"""Second, independent restatement of the reference renderer in numpy float32 scalars.

TEST INFRASTRUCTURE ONLY.  Written directly from the Rust source (not from oracle.cpp) so
that a transcription error in either restatement shows up as a bit difference between the
two on small frames (tests/test_oracle_cross.py).  Pure-Python loops: use it for frames of a
few thousand pixels and a few hundred triangles.

Every arithmetic step is one float32 operation on np.float32 scalars / arrays, in the order
the reference evaluates it (no FMA, no reassociation).  Reference lines are cited as
file:line relative to mororo18/draw.
"""
import ctypes
import ctypes.util

import numpy as np

F = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.tanf.restype = ctypes.c_float
_libm.tanf.argtypes = [ctypes.c_float]

EPS = F(0.0000001)  # linalg.rs:6
ZERO, ONE = F(0.0), F(1.0)


def v3(x, y, z):
    return np.array([x, y, z], F)


def dot(a, b):  # linalg.rs:182-184
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))


def norm(a):  # linalg.rs:167-171
    return F(np.sqrt(F(F(F(a[0] * a[0]) + F(a[1] * a[1])) + F(a[2] * a[2]))))


def normalized(a):  # linalg.rs:173-175
    with np.errstate(invalid="ignore", divide="ignore"):
        return (a / norm(a)).astype(F)


def cross(a, b):  # linalg.rs:186-200
    return v3(F(a[1] * b[2]) - F(a[2] * b[1]), F(a[2] * b[0]) - F(a[0] * b[2]), F(a[0] * b[1]) - F(a[1] * b[0]))


def matmul(a, b):  # linalg.rs:328-344
    c = np.zeros((4, 4), F)
    for i in range(4):
        for j in range(4):
            acc = ZERO
            for k in range(4):
                acc = F(acc + F(a[i, k] * b[k, j]))
            c[i, j] = acc
    return c


def matvec(m, v):  # linalg.rs:346-360
    out = np.zeros(4, F)
    for i in range(4):
        acc = ZERO
        for j in range(4):
            acc = F(acc + F(m[i, j] * v[j]))
        out[i] = acc
    return out


def sat_usize(v):  # Rust `f32 as usize`
    if np.isnan(v) or v <= 0:
        return 0
    if v >= F(18446744073709551616.0):
        return 2 ** 64 - 1
    return int(v)


def sat_u8(v):  # Rust `f32 as u8`
    if np.isnan(v) or v <= 0:
        return 0
    if v >= 255:
        return 255
    return int(v)


def total_key(v):  # f32::total_cmp
    i = int(np.array([v], F).view(np.int32)[0])
    if i < 0:
        i ^= 0x7FFFFFFF
    return i


class Camera:
    def __init__(self, pos, direction, ratio):  # scene/mod.rs:297-357
        near = F(-10.0)
        far = F(near - F(500.0))
        fov_x_rad = F(F(135.0) * F(F(np.pi) / F(180.0)))  # f32::to_radians
        self.right = F(abs(near) * F(_libm.tanf(F(fov_x_rad / F(2.0)))))
        self.left = F(-self.right)
        self.top = F(F(ONE / F(ratio)) * self.right)
        self.bottom = F(-self.top)
        self.position = np.array(pos, F)
        self.direction = normalized(np.array(direction, F))
        self.up = v3(0.0, 1.0, 0.0)
        self.near, self.far = near, far
        self.u = self.v = self.w = v3(0, 0, 0)

    def update_basis(self):  # scene/mod.rs:438-451
        g = self.direction
        w = ((g / norm(g)).astype(F) * F(-1.0)).astype(F)
        t_x_w = cross(self.up, w)
        u = (t_x_w / norm(t_x_w)).astype(F)
        v = cross(w, u)
        self.u, self.v, self.w = normalized(u), normalized(v), normalized(w)

    def basis_matrix(self):  # scene/mod.rs:415-428
        u, v, w = self.u, self.v, self.w
        return np.array([[u[0], v[0], w[0], 0], [u[1], v[1], w[1], 0], [u[2], v[2], w[2], 0], [0, 0, 0, 1]], F)

    def gen_matrix(self):  # scene/mod.rs:453-479
        p = self.position
        mp = np.array([[1, 0, 0, -p[0]], [0, 1, 0, -p[1]], [0, 0, 1, -p[2]], [0, 0, 0, 1]], F)
        self.update_basis()
        return matmul(self.basis_matrix().T.copy(), mp)

    def gen_view_planes(self):  # scene/mod.rs:481-593
        self.update_basis()
        mb = self.basis_matrix()
        cam = self.position
        n, f, r, l, t, b = self.near, self.far, self.right, self.left, self.top, self.bottom

        def world(x, y, z):
            p4 = matvec(mb, np.array([x, y, z, 1.0], F))
            return ((p4[:3] / p4[3]).astype(F) + cam).astype(F)

        ur, ul, lr, ll = world(r, t, n), world(l, t, n), world(r, b, n), world(l, b, n)
        xc, yc = F(F(l + r) / F(2.0)), F(F(b + t) / F(2.0))
        up_far = world(xc, F(F(f * t) / n), f)
        lo_far = world(xc, F(F(f * b) / n), f)
        ri_far = world(F(F(f * r) / n), yc, f)
        le_far = world(F(F(f * l) / n), yc, f)
        visible = ((ur + lo_far).astype(F) / F(2.0)).astype(F)
        depth = [plane_new(ur, lr, ll, visible), plane_new(le_far, ri_far, up_far, visible)]
        lateral = [plane_new(ri_far, ur, lr, visible), plane_new(le_far, ll, ul, visible),
                   plane_new(up_far, ul, ur, visible), plane_new(lo_far, ll, lr, visible)]
        return depth, lateral


def plane_new(a, b, c, visible):  # scene/mod.rs:603-632
    p, q = (b - a).astype(F), (c - b).astype(F)
    n = cross(p, q)
    k = F(-dot(n, a))
    if F(dot(n, visible) + k) < 0:
        n = cross(q, p)
        k = F(-dot(n, a))
    return n, k


def plane_func(pl, p):  # scene/mod.rs:634-636
    return F(dot(pl[0], p) + pl[1])


# A vertex's attributes travel as one float32 vector so that Add/Sub/Mul (canvas.rs:242-291)
# are elementwise: [normal3, light3, halfway3, uv3, screen2, depth] = 15 floats.
N0, L0, H0, T0, S0, D0 = 0, 3, 6, 9, 12, 14


def plane_clip(pl, tri):  # scene/mod.rs:662-746 ; tri = (verts[3], attrs[3])
    (a, b, c), (aa, ba, ca) = tri
    fa, fb, fc = plane_func(pl, a), plane_func(pl, b), plane_func(pl, c)
    if fa > 0 and fb > 0 and fc > 0:
        return [tri]
    if fa <= 0 and fb <= 0 and fc <= 0:
        return []
    if F(fa * fc) >= 0:
        fb, fc = fc, fb
        b, c = c, b
        ba, ca = ca, ba
        fa, fb = fb, fa
        a, b = b, a
        aa, ba = ba, aa
    elif F(fb * fc) >= 0:
        fa, fc = fc, fa
        a, c = c, a
        aa, ca = ca, aa
        fa, fb = fb, fa
        a, b = b, a
        aa, ba = ba, aa
    n = pl[0]
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        t_a = F(F(plane_func(pl, a) / dot(n, (a - c).astype(F))) - EPS)
        new_a = (a + ((c - a).astype(F) * t_a).astype(F)).astype(F)
        new_aa = (aa + ((ca - aa).astype(F) * t_a).astype(F)).astype(F)
        t_b = F(F(plane_func(pl, b) / dot(n, (b - c).astype(F))) - EPS)
        new_b = (b + ((c - b).astype(F) * t_b).astype(F)).astype(F)
        new_ba = (ba + ((ca - ba).astype(F) * t_b).astype(F)).astype(F)
    if fc <= 0:
        return [((a, new_a, new_b), (aa, new_aa, new_ba)), ((a, b, new_b), (aa, ba, new_ba))]
    return [((c, new_a, new_b), (ca, new_aa, new_ba))]


def clip_against_planes(tri, planes):  # scene/mod.rs:43-90
    depth, lateral = planes
    pool = []
    ok = True
    for pl in lateral:  # at_least_partially_visible, :641-660
        f = [plane_func(pl, p) for p in tri[0]]
        if f[0] <= 0 and f[1] <= 0 and f[2] <= 0:
            ok = False
            break
    if ok:
        pool = [tri]
    for pl in depth:
        new_pool = []
        for t in pool:
            new_pool.extend(plane_clip(pl, t))
        pool = new_pool
    return pool


class Canvas:
    def __init__(self, width, height):  # canvas.rs:366-380
        self.width, self.height = width, height
        self.frame = np.zeros((height, width, 4), np.uint8)
        self.frame[..., 3] = 255  # Pixel::black()
        self.offset = (F(0.0), F(0.0))
        self.depth_frame = None
        self.depth_max = F(0.0)
        self.depth_update = False

    def init_depth(self, d):  # canvas.rs:403-411
        self.depth_max = F(d)
        self.depth_frame = np.full((self.height, self.width), F(d), F)

    def apply_offset(self, x, y):  # canvas.rs:382-385
        self.offset = (F(x), F(y))

    def clear(self):  # canvas.rs:425-433 ; azul_bb = (r155,g186,b255), memory order b,g,r,pad
        self.frame[...] = np.array([255, 186, 155, 255], np.uint8)
        if self.depth_frame is not None:
            self.init_depth(self.depth_max)

    def draw_pixel(self, x, y, rgb, opacity, depth):  # canvas.rs:906-960
        row = self.height - y - 1
        if opacity < 1:
            bg = self.frame[row, x]  # b,g,r,pad
            k0, k1 = F(ONE - opacity), opacity
            r = (sat_u8(F(F(bg[2]) * k0)) + sat_u8(F(F(rgb[0]) * k1))) & 255
            g = (sat_u8(F(F(bg[1]) * k0)) + sat_u8(F(F(rgb[1]) * k1))) & 255
            b = (sat_u8(F(F(bg[0]) * k0)) + sat_u8(F(F(rgb[2]) * k1))) & 255
            new = (b, g, r, 0)
        else:
            new = (rgb[2], rgb[1], rgb[0], 255)
        if depth < self.depth_frame[y, x]:
            self.frame[row, x] = new
            if self.depth_update:
                self.depth_frame[y, x] = depth

    def draw_triangle(self, A, B, C, tex):  # canvas.rs:577-750 ; A,B,C = 15-float attribute vectors
        def center(p):  # :585-587 with pos_map_center :896-904 ; Vec2 sub = add of negation
            return (F(np.floor(F(F(p[S0] + F(-self.offset[0])) + F(0.5)))),
                    F(np.floor(F(F(p[S0 + 1] + F(-self.offset[1])) + F(0.5)))))

        with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
            (ax, ay), (bx, by), (cx, cy) = center(A), center(B), center(C)

            def edge(px, py, qx, qy):  # :597-616
                c0, c1, k1, k2 = F(py - qy), F(qx - px), F(px * qy), F(qx * py)
                return lambda x, y: F(F(F(F(c0 * x) + F(c1 * y)) + k1) - k2)

            f_ab, f_bc, f_ca = edge(ax, ay, bx, by), edge(bx, by, cx, cy), edge(cx, cy, ax, ay)

            def min3(*v):
                r = F(np.inf)
                for e in v:
                    if e < r:
                        r = e
                return r

            def max3(*v):
                r = F(-np.inf)
                for e in v:
                    if e > r:
                        r = e
                return r

            x0, y0 = sat_usize(min3(ax, bx, cx)), sat_usize(min3(ay, by, cy))
            x1, y1 = sat_usize(max3(ax, bx, cx)), sat_usize(max3(ay, by, cy))

            def from_coords(a0, b0, a1, b1):  # canvas.rs:315-330
                return (min(a0, a1), min(b0, b1), max(a0, a1), max(b0, b1))

            def clip(a, b):  # canvas.rs:332-350
                xm, ym = max(a[0], b[0]), max(a[1], b[1])
                xM, yM = min(a[2], b[2]), min(a[3], b[3])
                if xm > xM:
                    xm = xM = 0
                if ym > yM:
                    ym = yM = 0
                return from_coords(xm, ym, xM, yM)

            screen = from_coords(0, 0, self.width - 1, self.height - 1)
            drawable = clip(from_coords(x0, y0, x1, y1), screen)
            x0, y0, x1, y1 = clip(screen, drawable)

            f_alpha, f_beta, f_gama = f_bc(ax, ay), f_ca(bx, by), f_ab(cx, cy)
            m1 = F(-1.0)
            o_alpha, o_beta, o_gama = f_bc(m1, m1), f_ca(m1, m1), f_ab(m1, m1)

            for y in range(y0, y1 + 1):
                yf = F(y)
                for x in range(x0, x1 + 1):
                    xf = F(x)
                    alpha = F(f_bc(xf, yf) / f_alpha)
                    beta = F(f_ca(xf, yf) / f_beta)
                    gama = F(f_ab(xf, yf) / f_gama)
                    if not (alpha >= 0 and beta >= 0 and gama >= 0):
                        continue
                    if not ((alpha > 0 or F(f_alpha * o_alpha) > 0) and (beta > 0 or F(f_beta * o_beta) > 0)
                            and (gama > 0 or F(f_gama * o_gama) > 0)):
                        continue
                    depth = F(F(F(alpha * A[D0]) + F(beta * B[D0])) + F(gama * C[D0]))
                    P = (((A * alpha).astype(F) + (B * beta).astype(F)).astype(F) + (C * gama).astype(F)).astype(F)
                    uv = P[T0:T0 + 3]
                    dcol = texel(tex["map_kd"], uv[0], uv[1])
                    acol = texel(tex["map_ka"], uv[0], uv[1])
                    N, L, H = P[N0:N0 + 3], P[L0:L0 + 3], P[H0:H0 + 3]
                    c_l = tex["ks"]
                    c_r = (dcol * tex["kd"]).astype(F)
                    c_a = (acol * tex["ka"]).astype(F)
                    ln = dot(L, N)
                    s = F(ONE - (ln if ln > 0 else ZERO))  # 0.0_f32.max(x): NaN -> 0
                    hn = dot(H, N)
                    col = ((c_r * (c_a + (c_l * s).astype(F)).astype(F)).astype(F)
                           + (c_l * F(hn * hn)).astype(F)).astype(F)
                    scaled = (col * F(255.0)).astype(F)
                    rgb = (sat_u8(scaled[0]), sat_u8(scaled[1]), sat_u8(scaled[2]))
                    self.draw_pixel(x, y, rgb, tex["alpha"], depth)


    def draw_triangle_2d(self, va, vb, vc, tex, clipping_rect=None):  # canvas.rs:435-575 (the GUI's path)
        """va, vb, vc: (x, y, u, v, (r, g, b), alpha); tex: uint8 [h, w, 4]; clipping_rect: (x0, y0, x1, y1) or None."""
        def mul_px(p, f):  # Mul<f32> for Pixel, canvas.rs:154-169: per channel (c as f32 * f) as u8, in r, g, b order
            return tuple(sat_u8(F(F(c) * f)) for c in p)

        def add_px(p, q):  # Add for Pixel, canvas.rs:136-152: u8 add (wraps in a release build)
            return tuple((a + b) & 255 for a, b in zip(p, q))

        with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
            ctr = lambda v: (F(np.floor(F(F(v[0]) + F(0.5)))), F(np.floor(F(F(v[1]) + F(0.5)))))  # pos_map_center :896-904
            (ax, ay), (bx, by), (cx, cy) = ctr(va), ctr(vb), ctr(vc)

            def edge(px, py, qx, qy):  # :461-480: (p.y - q.y) * x + (q.x - p.x) * y + p.x * q.y - q.x * p.y
                c0, c1, k1, k2 = F(py - qy), F(qx - px), F(px * qy), F(qx * py)
                return lambda x, y: F(F(F(F(c0 * x) + F(c1 * y)) + k1) - k2)

            f_ab, f_bc, f_ca = edge(ax, ay, bx, by), edge(bx, by, cx, cy), edge(cx, cy, ax, ay)

            def min3(*v):  # :482-491
                r = F(np.inf)
                for e in v:
                    if e < r:
                        r = e
                return r

            def max3(*v):  # :493-502
                r = F(-np.inf)
                for e in v:
                    if e > r:
                        r = e
                return r

            def from_coords(a0, b0, a1, b1):  # canvas.rs:315-330
                return (min(a0, a1), min(b0, b1), max(a0, a1), max(b0, b1))

            def clip(a, b):  # canvas.rs:332-350
                xm, ym = max(a[0], b[0]), max(a[1], b[1])
                xM, yM = min(a[2], b[2]), min(a[3], b[3])
                if xm > xM:
                    xm = xM = 0
                if ym > yM:
                    ym = yM = 0
                return from_coords(xm, ym, xM, yM)

            screen = from_coords(0, 0, self.width - 1, self.height - 1)
            drawable = clip(from_coords(sat_usize(min3(ax, bx, cx)), sat_usize(min3(ay, by, cy)),
                                        sat_usize(max3(ax, bx, cx)), sat_usize(max3(ay, by, cy))), screen)  # :504-514
            x0, y0, x1, y1 = clip(from_coords(*clipping_rect) if clipping_rect is not None else screen, drawable)  # :516-517
            f_alpha, f_beta, f_gama = f_bc(ax, ay), f_ca(bx, by), f_ab(cx, cy)
            m1 = F(-1.0)
            o_alpha, o_beta, o_gama = f_bc(m1, m1), f_ca(m1, m1), f_ab(m1, m1)
            th, tw, _ = tex.shape
            for y in range(y0, y1 + 1):
                yf = F(y)
                for x in range(x0, x1 + 1):
                    xf = F(x)
                    alpha, beta, gama = F(f_bc(xf, yf) / f_alpha), F(f_ca(xf, yf) / f_beta), F(f_ab(xf, yf) / f_gama)
                    if not (alpha >= 0 and beta >= 0 and gama >= 0):
                        continue
                    if not ((alpha > 0 or F(f_alpha * o_alpha) > 0) and (beta > 0 or F(f_beta * o_beta) > 0)
                            and (gama > 0 or F(f_gama * o_gama) > 0)):
                        continue
                    col = add_px(add_px(mul_px(va[4], alpha), mul_px(vb[4], beta)), mul_px(vc[4], gama))  # :546
                    c_alpha = F(F(F(alpha * F(va[5])) + F(beta * F(vb[5]))) + F(gama * F(vc[5])))          # :548-550
                    u = F(F(F(F(va[2]) * alpha) + F(F(vb[2]) * beta)) + F(F(vc[2]) * gama))                  # :552
                    v = F(F(F(F(va[3]) * alpha) + F(F(vb[3]) * beta)) + F(F(vc[3]) * gama))
                    ui = min(sat_usize(F(np.floor(F(u * F(tw))))), tw - 1)  # get_rgba_slice, scene/mod.rs:137-152 (+ clamp)
                    vi = th - 1 - min(sat_usize(F(np.floor(F(v * F(th))))), th - 1)
                    px = tex[vi, ui]
                    t_alpha = F(F(px[3]) / F(255.0))                                                       # :555
                    col = add_px(mul_px(col, t_alpha), mul_px((int(px[0]), int(px[1]), int(px[2])), F(ONE - t_alpha)))  # :562-563
                    opacity = F(c_alpha * t_alpha)                                                       # :566
                    row = self.height - y - 1                                                            # draw_pixel_coord_with_depth(.., 0.0) :568
                    if opacity < 1:
                        bg = self.frame[row, x]
                        new = add_px(mul_px((int(bg[2]), int(bg[1]), int(bg[0])), F(ONE - opacity)), mul_px(col, opacity))
                    else:
                        new = col
                    if F(0.0) < self.depth_frame[y, x]:
                        self.frame[row, x] = (new[2], new[1], new[0], 0)  # every Pixel here came out of Add: padd = 0
                        if self.depth_update:
                            self.depth_frame[y, x] = F(0.0)


def texel(tmap, u, v):  # scene/mod.rs:154-168 (+ the never-triggering clamp, deviation 6)
    h, w, _ = tmap.shape
    ui = min(sat_usize(F(np.floor(F(u * F(F(w) - ONE))))), w - 1)
    vr = min(sat_usize(F(np.floor(F(v * F(F(h) - ONE))))), h - 1)
    px = tmap[h - 1 - vr, ui]
    return np.array([F(px[0]) / F(255.0), F(px[1]) / F(255.0), F(px[2]) / F(255.0)], F)


_WHITE = np.full((1, 1, 3), 255, np.uint8)


def _tex(t):
    return {"ka": np.array(t.ka, F), "kd": np.array(t.kd, F), "ks": np.array(t.ks, F), "alpha": F(t.alpha),
            "map_ka": _WHITE if t.map_ka is None else np.asarray(t.map_ka, np.uint8),
            "map_kd": _WHITE if t.map_kd is None else np.asarray(t.map_kd, np.uint8)}


class Scene:
    def __init__(self, width, height):  # scene/mod.rs:760-786
        self.width, self.height = width, height
        pos = v3(0.0, 0.0, 150.0)
        self.light = v3(0.0, 300.0, 300.0)
        self.camera = Camera(pos, (pos * F(-1.0)).astype(F), F(F(width) / F(height)))
        self.objects = []

    def set_camera(self, pos, direction):
        self.camera = Camera(pos, direction, F(F(self.width) / F(self.height)))

    def add_obj(self, obj):  # object.rs:34-71
        texs = [_tex(t) for t in obj.textures]
        opaque = [m for m in obj.meshes if not texs[m.texture_idx]["alpha"] < 1]
        transp = [m for m in obj.meshes if texs[m.texture_idx]["alpha"] < 1]
        self.objects.append({"v": np.asarray(obj.vertices, F), "n": np.asarray(obj.normals_vertices, F),
                             "t": np.asarray(obj.texture_vertices, F), "tex": texs,
                             "opaque": [(np.asarray(m.triangles, np.int64).reshape(-1, 9), m.texture_idx) for m in opaque],
                             "transp": [[np.asarray(m.triangles, np.int64).reshape(-1, 9), m.texture_idx] for m in transp]})

    def transformation_matrix(self):  # scene/mod.rs:817-899
        c = self.camera
        nx, ny = F(self.width), F(self.height)
        n, f, r, l, t, b = c.near, c.far, c.right, c.left, c.top, c.bottom
        cam = c.gen_matrix()
        persp = np.array([[n, 0, 0, 0], [0, n, 0, 0], [0, 0, F(n + f), F(-F(n * f))], [0, 0, 1, 0]], F)
        orth = np.array([[F(F(2.0) / F(r - l)), 0, 0, F(F(-F(r + l)) / F(r - l))],
                         [0, F(F(2.0) / F(t - b)), 0, F(F(-F(t + b)) / F(t - b))],
                         [0, 0, F(F(2.0) / F(n - f)), F(F(-F(n + f)) / F(n - f))],
                         [0, 0, 0, 1]], F)
        vp = np.array([[F(nx / F(2.0)), 0, 0, F(F(nx - ONE) / F(2.0))],
                       [0, F(ny / F(2.0)), 0, F(F(ny - ONE) / F(2.0))],
                       [0, 0, 1, 0], [0, 0, 0, 1]], F)
        return matmul(matmul(matmul(vp, orth), persp), cam)

    def render(self, canvas):  # scene/mod.rs:901-1249
        canvas.clear()
        M = self.transformation_matrix()
        cam = self.camera.position
        planes = self.camera.gen_view_planes()
        for obj in self.objects:
            V = obj["v"]
            vis = []
            with np.errstate(invalid="ignore", divide="ignore"):
                for p in V:  # :917-926
                    eye_dir = (p - cam).astype(F)
                    light = normalized((p - self.light).astype(F))
                    eye = normalized(eye_dir)
                    depth = norm(eye_dir)
                    halfway = normalized((light + eye).astype(F))
                    vis.append((light, halfway, depth))
            canvas.depth_update = True
            for tris, ti in obj["opaque"]:
                for t in tris:
                    self._draw(canvas, obj, vis, t, obj["tex"][ti], M, planes, cam, cull=True)
            canvas.depth_update = False
            for mesh in obj["transp"]:
                tris, ti = mesh

                def key(t):  # :1100-1115
                    c = ((((V[t[0]] + V[t[1]]).astype(F) + V[t[2]]).astype(F)) / F(3.0)).astype(F)
                    return -total_key(norm((c - cam).astype(F)))

                order = sorted(range(len(tris)), key=lambda i: key(tris[i]))  # stable
                mesh[0] = tris = tris[order]
                for t in tris:
                    self._draw(canvas, obj, vis, t, obj["tex"][ti], M, planes, cam, cull=False)

    def _draw(self, canvas, obj, vis, t, tex, M, planes, cam, cull):  # :933-1085 / :1117-1246
        verts, attrs = [], []
        for i in range(3):
            vi = vis[t[i]]
            a = np.zeros(15, F)
            a[N0:N0 + 3] = obj["n"][t[6 + i]]
            a[L0:L0 + 3] = vi[0]
            a[H0:H0 + 3] = vi[1]
            a[T0:T0 + 3] = obj["t"][t[3 + i]]
            a[D0] = vi[2]
            verts.append(obj["v"][t[i]])
            attrs.append(a)
        a, b, c = verts
        if cull:  # :1016-1027
            n = cross((b - a).astype(F), (c - b).astype(F))
            s = (np.zeros(3, F) + a).astype(F)
            s = (s + b).astype(F)
            s = (s + c).astype(F)
            center = (s / F(3.0)).astype(F)
            if dot((cam - center).astype(F), n) <= 0:
                return
        for cv, ca in clip_against_planes((tuple(verts), tuple(attrs)), planes):
            out = []
            with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
                for p, at in zip(cv, ca):  # :1047-1063
                    p4 = matvec(M, np.array([p[0], p[1], p[2], 1.0], F))
                    at = at.copy()
                    at[S0] = F(p4[0] / p4[3])
                    at[S0 + 1] = F(p4[1] / p4[3])
                    out.append(at)
            canvas.draw_triangle(out[0], out[1], out[2], tex)
