"""The library's PNG decoder (draw_b200/csrc/image_decode.cpp, TextureMap::load_from_file,
scene/mod.rs:174-202) against PIL on generated files: RGB / RGBA / palette (with and without tRNS) /
16-bit, sizes that exercise every scanline filter, and the error behaviour for what the reference rejects."""
import os

import numpy as np
import pytest

PIL = pytest.importorskip("PIL.Image")


def _lib_decode(path):
    import draw_b200.api as api
    return api.load_image(path)


def _noise(h, w, c, seed):
    rng = np.random.default_rng(seed)
    base = np.linspace(0, 255, w, dtype=np.float64)[None, :, None] * np.ones((h, 1, c))
    return np.clip(base + rng.normal(0, 25, (h, w, c)), 0, 255).astype(np.uint8)  # smooth + noise: PIL picks mixed filters


@pytest.mark.parametrize("mode,c", [("RGB", 3), ("RGBA", 4)])
@pytest.mark.parametrize("size", [(1, 1), (7, 5), (64, 33), (257, 130)])
def test_png_truecolour_matches_pil(tmp_path, mode, c, size):
    w, h = size
    a = _noise(h, w, c, seed=w * 31 + h + c)
    p = str(tmp_path / f"t_{mode}_{w}x{h}.png")
    PIL.fromarray(a, mode).save(p, optimize=(w % 2 == 0))
    got = _lib_decode(p)
    assert got.shape == (h, w, c) and got.dtype == np.uint8
    assert np.array_equal(got, a)


@pytest.mark.parametrize("bits", [1, 2, 4, 8])
def test_png_palette_matches_pil(tmp_path, bits):
    rng = np.random.default_rng(bits)
    n = 1 << bits
    idx = rng.integers(0, n, (37, 53), dtype=np.uint8)
    pal = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    im = PIL.fromarray(idx, "P")
    im.putpalette(pal.reshape(-1).tolist())
    p = str(tmp_path / f"pal{bits}.png")
    im.save(p, bits=bits)
    got = _lib_decode(p)
    assert np.array_equal(got, pal[idx])
    # with transparency: four components, alpha from tRNS
    alpha = rng.integers(0, 256, n, dtype=np.uint8)
    p2 = str(tmp_path / f"pal{bits}_trns.png")
    im.save(p2, bits=bits, transparency=bytes(alpha.tolist()))
    got = _lib_decode(p2)
    assert got.shape == (37, 53, 4)
    assert np.array_equal(got[..., :3], pal[idx]) and np.array_equal(got[..., 3], alpha[idx])


def test_png_16bit_keeps_the_high_byte(tmp_path):
    import struct, zlib
    rng = np.random.default_rng(5)
    w, h = 19, 11
    a = rng.integers(0, 65536, (h, w, 3), dtype=np.uint16)
    raw = b"".join(b"\x00" + a[y].astype(">u2").tobytes() for y in range(h))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    p = tmp_path / "rgb16.png"
    p.write_bytes(png)
    assert np.array_equal(_lib_decode(str(p)), (a >> 8).astype(np.uint8))


def test_rejects_what_the_reference_rejects(tmp_path):
    import draw_b200._native as N
    g = str(tmp_path / "grey.png")
    PIL.fromarray(_noise(8, 8, 1, 1)[..., 0], "L").save(g)
    with pytest.raises(N.DrawError):
        _lib_decode(g)                      # 1 component: `unreachable!()` in the reference (scene/mod.rs:187)
    b = str(tmp_path / "x.bmp")
    PIL.fromarray(_noise(8, 8, 3, 2), "RGB").save(b)
    with pytest.raises(N.DrawError):
        _lib_decode(b)                      # formats other than PNG and JPEG are left to the caller's loader
    with pytest.raises(N.DrawError):
        _lib_decode(str(tmp_path / "missing.png"))
    bad = tmp_path / "trunc.png"
    good = tmp_path / "good.png"
    PIL.fromarray(_noise(16, 16, 3, 3), "RGB").save(str(good))
    bad.write_bytes(good.read_bytes()[:60])
    with pytest.raises(N.DrawError):
        _lib_decode(str(bad))


def test_reference_texture_if_present():
    """models/lemur/lemurT.png (the C3 texture) decodes to the bytes PIL gives; skipped where the
    reference tree is not mounted (the GPU box)."""
    p = "/root/reference/models/lemur/lemurT.png"
    if not os.path.exists(p):
        pytest.skip("reference tree not mounted")
    want = np.asarray(PIL.open(p))
    got = _lib_decode(p)
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("c", [3, 4])
def test_png_writer_round_trips(tmp_path, c):
    """draw_image_write_png (stbi_write_png's job, app/mod.rs:362-378): the file decodes — with PIL and with the
    library's own decoder — to exactly the bytes written."""
    import draw_b200.api as api
    a = _noise(45, 67, c, seed=9 + c)
    a[10:30, 5:60] = a[10, 5]  # a constant region, like a rendered frame's background
    p = str(tmp_path / f"out{c}.png")
    api.write_png(p, a)
    assert np.array_equal(np.asarray(PIL.open(p)), a)
    assert np.array_equal(api.load_image(p), a)


def test_png_decoder_survives_damaged_files(tmp_path):
    """Truncated and randomly damaged PNG files: an image of the announced shape or an error, never a crash."""
    import draw_b200._native as N
    rng = np.random.default_rng(4)
    good = tmp_path / "good.png"
    PIL.fromarray(_noise(24, 40, 4, 5), "RGBA").save(str(good))
    data = good.read_bytes()
    n_err = 0
    for trial in range(200):
        bad = bytearray(data)
        if trial % 3 == 0:
            bad = bad[:int(rng.integers(1, len(bad)))]
        else:
            for _ in range(int(rng.integers(1, 5))):
                bad[int(rng.integers(8, len(bad)))] = int(rng.integers(0, 256))
        p = tmp_path / "bad.png"
        p.write_bytes(bytes(bad))
        try:
            a = _lib_decode(str(p))
            assert a.ndim == 3 and a.shape[2] in (3, 4)
        except N.DrawError:
            n_err += 1
    assert n_err > 0
