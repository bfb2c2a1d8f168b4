"""draw_image_load on JPEG files (draw_b200/csrc/jpeg_decode.cpp: T.81 entropy decoding + stb_image's IDCT, upsampling
and colour arithmetic; TextureMap::load_from_file, scene/mod.rs:174-202).  stb_image is not in this image, so the
checker is libjpeg through Pillow: the same coefficients go through a differently rounded IDCT / colour conversion, so
texels agree to within a few levels — a wrong Huffman / progressive / restart / sampling path would be off by far more.
Host code only: runs without a GPU."""
import os

import numpy as np
import pytest

PIL_Image = pytest.importorskip("PIL.Image")

AIRPLANE = "/root/reference/models/airplane/11804_Airplane_diff.jpg"


def _picture(h=203, w=317):
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([np.sin(xx * 0.05) * 127 + 128, np.cos(yy * 0.07) * 127 + 128, (xx + yy) % 256], -1).astype(np.uint8)
    noise = img[h // 4:h // 2, w // 5:2 * w // 3]
    noise[...] = rng.integers(0, 256, noise.shape)  # noise: every AC coefficient in use
    return img


def _compare(path):
    import draw_b200
    got = draw_b200.load_image(path)
    want = np.asarray(PIL_Image.open(path).convert("RGB"))
    assert got.shape == want.shape and got.dtype == np.uint8
    return np.abs(got.astype(int) - want.astype(int))


@pytest.mark.parametrize("progressive", [False, True])
@pytest.mark.parametrize("restart_rows", [0, 2])
@pytest.mark.parametrize("subsampling,name", [(0, "4:4:4"), (1, "4:2:2"), (2, "4:2:0")])
def test_decoder_against_libjpeg(tmp_path, subsampling, name, restart_rows, progressive):
    path = str(tmp_path / "t.jpg")
    kw = dict(quality=90, subsampling=subsampling, progressive=progressive)
    if restart_rows:
        kw["restart_marker_rows"] = restart_rows
    PIL_Image.fromarray(_picture()).save(path, **kw)
    d = _compare(path)
    if subsampling == 1:
        # stb_image's h2v1 filter weights the last column's two chroma samples the other way round than libjpeg's
        # (out[2w-2] = (3 * in[w-2] + in[w-1] + 2) >> 2); restated as is, so that column is left out of the bound
        d = d[:, :-2]
    assert d.max() <= 6, f"{name}: max difference {d.max()}"
    assert d.mean() < 0.25 and (d > 1).mean() < 0.06, f"{name}: mean {d.mean()}, share above 1: {(d > 1).mean()}"


@pytest.mark.parametrize("size", [(1, 1), (8, 8), (17, 9), (16, 33)])
def test_odd_sizes(tmp_path, size):
    w, h = size
    path = str(tmp_path / "s.jpg")
    PIL_Image.fromarray(_picture(h, w)).save(path, quality=95, subsampling=2)
    d = _compare(path)
    interior = d[:, :-1] if w > 1 else d
    assert interior.max() <= 8


@pytest.mark.skipif(not os.path.exists(AIRPLANE), reason="the reference's assets are only in the build container")
def test_reference_asset_progressive_444():
    """models/airplane/11804_Airplane_diff.jpg: 1024x1024, progressive, no chroma subsampling."""
    d = _compare(AIRPLANE)
    assert d.shape == (1024, 1024, 3)
    assert d.max() <= 3 and d.mean() < 0.01


def test_known_answer_flat_block(tmp_path):
    """A flat grey image has only DC coefficients: the IDCT's DC path, (dc * 4) >> ... per stb_image, returns exactly
    the level libjpeg returns, and the colour conversion of a neutral pixel is the identity."""
    import draw_b200
    path = str(tmp_path / "flat.jpg")
    PIL_Image.fromarray(np.full((16, 16, 3), 200, np.uint8)).save(path, quality=100, subsampling=0)
    got = draw_b200.load_image(path)
    assert (got == 200).all()


def test_rejected_files(tmp_path):
    import draw_b200
    grey = str(tmp_path / "grey.jpg")
    PIL_Image.fromarray(np.full((8, 8), 90, np.uint8)).save(grey)
    with pytest.raises(draw_b200.DrawError, match="greyscale"):
        draw_b200.load_image(grey)  # one native channel: the reference's `unreachable!` (scene/mod.rs:187)
    cut = str(tmp_path / "cut.jpg")
    PIL_Image.fromarray(_picture()).save(cut, quality=90)
    data = open(cut, "rb").read()
    open(cut, "wb").write(data[:200])
    with pytest.raises(draw_b200.DrawError):
        draw_b200.load_image(cut)


# ---- the writer (draw_image_write_jpg, stbi_write_jpg's scheme) -------------------------------------------------------

def _segments(path, marker):
    data, i, out = open(path, "rb").read(), 2, []
    while i < len(data):
        assert data[i] == 0xFF
        m, n = data[i + 1], data[i + 2] << 8 | data[i + 3]
        if m == marker:
            out.append(data[i + 4:i + 2 + n])
        if m == 0xDA:
            break
        i += 2 + n
    return out


@pytest.mark.parametrize("size", [(317, 203), (8, 8), (1, 1), (33, 7)])
def test_writer_round_trip(tmp_path, size):
    """Quality 100 = every quantiser 1: both libjpeg and the library's own decoder get the source back to within the
    rounding of the colour transform and the DCT; the alpha byte of a four-component source is ignored."""
    import draw_b200
    w, h = size
    src = np.concatenate([_picture(h, w), np.full((h, w, 1), 77, np.uint8)], -1)
    path = str(tmp_path / "w.jpg")
    draw_b200.write_jpg(path, src, quality=w * 4 if w * 4 > 90 else 100)  # the reference passes width * 4 (app/mod.rs:371-377)
    im = PIL_Image.open(path)
    assert im.size == (w, h) and im.mode == "RGB"
    assert all(v == 1 for table in im.quantization.values() for v in table)
    for decoded in (np.asarray(im), draw_b200.load_image(path)):
        d = np.abs(decoded.astype(int) - src[..., :3].astype(int))
        assert d.max() <= 4 and d.mean() < 0.6


def test_writer_uses_the_annex_k_huffman_tables(tmp_path):
    """stb_image_write ships the T.81 Annex K tables; so does libjpeg when it does not optimise: the DHT payloads agree."""
    import draw_b200
    src = _picture(40, 40)
    ours, theirs = str(tmp_path / "a.jpg"), str(tmp_path / "b.jpg")
    draw_b200.write_jpg(ours, src, quality=100)
    PIL_Image.fromarray(src).save(theirs, quality=100, subsampling=0, optimize=False)

    def tables(path):
        out = {}
        for seg in _segments(path, 0xC4):
            j = 0
            while j < len(seg):
                n = sum(seg[j + 1:j + 17])
                out[seg[j]] = seg[j + 1:j + 17 + n]
                j += 17 + n
        return out
    a, b = tables(ours), tables(theirs)
    assert sorted(a) == [0x00, 0x01, 0x10, 0x11] and a == b


def test_writer_quality_argument(tmp_path):
    import draw_b200
    src = _picture(16, 16)
    with pytest.raises(draw_b200.DrawError, match="4:2:0"):
        draw_b200.write_jpg(str(tmp_path / "q.jpg"), src, quality=90)  # stbi_write_jpg would subsample; the export never does
    with pytest.raises(draw_b200.DrawError):
        draw_b200.write_jpg(str(tmp_path / "q.jpg"), src, quality=0)   # 0 means 90
    draw_b200.write_jpg(str(tmp_path / "q.jpg"), src, quality=95)
    q = PIL_Image.open(str(tmp_path / "q.jpg")).quantization
    assert q[0][0] == (16 * 10 + 50) // 100 and max(q[1]) == (99 * 10 + 50) // 100  # the IJG scale: 200 - 2 * 95 = 10


def test_decoder_survives_damaged_files(tmp_path):
    """Truncations and random byte damage anywhere in baseline and progressive files: the decoder either returns an image
    of the announced size or fails with an error — it does not crash, hang or read outside the file."""
    import draw_b200
    rng = np.random.default_rng(3)
    src = _picture(40, 56)
    n_ok = n_err = 0
    for progressive in (False, True):
        good = str(tmp_path / "good.jpg")
        PIL_Image.fromarray(src).save(good, quality=85, subsampling=2, progressive=progressive, restart_marker_rows=1)
        data = bytearray(open(good, "rb").read())
        for trial in range(150):
            bad = bytearray(data)
            if trial % 3 == 0:
                bad = bad[:int(rng.integers(2, len(bad)))]
            else:
                for _ in range(int(rng.integers(1, 6))):
                    bad[int(rng.integers(2, len(bad)))] = int(rng.integers(0, 256))
            p = str(tmp_path / "bad.jpg")
            open(p, "wb").write(bytes(bad))
            try:
                a = draw_b200.load_image(p)
                assert a.ndim == 3 and a.shape[2] == 3 and a.dtype == np.uint8
                n_ok += 1
            except draw_b200.DrawError:
                n_err += 1
    assert n_ok + n_err == 300 and n_err > 0
