"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol that
include/draw_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "draw_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(draw_[a-z0-9_]+)\s*\(", text))
    names.discard("draw_image_loader")
    return names


def test_library_builds_and_exports_every_declared_symbol():
    from draw_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 35
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"declared in include/draw_b200.h but not exported: {missing}"


def test_binding_covers_the_header():
    from draw_b200 import _native
    assert set(_native.SIGNATURES) == _declared_symbols()


def test_rust_sys_crate_binds_every_declared_symbol():
    """bindings/draw-b200-sys is shipped as source (no Rust toolchain here): it must at least name every export."""
    src = open(os.path.join(ROOT, "bindings", "draw-b200-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (draw_[a-z0-9_]+)\s*\(", src))
    assert bound == _declared_symbols(), sorted(bound ^ _declared_symbols())
    for f in ("Cargo.toml", "build.rs"):
        assert os.path.exists(os.path.join(ROOT, "bindings", "draw-b200-sys", f))
    assert "DRAW_B200_VERSION: c_int = 200" in src


def test_version_and_error_string():
    from draw_b200 import _native
    L = _native.lib()
    assert L.draw_version() == 200
    assert isinstance(L.draw_last_error(), bytes)
    assert L.draw_tile_size() == 32


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import draw_b200
    with pytest.raises(draw_b200.DrawError) as e:
        draw_b200.Scene(64, 48)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(draw_b200.DrawError):
        draw_b200.Canvas(64, 48)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under draw_b200/ may reference it."""
    pkg = os.path.join(ROOT, "draw_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src and "liboracle" not in src, f


def test_struct_layouts_agree_across_the_bindings(tmp_path):
    """The header compiled as C says how big its structs are and where draw_frame_stats' fields sit; the ctypes mirrors
    (draw_b200/_native.py, draw_b200.api.VERTEX2D) and the Rust sys crate's field lists must agree with it."""
    import subprocess
    from draw_b200 import _native
    from draw_b200.synthetic import _VERTEX2D
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "draw_b200.h"\nint main(void) {\n'
                   '  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(draw_frame_stats), offsetof(draw_frame_stats, mirror_kbytes),\n'
                   '         offsetof(draw_frame_stats, front_phase_ns), offsetof(draw_frame_stats, front_block_ns), sizeof(draw_vertex2d),\n'
                   '         sizeof(draw_rect), sizeof(draw_texture_map));\n  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    S = _native.FrameStats
    assert got == [ctypes.sizeof(S), S.mirror_kbytes.offset, S.front_phase_ns.offset, S.front_block_ns.offset, _VERTEX2D.itemsize,
                   ctypes.sizeof(_native.Rect), ctypes.sizeof(_native.TextureMap)]
    # field order of draw_frame_stats: header == ctypes == Rust
    header = open(os.path.join(ROOT, "include", "draw_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header[header.index("typedef struct draw_frame_stats {"):header.index("} draw_frame_stats;")], flags=re.S)
    c_fields = [n for decl in re.findall(r"uint32_t\s+([^;]+);", body) for n in re.findall(r"([a-z_0-9]+)(?:\[\d+\])?\s*(?:,|$)", decl.strip())]
    assert c_fields == [n for n, _ in S._fields_]
    rust = open(os.path.join(ROOT, "bindings", "draw-b200-sys", "src", "lib.rs")).read()
    rbody = rust[rust.index("pub struct draw_frame_stats {"):]
    rbody = rbody[:rbody.index("}")]
    assert re.findall(r"pub ([a-z_0-9]+):", rbody) == c_fields


def test_bench_counts_the_uniform_upload_it_really_makes(tmp_path):
    """bench.py's h2d_bytes_per_step is frames x sizeof(FrameUniforms): the constant must follow the struct."""
    import subprocess
    import bench
    src = tmp_path / "sz.cpp"
    src.write_text('#include <cstdio>\n#include "device_types.h"\nint main() { std::printf("%zu\\n", sizeof(drawb200::FrameUniforms)); }\n')
    exe = tmp_path / "sz"
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "draw_b200", "csrc"), "-I", "/usr/local/cuda/include", str(src), "-o", str(exe)],
                   check=True)
    assert int(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout) == bench.UNIFORM_BYTES
