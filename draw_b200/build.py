"""Builds libdraw_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m draw_b200.build [--force] [--verbose]

Device code: -gencode arch=compute_100a,code=sm_100a -fmad=false (the reference never fuses
a*b+c; all parity-critical arithmetic additionally uses the __f*_rn intrinsics), -lineinfo so
ncu's source page maps to kernels.cu.  Host code: -ffp-contract=off for the per-frame uniforms.
The CUDA runtime is linked statically, so the .so only needs the driver on the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdraw_b200.so")
SOURCES = ["k_front.cu", "k_raster.cu", "k_tile.cu", "k_sort.cu", "k_sync.cu", "k_overlay.cu", "k_mirror.cu", "scene.cpp", "obj_loader.cpp", "image_decode.cpp", "jpeg_decode.cpp", "jpeg_encode.cpp"]
HEADERS = ["device_types.h", "device_math.cuh", "shading.cuh", "host_math.hpp", os.path.join("..", "..", "include", "draw_b200.h")]

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall", "-cudart", "static"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; libdraw_b200.so cannot be built")
    return p


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out build an experimental variant (e.g. another tile size) next to the default library;
    select it at run time with DRAW_B200_LIB=<path>."""
    lib = out or LIB
    if not force and not defines and up_to_date():
        return lib
    objs = []
    build_dir = os.path.join(HERE, "build", os.path.basename(lib))
    os.makedirs(build_dir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(build_dir, src + ".o")
        cmd = [nvcc_path(), *NVCC_FLAGS, *defines, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose and src.endswith(".cu"):
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    cmd = [nvcc_path(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", lib, *objs, "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return lib


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=outs[0] if outs else None))
