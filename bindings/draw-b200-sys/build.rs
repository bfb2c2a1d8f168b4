// Links libdraw_b200.so (built by `python -m draw_b200.build`: nvcc, sm_100a).  The library links the CUDA
// runtime statically and needs only the NVIDIA driver at run time.
//   DRAW_B200_LIB_DIR   directory that holds libdraw_b200.so (default: ../../draw_b200 relative to this crate)
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("DRAW_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../draw_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=draw_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=DRAW_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/draw_b200.h");
}
