// Links libomb200.so (built by `python -m openmeters_b200.build`; nvcc, sm_100a, static cudart).
// OMB200_LIB_DIR overrides the in-tree location.  src/lib.rs is generated from include/omb200.h by tools/gen_rust_sys.py,
// so no bindgen / libclang is needed at build time.
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("OMB200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../openmeters_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=omb200");
    println!("cargo:rerun-if-env-changed=OMB200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/omb200.h");
}
