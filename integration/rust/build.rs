// build.rs — builds libmosaic_b200.so with nvcc (sm_100a only) and links it into the mosaic-rust crate.
// UNVERIFIED SOURCE: there is no rustc/cargo in the image this repository was developed in.
//
// Usage: copy next to the reference's Cargo.toml, add `build = "build.rs"` to [package], and set
//   MOSAIC_B200_DIR=/path/to/this/repository
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("MOSAIC_B200_DIR").expect("MOSAIC_B200_DIR = checkout of the mosaic-b200 repository"));
    let pkg = root.join("mosaic-art-rust_b200");
    // the Makefile runs: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false ...
    let ok = Command::new("make")
        .arg("-C")
        .arg(&pkg)
        .arg("-j8")
        .arg("libmosaic_b200.so")
        .status()
        .expect("running make")
        .success();
    assert!(ok, "building libmosaic_b200.so failed");
    println!("cargo:rustc-link-search=native={}", pkg.display());
    println!("cargo:rustc-link-lib=dylib=mosaic_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", pkg.display());
    println!("cargo:rerun-if-changed={}", root.join("include/mosaic_b200.h").display());
    println!("cargo:rerun-if-changed={}", pkg.join("csrc").display());
    println!("cargo:rerun-if-env-changed=MOSAIC_B200_DIR");
}
