// Compiles libdvpari.so (hand-written sm_100a kernels + C ABI) and links it.
// The CUDA sources live in ../dv-pari_b200/csrc; the Makefile there runs
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 ...
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let pkg = root.join("dv-pari_b200");
    let status = Command::new("make")
        .arg("-C")
        .arg(&pkg)
        .args(["-j8", "libdvpari.so"])
        .status()
        .expect("make (nvcc) is required to build libdvpari.so");
    assert!(status.success(), "building libdvpari.so failed");
    println!("cargo:rustc-link-search=native={}", pkg.display());
    println!("cargo:rustc-link-lib=dylib=dvpari");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", pkg.display());
    println!("cargo:rerun-if-changed={}", pkg.join("csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include/dvpari.h").display());
}
