// build.rs -- builds libnumrs_b200 (nvcc, sm_100a only) via the repo's Makefile and links it.
// The `cc` crate is used only to locate a host C++ compiler for nvcc's -ccbin.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("numrs_b200").join("csrc");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let host_cxx = cc::Build::new().cpp(true).get_compiler().path().to_path_buf();
    let status = Command::new("make")
        .arg("-C").arg(&csrc)
        .arg(format!("NVCC={}", nvcc))
        .arg(format!("HOSTCXX={}", host_cxx.display()))
        .arg("-j8")
        .status()
        .expect("failed to run make for libnumrs_b200");
    assert!(status.success(), "libnumrs_b200 build failed (needs nvcc >= 12.9 with sm_100a)");
    println!("cargo:rustc-link-search=native={}", root.join("numrs_b200").display());
    println!("cargo:rustc-link-lib=dylib=numrs_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include").join("numrs_b200.h").display());
}
