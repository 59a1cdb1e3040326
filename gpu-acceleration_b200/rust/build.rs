// build.rs for the `cuda` feature of mopro-msm (replaces the Metal shader build,
// mopro-msm/build.rs:12-173: xcrun metal -> .air -> .metallib).  SOURCE ONLY (no cargo here).
use std::{env, path::PathBuf, process::Command};

fn main() {
    if env::var("CARGO_FEATURE_CUDA").is_err() {
        return;
    }
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let src = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("cuda/csrc/b200msm.cu");
    let lib = out.join("libb200msm.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
               "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(&src)
        .arg("-lcudart")
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=b200msm");
    println!("cargo:rerun-if-changed={}", src.display());
}
