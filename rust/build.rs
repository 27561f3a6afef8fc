// build.rs -- compiles the CUDA library for sm_100a with nvcc and links it (replaces the reference's
// `.cargo/config.toml` `+avx2` flag: the hot path no longer needs AVX2, it needs a B200).
// NOT EXERCISED in the build image (no Rust toolchain); the same nvcc command is what
// feature_detector_fast_b200/csrc/Makefile runs.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../feature_detector_fast_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".to_string());
    let lib = out.join("libfdf_cuda.so");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"])
        .args(["-Xcompiler", "-fPIC", "--shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("fdf_kernels.cu"))
        .arg(csrc.join("fdf_capi.cu"))
        .status()
        .expect("failed to run nvcc");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=fdf_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
    for f in ["fdf_kernels.cu", "fdf_capi.cu", "fdf_core.cuh", "fdf_strip.cuh", "fdf_kernels.cuh", "fdf_synth.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", manifest.join("../include/fdf.h").display());
}
