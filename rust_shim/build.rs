// build.rs -- the reference's build script (build.rs:1-29 compiles the WESL preview shaders) plus the CUDA backend.
// UNTESTED IN THIS REPO (no cargo here).  `PTB_SRC` points at this repository's path_tracer_rust_b200/csrc.
use std::{env, path::PathBuf, process::Command};

fn main() {
    // ... existing WESL shader compilation of the reference stays here unchanged ...

    let src = PathBuf::from(env::var("PTB_SRC").expect("set PTB_SRC to path_tracer_rust_b200/csrc"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libptb.so");
    // --fmad=false is a parity requirement: ray generation / intersection arithmetic must not be fused
    let flags = ["-arch=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false", "-prec-div=true", "-prec-sqrt=true",
                 "-ftz=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-shared"];
    let status = Command::new("nvcc")
        .args(flags)
        .arg("-o").arg(&lib)
        .args(["pt_kernels.cu", "pt_wavefront.cu", "pt_bvh_build.cu", "pt_api.cu"].iter().map(|f| src.join(f)))
        .arg("-x").arg("cu").arg(src.join("scene_io.cpp"))
        .status().expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=ptb");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
    for f in ["pt_kernels.cu", "pt_bvh_build.cu", "pt_api.cu", "scene_io.cpp", "pt_device.cuh", "pt_bvh.cuh"] {
        println!("cargo:rerun-if-changed={}", src.join(f).display());
    }
}
