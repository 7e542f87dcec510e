//! Compiles the CUDA sources for sm_100a with nvcc and links the result.
//!
//! Flags that matter for bit parity with the Rust reference (which never contracts a*b+c and uses
//! IEEE division / square root): `--fmad=false -prec-div=true -prec-sqrt=true -ftz=false`.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    // <repo>/rust/gears-physics-cuda-sys -> <repo>
    let repo = manifest.parent().unwrap().parent().unwrap().to_path_buf();
    let csrc = repo.join("gears_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let so = out.join("libgears_physics_cuda.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(&nvcc)
        .args([
            "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "--fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
            "-Xcompiler", "-fPIC", "-shared", "-o",
        ])
        .arg(&so)
        .arg(csrc.join("gp_api.cu"))
        .arg("-ldl") // NCCL is bound with dlopen at run time (gp_multi_impl.cuh): single-GPU users need no NCCL
        .status()
        .expect("nvcc not found: gears-physics-cuda has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=gears_physics_cuda");
    println!("cargo:rustc-link-lib=dylib=cudart");
    // every source under csrc/ (gp_api.cu includes all the .cuh files)
    for entry in std::fs::read_dir(&csrc).expect("gears_b200/csrc missing") {
        let path = entry.unwrap().path();
        if matches!(path.extension().and_then(|e| e.to_str()), Some("cu") | Some("cuh")) {
            println!("cargo:rerun-if-changed={}", path.display());
        }
    }
    println!("cargo:rerun-if-changed={}", repo.join("include/gears_physics_cuda.h").display());
}
