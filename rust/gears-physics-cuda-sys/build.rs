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
        .arg("-ldl")
        .status()
        .expect("nvcc not found: gears-physics-cuda has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=gears_physics_cuda");
    println!("cargo:rustc-link-lib=dylib=cudart");
    for f in [
        "gp_api.cu", "gp_kernels.cuh", "gp_common.cuh", "gp_mirror_kernels.cuh", "gp_cellsort_kernels.cuh",
        "gp_pairfind_kernels.cuh", "gp_chain_kernels.cuh", "gp_sweep_kernels.cuh", "gp_repair_kernels.cuh",
        "gp_slab_kernels.cuh", "gp_next_kernels.cuh", "gp_multi.cuh", "gp_multi_impl.cuh", "radix_sort.cuh", "scan.cuh",
    ] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", repo.join("include/gears_physics_cuda.h").display());
}
