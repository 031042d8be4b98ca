// build.rs — compiles the .cu sources for sm_100a with nvcc into libspam_cuda.so and links it.
// Mirrors sparse_matrix_b200/build.py (the recipe the Python side uses).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    // SPAM_CUDA_CSRC points at sparse_matrix_b200/csrc of this repository
    let csrc = PathBuf::from(env::var("SPAM_CUDA_CSRC").expect("set SPAM_CUDA_CSRC to .../sparse_matrix_b200/csrc"));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let so = out.join("libspam_cuda.so");
    let mut cmd = Command::new(nvcc);
    cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&so);
    for f in ["api.cu", "spgemm.cu", "scan.cu", "convert.cu", "spmv.cu", "dok.cu"] {
        cmd.arg(csrc.join(f));
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    for f in ["common.cuh", "merge.cuh", "rowhash.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    assert!(cmd.status().expect("nvcc not found").success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=spam_cuda");
}
