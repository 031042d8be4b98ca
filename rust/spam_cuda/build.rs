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
    // every .cu of the directory, like sparse_matrix_b200/build.py (its SOURCES list is checked against the
    // directory by tests/test_host.py); headers only trigger rebuilds
    let mut sources: Vec<PathBuf> = std::fs::read_dir(&csrc).expect("SPAM_CUDA_CSRC is not a directory")
        .filter_map(|e| e.ok().map(|e| e.path())).collect();
    sources.sort();
    for f in &sources {
        match f.extension().and_then(|e| e.to_str()) {
            Some("cu") => { cmd.arg(f); println!("cargo:rerun-if-changed={}", f.display()); }
            Some("cuh") => println!("cargo:rerun-if-changed={}", f.display()),
            _ => {}
        }
    }
    cmd.args(["-lpthread", "-ldl"]);
    assert!(cmd.status().expect("nvcc not found").success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=spam_cuda");
}
