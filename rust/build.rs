// build.rs of the drop-in: compiles the CUDA sources of the hot path (this repository's rustybam_b200/csrc/*.cu, vendored as
// cuda/ in the rustybam tree) with nvcc for sm_100a and links the result.  Added next to the reference's existing build.rs
// logic (the clap `include!`).  NOT compiled in this repository's image (no cargo / rustc here): kept as the file a maintainer
// adds; the runnable host of this repository is the C++ equivalent, rustybam_b200/host/rb_main.cpp.
// build.rs (added next to the existing clap include!)
fn main() {
    let out = std::path::PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let lib = out.join("librbcuda.so");
    let status = std::process::Command::new("nvcc")
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .args(["cuda/rb_kernels.cu", "cuda/trim_kernels.cu", "cuda/inflate_kernels.cu", "cuda/rbcuda.cu"])   // == rustybam_b200/csrc/*.cu of this repo
        .status().expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=rbcuda");
    println!("cargo:rerun-if-changed=cuda");
}
