// Links libngs_cuda.so (built by `make -C ngs_b200/csrc` in the engine's repository).
//   NGS_CUDA_LIB_DIR      directory that holds libngs_cuda.so (required)
//   NGS_CUDA_INCLUDE_DIR  directory that holds ngs_cuda.h (only used to re-run the build when the header changes)
fn main() {
    let dir = std::env::var("NGS_CUDA_LIB_DIR").expect("set NGS_CUDA_LIB_DIR to the directory that holds libngs_cuda.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ngs_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=NGS_CUDA_LIB_DIR");
    if let Ok(inc) = std::env::var("NGS_CUDA_INCLUDE_DIR") {
        println!("cargo:rerun-if-changed={inc}/ngs_cuda.h");
    }
}
