// Links the prebuilt C-ABI library.  HELIO_VOXEL_CUDA_LIB_DIR points at the directory holding
// libhelio_voxel_cuda.so (built by `make -C helio_b200/csrc`, nvcc sm_100a).
fn main() {
    let dir = std::env::var("HELIO_VOXEL_CUDA_LIB_DIR").unwrap_or_else(|_| "/usr/local/lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=helio_voxel_cuda");
    println!("cargo:rerun-if-env-changed=HELIO_VOXEL_CUDA_LIB_DIR");
}
