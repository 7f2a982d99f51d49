// Links the in-tree shared library built by `python -m mapad_b200.build` (nvcc, sm_100a).  MAPAD_GPU_LIB_DIR overrides the
// directory that holds libmapad_gpu.so.
fn main() {
    let dir = std::env::var("MAPAD_GPU_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").expect("set by cargo");
        format!("{manifest}/../../../mapad_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=mapad_gpu");
    println!("cargo:rerun-if-env-changed=MAPAD_GPU_LIB_DIR");
}
