// Links against the prebuilt libkmerust_gpu.so (built by `python -m krust_b200.build`).
fn main() {
    let dir = std::env::var("KMERUST_GPU_LIB_DIR").unwrap_or_else(|_| "../../../krust_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=kmerust_gpu");
    println!("cargo:rerun-if-env-changed=KMERUST_GPU_LIB_DIR");
}
