// Links against openwurli_b200/lib/libowgpu.so (built by `make -C openwurli_b200/csrc`).
// Override the search path with OWGPU_LIB_DIR.
fn main() {
    let dir = std::env::var("OWGPU_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{manifest}/../../../openwurli_b200/lib")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=owgpu");
    println!("cargo:rerun-if-env-changed=OWGPU_LIB_DIR");
}
