// Links libblbm.so (built by `python -m lbm_b200.build`, i.e. lbm_b200/csrc/Makefile).
// BLBM_LIB_DIR overrides the default in-tree location.
fn main() {
    let dir = std::env::var("BLBM_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../../lbm_b200", manifest)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=blbm");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=BLBM_LIB_DIR");
}
