// Links libbpperm_cuda.so (built by `python -c "import __graft_entry__ as g; g.build()"`).
// BPPERM_LIB_DIR = directory holding the shared library (default: ../../bulletproof-perm_b200).
fn main() {
    let dir = std::env::var("BPPERM_LIB_DIR").unwrap_or_else(|_| {
        format!("{}/../../bulletproof-perm_b200", std::env::var("CARGO_MANIFEST_DIR").unwrap())
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=bpperm_cuda");
    println!("cargo:rerun-if-env-changed=BPPERM_LIB_DIR");
}
