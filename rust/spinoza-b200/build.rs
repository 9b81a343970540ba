// Links libspinoza_b200.so (built by `python -m spinoza_b200._build` into spinoza_b200/lib/).
// Override the location with SPINOZA_B200_LIB_DIR.
fn main() {
    let dir = std::env::var("SPINOZA_B200_LIB_DIR")
        .unwrap_or_else(|_| format!("{}/../../spinoza_b200/lib", env!("CARGO_MANIFEST_DIR")));
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=spinoza_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=SPINOZA_B200_LIB_DIR");
}
