// Links the prebuilt libeon_kzg.so (nvcc, sm_100a; `make -C plonky3_eon_b200/csrc`).
fn main() {
    let dir = std::env::var("EON_KZG_LIB_DIR").unwrap_or_else(|_| {
        // default: the in-tree build output next to this repository's Python package
        format!("{}/../../plonky3_eon_b200", env!("CARGO_MANIFEST_DIR"))
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=eon_kzg");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=EON_KZG_LIB_DIR");
}
