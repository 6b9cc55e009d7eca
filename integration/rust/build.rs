// Link against the prebuilt C-ABI library. CUBECL_B200_LIB_DIR = directory holding libcubecl_b200.so
// (repo: cubecl_b200/lib after `python -m cubecl_b200.build`).
fn main() {
    let dir = std::env::var("CUBECL_B200_LIB_DIR").unwrap_or_else(|_| "../../cubecl_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=cubecl_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=CUBECL_B200_LIB_DIR");
}
