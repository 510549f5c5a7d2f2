// Links the prebuilt liborc_b200.so (built by `make -C orc_b200/csrc`); bindgen is not needed for ~45 plain-C symbols.
fn main() {
    let dir = std::env::var("ORC_B200_LIB_DIR").unwrap_or_else(|_| "../../orc_b200/lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=orc_b200");
    println!("cargo:rerun-if-changed=../../include/orc_b200.h");
}
