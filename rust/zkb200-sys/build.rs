// Link against libzkb200.so. ZKB200_LIB_DIR points at the directory that holds it (zk-circuits_b200/ in this repository).
fn main() {
    if let Ok(dir) = std::env::var("ZKB200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=zkb200");
    println!("cargo:rerun-if-env-changed=ZKB200_LIB_DIR");
}
