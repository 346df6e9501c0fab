// Link against libsolb.so: SOLB_LIB_DIR points at <repo>/sol_rs_b200 (where `make -C sol_rs_b200/csrc` leaves it).
fn main() {
    let dir = std::env::var("SOLB_LIB_DIR").unwrap_or_else(|_| "../../sol_rs_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=solb");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=SOLB_LIB_DIR");
}
