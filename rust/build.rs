// Link against librtb200.so built by `python -c 'import __graft_entry__ as g; g.build()'`.
fn main() {
    let dir = std::env::var("RTB200_LIB_DIR").unwrap_or_else(|_| "../surely_raytracing_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rtb200");
    println!("cargo:rerun-if-env-changed=RTB200_LIB_DIR");
}
