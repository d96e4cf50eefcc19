// Links libdeflate_b200.so.  DEFLATE_B200_LIB_DIR = directory holding the library built by
// `python -c "import __graft_entry__ as g; g.build()"` (deflate-rs_b200/ in this repository).
fn main() {
    if let Ok(dir) = std::env::var("DEFLATE_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=deflate_b200");
    println!("cargo:rerun-if-env-changed=DEFLATE_B200_LIB_DIR");
}
