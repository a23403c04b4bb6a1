// Links libxdtts_b200.so (include/xdtts_b200.h).  XDTTS_B200_LIB_DIR = directory that holds the library
// (xd-tts_b200/xdtts_b200/_lib after `python __graft_entry__.py`).
fn main() {
    println!("cargo:rerun-if-env-changed=XDTTS_B200_LIB_DIR");
    let dir = std::env::var("XDTTS_B200_LIB_DIR").expect("set XDTTS_B200_LIB_DIR to the directory of libxdtts_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=xdtts_b200");
}
