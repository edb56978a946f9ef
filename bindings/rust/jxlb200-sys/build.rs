// Links against the in-tree libjxlb200.so: JXLB200_LIB_DIR = <repo>/jpeg-xl-lossy-image-compression-thesis_b200
fn main() {
    println!("cargo:rerun-if-env-changed=JXLB200_LIB_DIR");
    if let Ok(dir) = std::env::var("JXLB200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=jxlb200");
}
