// build.rs of the reference crate once src/gpu.rs is added: where libstroemung_b200.so lives.
// STROEMUNG_B200_LIB_DIR = the directory holding the library built by `make -C stroemung_b200/csrc`.
fn main() {
    if let Ok(dir) = std::env::var("STROEMUNG_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=stroemung_b200");
    println!("cargo:rerun-if-env-changed=STROEMUNG_B200_LIB_DIR");
}
