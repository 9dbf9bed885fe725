// Link against nnnoiseless_b200/lib/libnnnoiseless_b200.so (built by `python -m nnnoiseless_b200.build`).
fn main() {
    let root = std::env::var("NNNOISELESS_B200_LIB_DIR")
        .unwrap_or_else(|_| format!("{}/../../nnnoiseless_b200/lib", env!("CARGO_MANIFEST_DIR")));
    println!("cargo:rustc-link-search=native={}", root);
    println!("cargo:rustc-link-lib=dylib=nnnoiseless_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", root);
}
