// Link against libgadcu.so built by `python -m gad_b200.build` (GADCU_LIB_DIR overrides the location).
fn main() {
    let dir = std::env::var("GADCU_LIB_DIR").unwrap_or_else(|_| "../gad_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=gadcu");
}
