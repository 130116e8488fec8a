// Links libgooey_b200.so (built by `python -c "import __graft_entry__ as g; g.build()"` into libgooey_b200/lib/).
fn main() {
    let dir = std::env::var("GOOEY_B200_LIB_DIR").unwrap_or_else(|_| "../libgooey_b200/lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=gooey_b200");
    println!("cargo:rerun-if-env-changed=GOOEY_B200_LIB_DIR");
}
