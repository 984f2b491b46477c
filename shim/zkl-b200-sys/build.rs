// Links libzkl_b200.so, built by `python __graft_entry__.py` (nvcc, sm_100a) in the zk-lisp_b200 checkout.
fn main() {
    let root = std::env::var("ZKL_B200_ROOT").expect("ZKL_B200_ROOT = path to the zk-lisp_b200 checkout");
    println!("cargo:rustc-link-search=native={root}/zk-lisp_b200/csrc");
    println!("cargo:rustc-link-lib=dylib=zkl_b200");
    println!("cargo:rerun-if-env-changed=ZKL_B200_ROOT");
    println!("cargo:rerun-if-changed={root}/include/zkl_b200.h");
}
