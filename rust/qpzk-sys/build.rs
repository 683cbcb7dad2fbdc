// cc-built .cu, sm_100a only: no Triton, no multi-backend dispatch, no CPU fallback.
// The kernels are a unity build (csrc/qpzk.cu includes every .cuh), so one translation unit suffices.
use std::path::PathBuf;

fn main() {
    let root = PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("qp-zk-circuits-rm_b200/csrc");
    cc::Build::new()
        .cuda(true)
        .cudart("shared")
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .flag("-O3")
        .flag("-lineinfo")
        .flag("-std=c++17")
        .include(root.join("include"))
        .file(csrc.join("qpzk.cu"))
        .compile("qpzk");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/qpzk.h").display());
}
