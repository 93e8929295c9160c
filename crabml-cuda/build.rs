// Locates libcrabml_cuda.so (built with `python -m crabml_b200.build`: nvcc only, cudart linked statically, so the shim has
// no other native dependency) and tells rustc to link it.
//   CRABML_CUDA_LIB_DIR   directory that holds libcrabml_cuda.so (default: <repo>/crabml_b200/lib next to this crate)
use std::env;
use std::path::PathBuf;

fn main() {
    println!("cargo:rerun-if-env-changed=CRABML_CUDA_LIB_DIR");
    println!("cargo:rerun-if-changed=../include/crabml_cuda.h");
    let dir = match env::var("CRABML_CUDA_LIB_DIR") {
        Ok(d) => PathBuf::from(d),
        Err(_) => {
            let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").expect("CARGO_MANIFEST_DIR"));
            manifest.join("..").join("crabml_b200").join("lib")
        }
    };
    if !dir.join("libcrabml_cuda.so").exists() {
        panic!(
            "libcrabml_cuda.so not found in {}: build it with `python -m crabml_b200.build` or set CRABML_CUDA_LIB_DIR",
            dir.display()
        );
    }
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=crabml_cuda");
    // let binaries find the library without LD_LIBRARY_PATH
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
