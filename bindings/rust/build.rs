// Links libntgpu.so (built by `python needletail_b200/build.py`: nvcc, sm_100a).  NTGPU_LIB_DIR = directory of the .so
// (default: ../../needletail_b200).  With `--features bindgen` the declarations are regenerated from include/ntgpu.h.
use std::env;
use std::path::PathBuf;

fn main() {
    let here = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let lib_dir = env::var("NTGPU_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| here.join("../../needletail_b200"));
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=ntgpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    println!("cargo:rerun-if-changed=../../include/ntgpu.h");
    #[cfg(feature = "bindgen")]
    {
        let out = PathBuf::from(env::var("OUT_DIR").unwrap()).join("ntgpu.rs");
        bindgen::Builder::default()
            .header(here.join("../../include/ntgpu.h").to_str().unwrap())
            .allowlist_function("ntg_.*")
            .allowlist_type("ntg_.*")
            .allowlist_var("NTG_.*")
            .generate()
            .expect("bindgen on include/ntgpu.h")
            .write_to_file(out)
            .unwrap();
    }
}
