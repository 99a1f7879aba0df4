// Compiles the CUDA renderer (one translation unit) with nvcc for sm_100a and links it.
// PHONIC_B200_CSRC may point at phonic_b200/csrc of this repository (default: ../phonic_b200/csrc).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = env::var("PHONIC_B200_CSRC")
        .map(PathBuf::from)
        .unwrap_or_else(|_| manifest.join("../phonic_b200/csrc"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libphonic_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    // -fmad=false: phonic never fuses a*b+c; every f32/f64 operation stays separately rounded (bit-exact voice path)
    let status = Command::new(&nvcc)
        .args([
            "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
            "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-o",
        ])
        .arg(&lib)
        .arg(csrc.join("renderer.cu"))
        .status()
        .expect("nvcc not found: the B200 renderer has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    for f in ["renderer.cu", "voice.cuh", "skeleton_kernel.cuh", "replay_kernel.cuh", "phase_table.cuh", "mixer_kernel.cuh",
              "effects.cuh", "effects_par.cuh", "sinc_kernel.cuh", "gran.cuh", "hq.cuh", "host_fx.h", "dev_structs.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=phonic_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
}
