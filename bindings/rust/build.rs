// build.rs — compiles the CUDA side for sm_100a with nvcc and links it statically.
// (What `make -C hijiki_b200/csrc` does)
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("HIJIKI_B200_CSRC").unwrap_or_else(|_| "hijiki_b200/csrc".into()));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());

    // device side: one translation unit (kernels + C ABI)
    let ctx_o = out.join("context.o");
    let ok = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
               "-Xcompiler", "-fPIC", "-c"])
        .arg(csrc.join("device/context.cu"))
        .arg("-o")
        .arg(&ctx_o)
        .status()
        .expect("nvcc not found")
        .success();
    assert!(ok, "nvcc failed");

    // host side: plain C++ (wide-BVH builder, pass planner; loaders are optional for a Rust host)
    let mut objs = vec![ctx_o];
    for f in ["cwbvh_build", "pass_plan", "scene_compile", "obj_loader", "bvh2_sah", "synthetic", "host_api"] {
        let o = out.join(format!("{f}.o"));
        let ok = Command::new("g++")
            .args(["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-c"])
            .arg(csrc.join(format!("host/{f}.cpp")))
            .arg("-o")
            .arg(&o)
            .status()
            .expect("g++ not found")
            .success();
        assert!(ok, "g++ failed on {f}.cpp");
        objs.push(o);
    }
    let lib = out.join("libhijiki_b200.a");
    let _ = std::fs::remove_file(&lib);
    assert!(Command::new("ar").arg("rcs").arg(&lib).args(&objs).status().unwrap().success());

    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=hijiki_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=static=cudart_static");
    for l in ["stdc++", "dl", "pthread", "rt"] {
        println!("cargo:rustc-link-lib={l}");
    }
    println!("cargo:rerun-if-changed={}", csrc.display());
}
