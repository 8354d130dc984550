"""Builds librestir_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["csrc/restir_kernels.cu", "csrc/restir_trace.cu", "csrc/restir_halo.cu", "csrc/restir_gbuffer.cu", "csrc/restir_bvh_build.cu", "csrc/restir_wide_build.cu", "csrc/restir_generic.cu", "csrc/restir_capi.cu", "csrc/traversal_image.cpp", "csrc/wide_image.cpp", "host/scene_build.cpp"]
DRIVER_SOURCES = ["host/restir_driver.cpp", "host/passes.hpp", "host/capture.hpp", "../include/restir_capture.h"]
HEADERS = ["csrc/restir_math.cuh", "csrc/restir_pixel.cuh", "csrc/restir_device.cuh", "csrc/restir_kernels.h", "csrc/restir_trace.cuh", "csrc/restir_wide.cuh", "csrc/traversal_image.h", "csrc/wide_image.h", "../include/restir_b200.h",
           "../include/restir_layouts.h", "host/passes.hpp"]
OUT = os.path.join(HERE, "librestir_b200.so")
DRIVER = os.path.join(HERE, "restir_driver")

# -fmad=false: no FMA contraction (arithmetic policy P1); division and sqrt stay IEEE (-prec-div/-prec-sqrt
# default true), denormals are kept (-ftz default false).
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-shared",
]


def nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build():
    if not os.path.exists(OUT) or not os.path.exists(DRIVER):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.exists(os.path.join(HERE, s)) and os.path.getmtime(os.path.join(HERE, s)) > t for s in SOURCES + HEADERS + DRIVER_SOURCES + ["build.py"])


def build_variant(out, defines, verbose=False):
    """Experiment builds (same sources, extra -D flags) loaded through RESTIR_B200_LIB; not part of build()."""
    cmd = [nvcc(), "-ccbin", "/usr/bin/g++"] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(HERE, s) for s in SOURCES]
    subprocess.check_call(cmd, cwd=HERE)
    return out


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(HERE, s) for s in SOURCES]
    ccbin = "/usr/bin/g++"
    if os.path.exists(ccbin):
        cmd[1:1] = ["-ccbin", ccbin]
    subprocess.check_call(cmd, cwd=HERE)
    # the C++ host driver (mirror of the reference's pass classes + App main loop) links against the C ABI
    drv = [nvcc()] + (["-ccbin", ccbin] if os.path.exists(ccbin) else []) + [
        "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler", "-Wall,-ffp-contract=off", "-o", DRIVER, os.path.join(HERE, "host", "restir_driver.cpp"),
        "-L" + HERE, "-lrestir_b200", "-Xlinker", "-rpath,$ORIGIN"]
    subprocess.check_call(drv, cwd=HERE)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
