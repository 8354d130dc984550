"""The C-ABI library loads without a GPU and exports every symbol include/restir_b200.h declares."""
import ctypes as C
import os
import re

import parity_harness as ph

ROOT = ph.ROOT


def test_header_symbols_are_exported():
    header = open(os.path.join(ROOT, "include", "restir_b200.h")).read()
    declared = set(re.findall(r"\b(restir_[a-z_0-9]+)\s*\(", header))
    assert declared == set(ph.capi.EXPORTS), declared ^ set(ph.capi.EXPORTS)
    lib = ph.capi.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), name


def test_headers_compile_as_c11():
    import subprocess
    src = '#include "restir_b200.h"\n#include "restir_capture.h"\nint main(void){return sizeof(restir_reservoir)==64&&sizeof(restir_capture_header)==72?0:1;}\n'
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                   input=src.encode(), check=True)


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        return
    try:
        ph.capi.RestirContext(0)
    except ph.capi.RestirError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("context creation must fail without a GPU")


def test_product_never_touches_the_oracle():
    pkg_dir = os.path.join(ROOT, "restir-vulkan_b200")
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(base, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text and "restir_oracle" not in text, f
