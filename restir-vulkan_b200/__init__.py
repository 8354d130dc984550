"""restir-vulkan_b200 — B200-native ReSTIR resampling passes behind the reference's pass interface.

Layout:
  csrc/      hand-written sm_100a CUDA kernels + the C-ABI context (include/restir_b200.h)
  host/      host C++: scene-side builders (AABB tree, lights, alias table) and the C++ mirror of the
             reference's pass classes (passes.hpp)
  capi.py    ctypes binding of the C ABI (what tests and bench.py call)
  fixtures.py  scene inputs (baked reference scenes, procedural scenes, material tables)
  bands.py   row-band multi-GPU driver (band cutting, halo sizing, neighbour wiring; torch.distributed is plumbing)
  capture.py frame captures (include/restir_capture.h): numpy writer / reader

The directory name carries a hyphen, so it is loaded through __graft_entry__.load_package() as the
module `restir_vulkan_b200`.
"""
from . import capi, fixtures  # noqa: F401
from .capi import RestirContext, RestirError  # noqa: F401
