#!/bin/bash
# usage: gpurun_wide.sh [variant ...] -- parity tests of the default build (wide walk), then A/B benches of the headline: wide
# (default, with the whole-frame parity sample against the CPU oracle), binary image, restir-vulkan_b200/variants/lib_<variant>.so
python -m pytest tests/test_golden_frames.py tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -3
run() { # tag, args...
  tag=$1; shift
  python bench.py --steps 10 --warmup 3 --no-e2e --no-suite "$@" > gpurun_out/try_$tag.json 2> gpurun_out/try_$tag.err || tail -5 gpurun_out/try_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/try_$tag.json"))
print("$tag", "| ms/frame", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernel_ms"].items()}, "walked", d["rays_walked_per_frame"], "parity", d.get("parity_sample"), d["bvh"])
PY
}
run wide
run image --traversal image --no-cpu-baseline
for v in "$@"; do
  export RESTIR_B200_LIB=$PWD/restir-vulkan_b200/variants/lib_$v.so
  run $v --no-cpu-baseline
done
unset RESTIR_B200_LIB
