#!/bin/bash
# usage: gpurun_cache.sh -- GPU parity tests, then the headline with the occluder cache on / off (A/B)
python -m pytest tests/test_golden_frames.py tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -3
run() { # tag, args...
  tag=$1; shift
  python bench.py --steps 10 --warmup 3 --no-e2e --no-suite "$@" > gpurun_out/try_$tag.json 2> gpurun_out/try_$tag.err || tail -5 gpurun_out/try_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/try_$tag.json"))
print("$tag", "| ms/frame", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernel_ms"].items()}, "asked", d["rays_per_frame"], "walked", d["rays_walked_per_frame"], "cached", d.get("rays_answered_by_cached_occluder_per_frame"), "parity", d.get("parity_sample"))
PY
}
run cache_on
run cache_off --occluder-cache off --no-cpu-baseline
