#!/bin/bash
# usage: gpurun_dedupe.sh -- GPU parity tests, then the headline and the 8K / 1 M-light frame with and without the segment table (A/B)
python -m pytest tests/test_golden_frames.py tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -3
run() { # tag, args...
  tag=$1; shift
  python bench.py --warmup 3 --no-e2e --no-suite "$@" > gpurun_out/try_$tag.json 2> gpurun_out/try_$tag.err || tail -5 gpurun_out/try_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/try_$tag.json"))
print("$tag", "| ms/frame", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernel_ms"].items()}, "asked", d["rays_per_frame"], "walked", d["rays_walked_per_frame"], "cached", d.get("rays_answered_by_cached_occluder_per_frame"), "parity", d.get("parity_sample"))
PY
}
run dedupe_on --steps 10 --ray-elision dedupe
run dedupe_off --steps 10  --no-cpu-baseline
run 8k_dedupe_on --steps 4 --config sponza_8k_1m_lights --no-cpu-baseline --ray-elision dedupe
run 8k_dedupe_off --steps 4 --config sponza_8k_1m_lights  --no-cpu-baseline
