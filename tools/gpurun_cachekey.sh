#!/bin/bash
# usage: gpurun_cachekey.sh -- occluder cache keyed by light index vs by direction on three configurations (A/B)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "occluder or elision or golden" 2>&1 | tail -2
run() { tag=$1; shift
  python bench.py --warmup 3 --no-e2e --no-suite --no-cpu-baseline "$@" > gpurun_out/try_$tag.json 2> gpurun_out/try_$tag.err || tail -5 gpurun_out/try_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/try_$tag.json"))
print("$tag", "| ms/frame", round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["kernel_ms"].items() if "trace" in k}, "walked", round(d["rays_walked_per_frame"]), "cached", round(d["rays_answered_by_cached_occluder_per_frame"]))
PY
}
for key in light direction; do
  run head_$key --steps 10 --occluder-cache $key
  run office_$key --steps 6 --config office_2160p_unbiased3 --occluder-cache $key
  run 8k_$key --steps 4 --config sponza_8k_1m_lights --occluder-cache $key
done
