#!/bin/bash
# usage: gpurun_env.sh "ENV=1 ENV2=2" tag ...   (pairs) -- bench under each environment
while [ $# -ge 2 ]; do
  envs=$1; tag=$2; shift 2
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/try_$tag.json 2> gpurun_out/try_$tag.err || tail -5 gpurun_out/try_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/try_$tag.json"))
print("$tag", "ms/frame", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernel_ms"].items() if "trace" in k})
PY
done
