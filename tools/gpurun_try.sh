#!/bin/bash
# usage: gpurun_try.sh <tag> [bench args]  -- runs GPU parity tests + a short bench, prints kernel_ms
tag=$1; shift
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/try_$tag.json 2> gpurun_out/try_$tag.err || tail -5 gpurun_out/try_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/try_$tag.json"))
print("ms/frame", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernel_ms"].items()})
PY
