#!/bin/bash
# usage: gpurun_variants.sh name1 name2 ... -- benches each restir-vulkan_b200/variants/lib_<name>.so (no tests)
for v in "$@"; do
  RESTIR_B200_LIB=$PWD/restir-vulkan_b200/variants/lib_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/try_$v.json 2> gpurun_out/try_$v.err || tail -5 gpurun_out/try_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/try_$v.json"))
print("$v", "ms/frame", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernel_ms"].items() if "trace" in k}, d["parity_sample"])
PY
done
