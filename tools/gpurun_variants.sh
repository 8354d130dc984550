#!/bin/bash
# usage: gpurun_variants.sh [--config NAME] name1 name2 ... -- per restir-vulkan_b200/variants/lib_<name>.so ("default" = the in-tree build): the
# golden-frame and shadow-ray parity tests against that build, then a short device-resident bench (ms/frame, per-kernel ms)
cfg=sponza_1080p_unbiased5
if [ "$1" = "--config" ]; then cfg=$2; shift 2; fi
for v in "$@"; do
  if [ "$v" = "default" ]; then unset RESTIR_B200_LIB; else export RESTIR_B200_LIB=$PWD/restir-vulkan_b200/variants/lib_$v.so; fi
  par=$(python -m pytest tests/test_golden_frames.py tests/test_gpu_parity.py -m gpu -x -q -k "golden or shadow_rays or unbiased or occluder" 2>&1 | tail -1)
  python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-suite > gpurun_out/try_$v.json 2> gpurun_out/try_$v.err || tail -5 gpurun_out/try_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/try_$v.json"))
print("$v", "| parity:", "$par", "| ms/frame", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernel_ms"].items()}, "walked", d["rays_walked_per_frame"])
PY
done
unset RESTIR_B200_LIB
