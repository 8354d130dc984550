#!/bin/bash
# benches every configuration of bench.py's CONFIGS table (device-resident numbers only)
for c in sponza_1080p_biased4 sponza_1080p_unbiased3 cornell_720p_biased4 office_2160p_unbiased3 sponza_8k_1m_lights; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --config $c > gpurun_out/cfg_$c.json 2> gpurun_out/cfg_$c.err || tail -5 gpurun_out/cfg_$c.err
  python - <<PY
import json
d=json.load(open("gpurun_out/cfg_$c.json"))
print("$c", "ms/frame", round(d["ms_per_step"],4), "Mrays/s", round(d["value"],1), "walked/frame", d["rays_walked_per_frame"], "of", d["rays_per_frame"], {k:round(v,3) for k,v in d["kernel_ms"].items()})
PY
done
