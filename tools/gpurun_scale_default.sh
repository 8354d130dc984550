#!/bin/bash
# usage: gpurun_scale_default.sh N tag -- the driver's own command at N GPUs (weak headline + strong 8K in one run), summary printed
N=$1; tag=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo rc=$?
tail -2 gpurun_out/${tag}_bench_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_n$N.json"))
print("N=$N weak: ms", round(d["ms_per_step"],3), "Mrays/s", round(d["value"],1), "walked", round(d["value_walked"],1), "e2e ms", round(d["e2e"]["ms_per_frame"],3), "e2e devG ms", (d["e2e"].get("device_gbuffer") or {}).get("ms_per_frame"), "band_parity", d["band_parity"] and (d["band_parity"]["mismatching_reservoirs"], d["band_parity"]["mismatching_pixels_rgba8"], d["band_parity"]["pixels"]), "invalid", d["invalid"], "clocks", d["clocks"])
for k,v in d["configs"].items(): print("  ", k, "ms", round(v["ms_per_frame"],3), "band_parity", v.get("band_parity") and (v["band_parity"]["mismatching_reservoirs"], v["band_parity"]["pixels"]), "clocks", v.get("clocks"))
PY
