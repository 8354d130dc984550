#!/bin/bash
# usage: gpurun_scale.sh N config scaling [steps]
N=$1; cfg=$2; sc=$3; steps=${4:-5}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps $steps --warmup 3 --config $cfg --scaling $sc --no-e2e --no-cpu-baseline > gpurun_out/scale_${cfg}_${sc}_n$N.json 2> gpurun_out/scale_${cfg}_${sc}_n$N.err || tail -5 gpurun_out/scale_${cfg}_${sc}_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/scale_${cfg}_${sc}_n$N.json"))
print("$cfg $sc N=$N", "ms/frame", round(d["ms_per_step"],3), "Mrays/s", round(d["value"],1), "halo_misses", d["halo_misses"], "timeouts", d.get("halo_wait_timeouts"), d["config"]["parallelism"])
PY
