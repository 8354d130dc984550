#!/bin/bash
# usage: gpurun_capture.sh <tag>   -- the evidence set of a round: GPU tests, the default bench line, the ncu launch list of the same
# command shape, and `ncu --set full` captures of one headline frame and one biased frame (read here with profiles/ncu_extract.py)
tag=$1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-suite > gpurun_out/${tag}_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"trace_|omni_|unbiased_|lighting_kernel|spatial_reuse" -s 35 -c 7 -f \
    -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-suite > gpurun_out/${tag}_full_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"trace_|omni_|unbiased_|lighting_kernel|spatial_reuse" -s 25 -c 5 -f \
    -o gpurun_out/${tag}_full_biased python bench.py --config sponza_1080p_biased4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-suite > gpurun_out/${tag}_full_biased_run.log 2>&1
ls -la gpurun_out/${tag}_*
