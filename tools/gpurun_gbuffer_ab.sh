python -m pytest tests/test_gbuffer_pass.py -m gpu -x -q 2>&1 | tail -2
for v in default base default base; do
  if [ "$v" = "default" ]; then unset RESTIR_B200_LIB; else export RESTIR_B200_LIB=$PWD/restir-vulkan_b200/variants/lib_$v.so; fi
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-suite > gpurun_out/try_gb_$v.json 2> gpurun_out/try_gb_$v.err || tail -5 gpurun_out/try_gb_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/try_gb_$v.json"))
print("$v", "frame", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_frame"],4), "e2e devG", round(d["e2e"]["device_gbuffer"]["ms_per_frame"],4), "parity", d["parity_sample"] and d["parity_sample"]["mismatching_reservoirs"])
PY
done
