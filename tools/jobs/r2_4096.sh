#!/bin/bash
for g in 1 0; do
GCA_TILED_GRAPH=$g python bench.py --size 4096 --envs-per-gpu 1 --hidden device --steps 24 --warmup 24 --preroll 0 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 > gpurun_out/t4096.json 2> gpurun_out/t4096.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/t4096.json"))
    print("4096x4096 graph=$g: us/step %.1f value %.3e warm %.3e e2e %.1f us" % (d["ms_per_step"]*1e3, d["value"], d["value_l2_warm"], d["e2e"]["us_per_step"]), d["workload_stats"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/t4096.err").read()[-1500:])
PY
done
python -m pytest tests -x -q -m gpu -k "tiled or other_widths or many_envs" 2>&1 | tail -3
