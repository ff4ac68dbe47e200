#!/bin/bash
bash tools/jobs/r2_ab.sh "$@"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -3 gpurun_out/r2_bench_d.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_d.json"))
print("bench: us/step %.2f value %.3e frac %.3f e2e %.3e (%.1f us) warm %.3e window/longrun %.3f" % (d["ms_per_step"]*1e3, d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["us_per_step"], d["value_l2_warm"], d["workload_stats"]["long_run"]["timed_window_over_long_run"]))
print(d["with_observation"])
print(d.get("cpu_baseline"))
PY
