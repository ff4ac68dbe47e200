#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err ) 2>&1 | grep real; tail -3 gpurun_out/r2_bench_e.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref_e.json 2> gpurun_out/r2_ref_e.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_e.json"))
print("bench: us/step %.2f value %.3e frac %.3f e2e %.3e (%.1f us) warm %.3e window/longrun %.3f launches %d" % (d["ms_per_step"]*1e3, d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["us_per_step"], d["value_l2_warm"], d["workload_stats"]["long_run"]["timed_window_over_long_run"], d["gpu_launches"]))
print(d["step_us"]["series"])
print({k: round(v["us_per_step"],1) for k,v in d["with_observation"].items()})
for o in [x for x in d.get("other_configs", []) if "roofline" in x]: print(o["workload"][:40], "us/step %.1f value %.3e frac %.3f launches/step %.1f" % (o["us_per_step"], o["value"], o["roofline"]["frac"], o["launches_per_step"]))
print(d.get("cpu_baseline"))
r=json.load(open("gpurun_out/r2_ref_e.json")); print("reference arm:", r["value"], r["config"].get("cpu_sample_envs_per_step"))
PY
