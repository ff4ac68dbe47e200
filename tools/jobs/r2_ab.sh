#!/bin/bash
# A/B of variant libraries on the default bench workload, interleaved twice
for rep in 1 2; do
for v in "$@"; do
  GCA_SKIP_VERSION_CHECK=1 GCA_LIB_PATH=build/variants/$v.so python bench.py --steps 128 --warmup 8 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$v.json"))
    print("$v: us/step %.2f  value %.3e  e2e %.3e (%.1f us) sync %.3e (%.1f us) warm %.3e" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"], d["e2e"].get("us_per_step",0), d["e2e"].get("sync",{}).get("value",0), d["e2e"].get("sync",{}).get("us_per_step",0), d["value_l2_warm"]))
except Exception as e:
    print("$v: failed", e); print(open("gpurun_out/ab_$v.err").read()[-1500:])
PY
done
done
