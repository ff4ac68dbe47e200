#!/bin/bash
for rep in 1 2; do
for m in write write+read; do
  python bench.py --steps 128 --warmup 8 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 --flush-mode $m > gpurun_out/fl_$m.json 2> gpurun_out/fl_$m.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/fl_$m.json"))
    s=d["step_us"]
    print("$m: us/step %.2f (min %.1f med %.1f max %.1f) value %.3e  e2e %.3e (%.1f us) sync %.3e (%.1f us) warm %.3e" % (d["ms_per_step"]*1e3, s["min"], s["median"], s["max"], d["value"], d["e2e"]["value"], d["e2e"].get("us_per_step",0), d["e2e"].get("sync",{}).get("value",0), d["e2e"].get("sync",{}).get("us_per_step",0), d["value_l2_warm"]))
except Exception as e:
    print("$m: failed", e); print(open("gpurun_out/fl_$m.err").read()[-1500:])
PY
done
done
