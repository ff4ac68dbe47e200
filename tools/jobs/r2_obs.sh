#!/bin/bash
python -m pytest tests -x -q -m gpu -k "fused_observation or step_host or async" 2>&1 | tail -3
for v in "" stcs; do
  if [ -n "$v" ]; then export GCA_LIB_PATH=build/variants/$v.so; fi
  python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-other-configs --long-run 0 > gpurun_out/obs.json 2> gpurun_out/obs.err
  python - <<PY
import json
d=json.load(open("gpurun_out/obs.json"))
print("[$v] us/step %.1f warm %.1f e2e %.1f us ; obs:" % (d["ms_per_step"]*1e3, 4096*4096*4/d["value_l2_warm"]*1e6, d["e2e"]["us_per_step"]), {k: round(x["us_per_step"],1) for k,x in d["with_observation"].items()})
PY
done
