#!/bin/bash
mkdir -p gpurun_out
for s in 0 3 5 7 9 11 14; do
  GCA_BALANCE_SKEW=$s python bench.py --steps 192 --warmup 16 --preroll 512 --preroll-groups 32 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 > gpurun_out/r2_skew_$s.json 2> gpurun_out/r2_skew_$s.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_skew_$s.json"))
print("skew $s: us/step %.2f  value %.3e  e2e %.3e  warm %.3e" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"], d["value_l2_warm"]))
PY
done
