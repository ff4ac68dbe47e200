#!/bin/bash
for b in 0 2 4 8 16 32; do
python bench.py --steps 128 --warmup 8 --balance-every $b --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 > gpurun_out/bal.json 2> gpurun_out/bal.err
python - <<PY
import json
d=json.load(open("gpurun_out/bal.json"))
print("balance_every $b: us/step %.2f value %.3e warm %.3e e2e %.1f us" % (d["ms_per_step"]*1e3, d["value"], d["value_l2_warm"], d["e2e"]["us_per_step"]))
PY
done
