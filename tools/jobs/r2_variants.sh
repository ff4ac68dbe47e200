#!/bin/bash
# bench the listed variant libraries (build/variants/NAME.so) on the stationary workload
for v in "$@"; do
  GCA_LIB_PATH=build/variants/$v.so python bench.py --steps 192 --warmup 16 --preroll 512 --preroll-groups 32 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_$v.json"))
    print("$v: us/step %.2f  value %.3e  e2e %.3e  warm %.3e" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"], d["value_l2_warm"]))
except Exception as e:
    print("$v: failed", e); print(open("gpurun_out/var_$v.err").read()[-800:])
PY
done
