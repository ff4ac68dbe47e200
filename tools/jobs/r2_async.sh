#!/bin/bash
# parity of the asynchronous variant (quick subset, hang guard), then A/B on the bench workload: r2_async.sh "name:skew ..."
v0=$1; shift
GCA_SKIP_VERSION_CHECK=1 GCA_LIB_PATH=build/variants/$v0.so timeout 300 python -m pytest tests -x -q -m gpu -k "env_step_parity or env_counts_around or full_batch_4096 or dense_front or load_balancing or ignition_burst or dousing_heavy" 2>&1 | tail -6
run() {  # name skew
  GCA_BALANCE_SKEW=$2 GCA_SKIP_VERSION_CHECK=1 GCA_LIB_PATH=build/variants/$1.so timeout 300 python bench.py --steps 128 --warmup 8 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 > gpurun_out/ab_$1_$2.json 2> gpurun_out/ab_$1_$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$1_$2.json"))
    print("$1 skew=$2: us/step %.2f  value %.3e  e2e %.3e (%.1f us) warm %.3e" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"], d["e2e"].get("us_per_step",0), d["value_l2_warm"]))
except Exception as e:
    print("$1 skew=$2: failed", e); print(open("gpurun_out/ab_$1_$2.err").read()[-1500:])
PY
}
for vs in "$@"; do run ${vs%%:*} ${vs##*:}; done
