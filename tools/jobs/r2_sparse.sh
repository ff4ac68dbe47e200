#!/bin/bash
# sparse tile activity: parity of the tiled path, then config 4 / config 3 (generic tiles) A/B against the dense passes
export GCA_SKIP_VERSION_CHECK=1 GCA_LIB_PATH=build/variants/cur.so
timeout 900 python -m pytest tests -x -q -m gpu -k "tiled or other_widths or many_envs or large_single or injected_uniforms_128 or windy or v3 or conditional_reset or operator_level or golden" 2>&1 | tail -5
for sp in 1 0; do
GCA_TILED_SPARSE=$sp timeout 300 python bench.py --size 4096 --envs-per-gpu 1 --hidden device --steps 24 --warmup 24 --preroll 0 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 > gpurun_out/t4096_$sp.json 2> gpurun_out/t4096_$sp.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/t4096_$sp.json"))
    print("4096x4096 sparse=$sp: us/step %.1f value %.3e warm %.3e e2e %.1f us" % (d["ms_per_step"]*1e3, d["value"], d["value_l2_warm"], d["e2e"]["us_per_step"]), d["workload_stats"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/t4096_$sp.err").read()[-1500:])
PY
GCA_TILED_SPARSE=$sp timeout 300 python bench.py --size 256 --envs-per-gpu 1024 --hidden device --generic-tiles --steps 30 --warmup 40 --preroll 0 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 > gpurun_out/t256_$sp.json 2> gpurun_out/t256_$sp.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/t256_$sp.json"))
    print("256x256x1024 generic sparse=$sp: us/step %.1f value %.3e warm %.3e" % (d["ms_per_step"]*1e3, d["value"], d["value_l2_warm"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/t256_$sp.err").read()[-1500:])
PY
done
