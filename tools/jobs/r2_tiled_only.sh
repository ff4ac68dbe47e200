#!/bin/bash
export GCA_SKIP_VERSION_CHECK=1 GCA_LIB_PATH=build/variants/cur.so
timeout 900 python -m pytest tests -x -q -m gpu -k "tiled or large_single or injected_uniforms_128 or other_widths or many_envs or conditional_reset or env_api_surface" 2>&1 | tail -4
timeout 300 python bench.py --size 4096 --envs-per-gpu 1 --hidden device --steps 24 --warmup 24 --preroll 0 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 > gpurun_out/t4096_1.json 2> gpurun_out/t4096_1.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/t4096_1.json"))
    print("4096: us/step %.1f value %.3e warm %.3e e2e %.1f" % (d["ms_per_step"]*1e3, d["value"], d["value_l2_warm"], d["e2e"]["us_per_step"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/t4096_1.err").read()[-800:])
PY
