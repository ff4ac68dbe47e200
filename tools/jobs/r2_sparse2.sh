#!/bin/bash
export GCA_SKIP_VERSION_CHECK=1 GCA_LIB_PATH=build/variants/cur.so
timeout 900 python -m pytest tests -x -q -m gpu -k "tiled or other_widths or many_envs or large_single or injected_uniforms_128" 2>&1 | tail -3
timeout 300 python bench.py --size 4096 --envs-per-gpu 1 --hidden device --steps 24 --warmup 24 --preroll 0 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 > gpurun_out/t4096_1.json 2> gpurun_out/t4096_1.err
timeout 300 python bench.py --size 256 --envs-per-gpu 1024 --hidden device --generic-tiles --steps 30 --warmup 40 --preroll 0 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 > gpurun_out/t256_1.json 2> gpurun_out/t256_1.err
python - <<PY
import json
for f in ("t4096_1","t256_1"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "us/step %.1f value %.3e warm %.3e e2e %.1f us" % (d["ms_per_step"]*1e3, d["value"], d["value_l2_warm"], d["e2e"]["us_per_step"]))
    except Exception as e:
        print("failed", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
CMD="python bench.py --size 4096 --envs-per-gpu 1 --hidden device --steps 2 --warmup 24 --preroll 0 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0"
timeout 600 ncu --nvtx --nvtx-include "timed_steps/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_4096.csv $CMD > gpurun_out/r2_ncu4096.json 2> gpurun_out/r2_ncu4096.err
