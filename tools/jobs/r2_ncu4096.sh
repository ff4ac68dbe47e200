#!/bin/bash
export GCA_SKIP_VERSION_CHECK=1 GCA_LIB_PATH=build/variants/cur.so
CMD="python bench.py --size 4096 --envs-per-gpu 1 --hidden device --steps 2 --warmup 24 --preroll 0 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0"
timeout 600 ncu --nvtx --nvtx-include "timed_steps/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_4096.csv $CMD > gpurun_out/r2_ncu4096.json 2> gpurun_out/r2_ncu4096.err
echo "rc=$?"; wc -l gpurun_out/r2_launches_4096.csv
