#!/bin/bash
export GCA_SKIP_VERSION_CHECK=1 GCA_LIB_PATH=build/variants/cur.so
CMD="python bench.py --size 4096 --envs-per-gpu 1 --hidden device --steps 2 --warmup 24 --preroll 0 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0"
timeout 600 ncu --nvtx --nvtx-include "timed_steps/" --set full --clock-control none --import-source on -k regex:ca_tiled_list -c 2 -f -o gpurun_out/r2_tiled_list $CMD > gpurun_out/r2_ncu4096f.json 2> gpurun_out/r2_ncu4096f.err
echo "rc=$?"; ls -la gpurun_out/r2_tiled_list.ncu-rep
