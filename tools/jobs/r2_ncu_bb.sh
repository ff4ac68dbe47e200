#!/bin/bash
export GCA_SKIP_VERSION_CHECK=1 GCA_LIB_PATH=build/variants/cur.so
CMD="python bench.py --size 256 --envs-per-gpu 1024 --hidden device --steps 4 --warmup 40 --preroll 0 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0"
timeout 600 ncu --nvtx --nvtx-include "timed_steps/" --set full --clock-control none --import-source on -k regex:env_step_bb -c 1 -f -o gpurun_out/r2c_env_step_bb $CMD > gpurun_out/r2c_ncu_bb.json 2> gpurun_out/r2c_ncu_bb.err
echo "bb capture rc=$?"; ls -la gpurun_out/r2c_env_step_bb.ncu-rep
