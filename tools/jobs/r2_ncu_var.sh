#!/bin/bash
# full ncu capture of env_step64 for a variant library: r2_ncu_var.sh NAME [extra env]
v=$1
export GCA_SKIP_VERSION_CHECK=1 GCA_LIB_PATH=build/variants/$v.so
CMD="python bench.py --steps 4 --warmup 5 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0"
timeout 600 ncu --nvtx --nvtx-include "timed_steps/" --set full --clock-control none --import-source on -k regex:env_step64 -c 1 -f -o gpurun_out/r2_env_step64_$v $CMD > gpurun_out/r2_ncu_$v.json 2> gpurun_out/r2_ncu_$v.err
echo "full capture rc=$?"; ls -la gpurun_out/r2_env_step64_$v.ncu-rep
