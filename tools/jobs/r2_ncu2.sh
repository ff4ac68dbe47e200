#!/bin/bash
# full capture of the whole-grid bit-board kernel on BASELINE config 3
mkdir -p gpurun_out
CMD="python bench.py --size 256 --envs-per-gpu 1024 --hidden device --steps 4 --warmup 40 --preroll 0 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0"
$CMD > gpurun_out/r2_ncu_bb_plain.json 2> gpurun_out/r2_ncu_bb_plain.err && \
ncu --nvtx --nvtx-include "timed_steps/" --set full --clock-control none --import-source on -k regex:env_step_bb -c 2 -f -o gpurun_out/r2_env_step_bb $CMD > gpurun_out/r2_ncu_bb.json 2> gpurun_out/r2_ncu_bb.err
echo "bb capture rc=$?"; ls -la gpurun_out/r2_env_step_bb.ncu-rep
python -c "
import json; d=json.load(open('gpurun_out/r2_ncu_bb_plain.json')); print('plain run: us/step', d['ms_per_step']*1e3, d['step_us']['series'])"
