#!/bin/bash
# config 4 (one 4096x4096 grid): the command without ncu, then the launch list of its two timed env steps
CMD="python bench.py --size 4096 --envs-per-gpu 1 --hidden device --steps 2 --warmup 24 --preroll 0 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0"
$CMD > gpurun_out/r2_4096_plain.json 2> gpurun_out/r2_4096_plain.err && \
timeout 600 ncu --nvtx --nvtx-include "timed_steps/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_4096.csv $CMD > gpurun_out/r2_ncu4096.json 2> gpurun_out/r2_ncu4096.err
echo "launch list rc=$?"; wc -l gpurun_out/r2_launches_4096.csv
python -c "
import json; d=json.load(open('gpurun_out/r2_4096_plain.json')); print('plain run: us/step', d['ms_per_step']*1e3, d['step_us']['series'])"
