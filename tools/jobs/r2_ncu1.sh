#!/bin/bash
# launch list of the timed window of the default bench (same command line first without ncu), then one full capture
mkdir -p gpurun_out
CMD="python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0"
$CMD > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err && \
ncu --nvtx --nvtx-include "timed_steps/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_l.json 2> gpurun_out/r2_ncu_l.err
echo "launch list rc=$?"; wc -l gpurun_out/r2_launches.csv
$CMD > gpurun_out/r2_ncu_plain2.json 2> gpurun_out/r2_ncu_plain2.err && \
ncu --nvtx --nvtx-include "timed_steps/" --set full --clock-control none --import-source on -k regex:env_step64 -c 2 -f -o gpurun_out/r2_env_step64 $CMD > gpurun_out/r2_ncu_f.json 2> gpurun_out/r2_ncu_f.err
echo "full capture rc=$?"; ls -la gpurun_out/r2_env_step64.ncu-rep
python -c "
import json; d=json.load(open('gpurun_out/r2_ncu_plain.json')); print('plain run: us/step', d['ms_per_step']*1e3, d['step_us']['series'])"
