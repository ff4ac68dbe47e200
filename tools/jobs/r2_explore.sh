#!/bin/bash
# GPU job 1 of round 2: episode profile of the bench workload, stationarity of the pre-roll, new bench line
set -x
mkdir -p gpurun_out
python tools/phase_series.py gpurun_out/r2_series_reset.json --steps 1536 > gpurun_out/r2_series_reset.log 2>&1
python tools/phase_series.py gpurun_out/r2_series_pre352.json --steps 768 --preroll 352 --groups 16 > gpurun_out/r2_series_pre352.log 2>&1
python tools/phase_series.py gpurun_out/r2_series_pre512.json --steps 768 --preroll 512 --groups 32 > gpurun_out/r2_series_pre512.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
python bench.py --no-cpu-baseline --no-obs-leg --no-other-configs > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -3 gpurun_out/r2_series_*.log; cut -c1-1500 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
