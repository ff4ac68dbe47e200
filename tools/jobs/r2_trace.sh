#!/bin/bash
mkdir -p gpurun_out
PREROLL=512 GCA_LIB_PATH=build/variants/trace.so python tools/phase_trace.py > gpurun_out/r2_trace_stationary.log 2>&1
tail -40 gpurun_out/r2_trace_stationary.log
