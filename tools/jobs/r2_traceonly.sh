#!/bin/bash
TAG=${1:-t}
PREROLL=512 GCA_LIB_PATH=build/variants/trace.so python tools/phase_trace.py > gpurun_out/${TAG}_trace.log 2>&1
grep -B2 -A3 "^cold kernel" gpurun_out/${TAG}_trace.log | head -14; grep "CTA time\|launch order" gpurun_out/${TAG}_trace.log | head -2
