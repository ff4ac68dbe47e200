#!/bin/bash
# parity tests, then the stationary bench (short), then the phase trace of the same build
mkdir -p gpurun_out
TAG=${1:-iter}
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --steps 192 --warmup 16 --preroll 512 --preroll-groups 32 --no-cpu-baseline --no-obs-leg --no-other-configs --long-run 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("${TAG}: us/step %.2f  value %.3e  e2e %.3e  warm %.3e  frac %.3f" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"], d["value_l2_warm"], d["roofline"]["frac"]))
PY
if [ -f build/variants/trace.so ]; then
PREROLL=512 GCA_LIB_PATH=build/variants/trace.so python tools/phase_trace.py > gpurun_out/${TAG}_trace.log 2>&1
grep -A3 "^cold kernel" gpurun_out/${TAG}_trace.log | head -8; grep "CTA time\|launch order" gpurun_out/${TAG}_trace.log | head -4
fi
