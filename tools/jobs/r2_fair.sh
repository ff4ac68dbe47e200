#!/bin/bash
# parity of the first variant (quick subset), then A/B of all variants on the bench workload
v0=$1
GCA_SKIP_VERSION_CHECK=1 GCA_LIB_PATH=build/variants/$v0.so timeout 600 python -m pytest tests -x -q -m gpu -k "env_step_parity or env_counts_around or full_batch_4096 or dense_front or load_balancing" 2>&1 | tail -4
bash tools/jobs/r2_ab.sh "$@"
