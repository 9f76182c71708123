#!/bin/bash
# GPU visit: A/B of the CTA-per-env env_step variants (resident CTAs, health copy, table size) on C3, tests of the new table.
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py tests/test_multi_layout_gpu.py -m gpu -q --maxfail=20 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
run() { echo "== $1"; env $1 timeout 300 python scripts/step_time_trace.py c3 300 2>&1 | awk '/us per launch/{s+=$(NF-3); n++; if (n==1||n==8||n==15) printf "%s ", $(NF-3)} END{printf " mean %.1f us\n", s/n}'; }
run "MQ_X=0" | tee gpurun_out/${tag}_ab.txt
run "MQ_ENV_SMEM_PAD=12000" | tee -a gpurun_out/${tag}_ab.txt
run "MQ_ENV_HSM=0" | tee -a gpurun_out/${tag}_ab.txt
run "MQ_ENV_HSM=0 MQ_ENV_HASH_POW2=1" | tee -a gpurun_out/${tag}_ab.txt
run "MQ_ENV_HASH_POW2=1 MQ_ENV_SMEM_PAD=4000" | tee -a gpurun_out/${tag}_ab.txt
run "MQ_X=0" | tee -a gpurun_out/${tag}_ab.txt
