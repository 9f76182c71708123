#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
run() { echo "== $1"; env $1 timeout 300 python scripts/step_time_trace.py c3 300 2>&1 | awk '/us per launch/{s+=$(NF-3); n++; if (n==1||n==8||n==15) printf "%s ", $(NF-3)} END{printf " mean %.1f us\n", s/n}'; }
D=$PWD/dqn_marl_b200
run "MQ_X=0" | tee gpurun_out/${tag}_ab.txt
run "MARL_B200_SO=$D/libmarl_b200_u1.so" | tee -a gpurun_out/${tag}_ab.txt
run "MARL_B200_SO=$D/libmarl_b200_c3.so" | tee -a gpurun_out/${tag}_ab.txt
run "MARL_B200_SO=$D/libmarl_b200_c3u1.so" | tee -a gpurun_out/${tag}_ab.txt
run "MARL_B200_SO=$D/libmarl_b200_c3.so MQ_ENV_HASH_POW2=1" | tee -a gpurun_out/${tag}_ab.txt
