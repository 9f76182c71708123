#!/bin/bash
# GPU visit: lazy-noise scoring — parity tests of every env kernel variant, ambiguity rate, step times on C3 / C2 / C5.
tag=${1:-r02g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py tests/test_multi_layout_gpu.py tests/test_wire_and_qmix_gpu.py -m gpu -q --maxfail=20 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
D=$PWD/dqn_marl_b200
for wl in c3 c2; do
  MARL_B200_SO=$D/libmarl_b200_envtrace.so timeout 300 python scripts/env_phase_trace.py $wl 300 > gpurun_out/${tag}_phase_$wl.txt 2>&1; tail -3 gpurun_out/${tag}_phase_$wl.txt
done
for wl in c3 c2 c5; do
  timeout 300 python scripts/step_time_trace.py $wl 300 > gpurun_out/${tag}_steps_$wl.txt 2>&1
  awk '/us per launch/{s+=$(NF-3); n++} END{printf "'$wl' mean %.1f us per launch\n", s/n}' gpurun_out/${tag}_steps_$wl.txt
done
