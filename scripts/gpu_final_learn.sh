#!/bin/bash
# Final-evidence visit, part 2: --set full capture of one whole learn step (bf16 path, B = 4096): every tcgen05 kernel, the
# replay sample and clip + Adam.
tag=${1:-r02}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
python scripts/learn_step_profile.py bf16 4096 > gpurun_out/${tag}_learn_plain.log 2>&1 && \
  $NCU --set full --import-source on -k regex:"pair|persistent|conv1_obs|gemm_bf16_tn|gemm_bf16_tc|clip_adam|replay" -s 141 -c 47 -o gpurun_out/${tag}_learn_kernels python scripts/learn_step_profile.py bf16 4096 > gpurun_out/${tag}_learn_ncu.log 2>&1
echo "learn capture rc=$?"; cat gpurun_out/${tag}_learn_plain.log
python scripts/split_precision_probe.py > gpurun_out/${tag}_split_precision_probe.txt 2>&1; echo "probe rc=$?"
python scripts/gemm_tc_perf.py 2>&1 | grep -E "dgrad|wgrad|fwd" > gpurun_out/${tag}_gemm_tc_perf.txt
python scripts/learn_step_profile.py fp32 4096 > gpurun_out/${tag}_learn_fp32_plain.log 2>&1 && \
  $NCU --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.sum -c 400 --csv --log-file gpurun_out/${tag}_launches_learn_fp32.csv python scripts/learn_step_profile.py fp32 4096 > /dev/null 2>&1
echo "fp32 launch list rc=$?"; cat gpurun_out/${tag}_learn_fp32_plain.log
