#!/bin/bash
# One GPU-box visit for profiles: tests touched since the last visit, then ncu launch list of the full loop and --set full
# captures of the env step (C3) and of the learner's tensor-core kernels.  Every profiled command first runs plain.
#   gpurun --timeout 2400 -- 'bash scripts/ncu_round.sh [tag]'
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wire_and_qmix_gpu.py tests/test_runners_gpu.py::test_train_dqn_c1_config_then_evaluate_and_main \
    tests/test_agent_gpu.py::test_bf16_path_does_not_drift_over_50_learn_steps tests/test_agent_gpu.py::test_bf16_path_tracks_fp32_at_the_bench_batch_sizes \
    -m gpu -q --maxfail=20 > gpurun_out/${tag}_pytest_new.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest_new.log
NCU="ncu --clock-control none"
python scripts/loop_profile.py c3 4 > gpurun_out/${tag}_loop_plain.log 2>&1 && \
  $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file gpurun_out/${tag}_launches_loop_c3.csv python scripts/loop_profile.py c3 4 > gpurun_out/${tag}_loop_ncu.log 2>&1
echo "launch list rc=$?"; cat gpurun_out/${tag}_loop_plain.log
python scripts/env_step_profile.py c3 40 > gpurun_out/${tag}_env_plain.log 2>&1 && \
  $NCU --set full --import-source on -k regex:env_step_kernel -s 40 -c 2 -o gpurun_out/${tag}_env_step_c3 python scripts/env_step_profile.py c3 40 > gpurun_out/${tag}_env_ncu.log 2>&1
echo "env capture rc=$?"
python scripts/learn_step_profile.py bf16 4096 > gpurun_out/${tag}_learn_plain.log 2>&1 && \
  $NCU --set full --import-source on -k regex:"pair|persistent|conv1_obs|gemm_bf16_tn|gemm_bf16_tc|conv_wgrad|clip_adam" -s 150 -c 40 -o gpurun_out/${tag}_learn_kernels python scripts/learn_step_profile.py bf16 4096 > gpurun_out/${tag}_learn_ncu.log 2>&1
echo "learn capture rc=$?"; cat gpurun_out/${tag}_learn_plain.log
ls -la gpurun_out | tail -12
