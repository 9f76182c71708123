#!/bin/bash
# Final-evidence visit (1 GPU), part 1: full GPU test suite, both bench arms, launch lists, --set full captures of the env step
# (gpurun returns at most 64 MiB per call: the learner capture is scripts/gpu_final_learn.sh).
tag=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "reference arm rc=$?"
NCU="ncu --clock-control none"
# launch list of the bench command itself (env leg = the headline's timed region; capped) and of the learner loop
python bench.py --steps 20 --warmup 3 --no-learner --no-cpu --replay-batch 0 --c2-steps 0 > /dev/null 2>&1 && \
  $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/${tag}_launches_bench_env.csv python bench.py --steps 20 --warmup 3 --no-learner --no-cpu --replay-batch 0 --c2-steps 0 > gpurun_out/${tag}_bench_env_ncu.log 2>&1
echo "bench launch list rc=$?"
python scripts/loop_profile.py c3 4 > gpurun_out/${tag}_loop_plain.log 2>&1 && \
  $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file gpurun_out/${tag}_launches_loop_c3.csv python scripts/loop_profile.py c3 4 > gpurun_out/${tag}_loop_ncu.log 2>&1
echo "loop launch list rc=$?"; cat gpurun_out/${tag}_loop_plain.log
python scripts/env_step_profile.py c3 40 > gpurun_out/${tag}_env_plain.log 2>&1 && \
  $NCU --set full --import-source on -k regex:env_step_kernel -s 40 -c 2 -o gpurun_out/${tag}_env_step_c3 python scripts/env_step_profile.py c3 40 > gpurun_out/${tag}_env_ncu.log 2>&1
echo "env capture rc=$?"
python scripts/env_step_profile.py c2 40 > /dev/null 2>&1 && \
  $NCU --set full --import-source on -k regex:env_step_kernel -s 40 -c 2 -o gpurun_out/${tag}_env_step_c2 python scripts/env_step_profile.py c2 40 > gpurun_out/${tag}_env_c2_ncu.log 2>&1
echo "env c2 capture rc=$?"
ls -la gpurun_out | grep ${tag}_ | tail -20
