#!/bin/bash
# GPU visit: env C3 with 4 resident CTAs (tests + bench + phase trace), conv epilogue A/B, remaining new tests.
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py tests/test_multi_layout_gpu.py tests/test_checkpoint_gpu.py tests/test_agent_gpu.py -m gpu -q --maxfail=20 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 300 --warmup 20 --no-learner --no-cpu --replay-batch 0 > gpurun_out/${tag}_bench_env.json 2> gpurun_out/${tag}_bench_env.err; echo "bench rc=$?"
python -c "
import json,sys
d=json.loads(open('gpurun_out/${tag}_bench_env.json').read().strip().splitlines()[-1])
print('C3 env', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'wire', d['e2e']['wire']['value'], 'c2', d['secondary_c2']['value'], d['secondary_c2']['ms_per_step'])
"
timeout 300 python scripts/step_time_trace.py c3 400 > gpurun_out/${tag}_step_trace_c3.txt 2>&1; tail -22 gpurun_out/${tag}_step_trace_c3.txt
MARL_B200_SO=$PWD/dqn_marl_b200/libmarl_b200_envtrace.so timeout 300 python scripts/env_phase_trace.py c3 200 > gpurun_out/${tag}_phase_c3.txt 2>&1; cat gpurun_out/${tag}_phase_c3.txt
timeout 600 python scripts/qnet_ab.py 0 1 2 4 8 9 3 15 > gpurun_out/${tag}_qnet_ab.txt 2>&1; cat gpurun_out/${tag}_qnet_ab.txt
