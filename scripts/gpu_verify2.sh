#!/bin/bash
tag=${1:-r02w}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_agent_gpu.py tests/test_abi.py tests/test_bench_contract.py -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
print(d.get("value"), d.get("ms_per_step"), "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
for k in ("learner", "learner_fp32"):
    print("  ", k, d[k]["value"], d[k].get("segments_ms"))
print("   c1", d["c1_dropin"]["ms_per_call"], d["c1_dropin"]["ms_per_iteration"])
print("   replay", d["replay"]["sample"]["frac_of_hbm_peak"], d["replay"]["push"]["frac_of_hbm_peak"], "c2", d["secondary_c2"]["value"], d["secondary_c2"]["ms_per_step"])
PY
