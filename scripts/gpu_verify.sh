#!/bin/bash
# Verification visit (1 GPU): full GPU test suite, smoke, both bench arms at their defaults, the C1 wall clock, the C5 stress shape.
tag=${1:-r02v}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
python scripts/c1_profile.py 600 > gpurun_out/${tag}_c1_wall.txt 2>&1; head -1 gpurun_out/${tag}_c1_wall.txt
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "reference arm rc=$?"
python bench.py --workload c5 --steps 200 --warmup 10 --no-cpu --loop-steps 20 --fp32-loop-steps 0 --overlap-loop 0 --host-fed-steps 0 --c2-steps 0 --c1-iters 0 > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err; echo "c5 rc=$?"
python - <<PY
import json
for f in ("bench_n1", "bench_reference_arm", "bench_c5"):
    try:
        d = json.loads(open("gpurun_out/${tag}_%s.json" % f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, d.get("value"), d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"))
    for k in ("learner", "learner_fp32"):
        if d.get(k): print("  ", k, d[k]["value"], d[k].get("segments_ms"))
    if d.get("c1_dropin"): print("   c1", d["c1_dropin"]["ms_per_call"], d["c1_dropin"]["ms_per_iteration"])
    if d.get("python_reference"): print("   pyref c1", (d["python_reference"].get("c1_loop") or {}).get("ms_per_call"))
    if d.get("replay"): print("   replay", d["replay"]["sample"]["frac_of_hbm_peak"], d["replay"]["push"]["frac_of_hbm_peak"])
PY
