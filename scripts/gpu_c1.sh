#!/bin/bash
# C1 facade latency visit: GPU tests, wall clock per call (warp-per-env variants A/B), host profile, launch list of a few iterations.
tag=${1:-r02c1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
python scripts/c1_profile.py 500 > gpurun_out/${tag}_wall.txt 2>&1; echo "wall rc=$?"; head -1 gpurun_out/${tag}_wall.txt
for w in 2 4; do
  MQ_SMALL_WPE=$w python scripts/c1_profile.py 500 > gpurun_out/${tag}_wall_wpe$w.txt 2>&1; echo "wpe=$w rc=$?"; head -1 gpurun_out/${tag}_wall_wpe$w.txt
done
python scripts/c1_profile.py 500 --cprofile > gpurun_out/${tag}_cprofile.txt 2>&1; echo "cprofile rc=$?"
ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${tag}_launches.csv python scripts/c1_profile.py 44 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/${tag}_launches.csv 2>/dev/null | head -45
