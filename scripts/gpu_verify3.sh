#!/bin/bash
# Final check of the build: full GPU test suite + smoke (+ the C1 wall clock).
tag=${1:-r02z}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${tag}_smoke.log
python scripts/c1_profile.py 600 > gpurun_out/${tag}_c1_wall.txt 2>&1; head -1 gpurun_out/${tag}_c1_wall.txt
