#!/bin/bash
# Iteration call: GPU parity tests (optional), bench, developer phase profile.   usage: scripts/gpu_iter.sh TAG [notest]
TAG=${1:-it}
mkdir -p gpurun_out
if [ "$2" != "notest" ]; then
  echo "== pytest gpu"; timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest.log
fi
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json | cut -c1-600
echo "== phase profile"; RSB_PROF2=1 timeout 300 python -u scripts/phase_profile.py 4096 240 0 2 > gpurun_out/${TAG}_phase.log 2>&1; tail -14 gpurun_out/${TAG}_phase.log | cut -c1-150
