#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench, ncu launch list.  usage: scripts/gpu_check.sh [tag]
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/${TAG}_bench.json
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --ticks-per-step 20 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -5 gpurun_out/${TAG}_launches.csv
