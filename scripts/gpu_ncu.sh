#!/bin/bash
# ncu --set full capture of the tick kernel.  usage: scripts/gpu_ncu.sh TAG [ticks]
TAG=${1:-ncu}; TICKS=${2:-20}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:arena_step -s 3 -c 1 -o gpurun_out/${TAG} -f python bench.py --steps 1 --warmup 3 --ticks-per-step ${TICKS} --no-cpu-baseline > gpurun_out/${TAG}_log.txt 2>&1
tail -3 gpurun_out/${TAG}_log.txt
ls -la gpurun_out/${TAG}.ncu-rep
