#!/bin/bash
# ncu --set full capture of one launch of the tick kernel, taken after SKIP earlier launches so that the arenas are
# in mid-game states.  usage: scripts/gpu_ncu.sh TAG [ticks_per_launch] [skip_launches]
TAG=${1:-ncu}; TICKS=${2:-60}; SKIP=${3:-5}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:arena_step -s ${SKIP} -c 1 -o gpurun_out/${TAG} -f python bench.py --steps 1 --warmup ${SKIP} --ticks-per-step ${TICKS} --no-cpu-baseline > gpurun_out/${TAG}_log.txt 2>&1
tail -2 gpurun_out/${TAG}_log.txt | cut -c1-300
ls -la gpurun_out/${TAG}.ncu-rep
