#!/bin/bash
# One gpurun call for the round's record: tests, both bench arms, ncu launch list + full capture of the bench's own launch
# shape, developer phase profile.   usage: scripts/gpu_round.sh TAG
TAG=${1:-round}
mkdir -p gpurun_out; rm -f gpurun_out/rollout_divergence.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1; echo "host cores: $(nproc)" >> gpurun_out/${TAG}_gpu.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.log
cp gpurun_out/rollout_divergence.jsonl gpurun_out/${TAG}_rollout_divergence.jsonl 2>/dev/null
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-300
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json | cut -c1-400
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --long-window-ticks 0 > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "== ncu full"; timeout 1500 ncu --set full --clock-control none --import-source on -k regex:arena_step -s 5 -c 1 -o gpurun_out/${TAG}_full -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --long-window-ticks 0 > gpurun_out/${TAG}_full_log.txt 2>&1
echo "== phase profile"; RSB_PROF2=1 timeout 300 python -u scripts/phase_profile.py 4096 240 0 2 > gpurun_out/${TAG}_phase.log 2>&1; tail -14 gpurun_out/${TAG}_phase.log | cut -c1-100
