#!/bin/bash
# Robustness record of the round: compute-sanitizer (memcheck, racecheck) on the mixed workload of scripts/dev_sanitize.py
# (Soccar 3v3, Snowday 2v2, Dropshot 2v2) and the soak runs of every game mode.   usage: scripts/gpu_extra.sh TAG
TAG=${1:-extra}
mkdir -p gpurun_out
echo "== memcheck"; timeout 1500 compute-sanitizer --tool memcheck python scripts/dev_sanitize.py 32 40 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/${TAG}_sanitizer_memcheck.log
echo "== racecheck"; timeout 1500 compute-sanitizer --tool racecheck python scripts/dev_sanitize.py 32 40 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/${TAG}_sanitizer_racecheck.log
: > gpurun_out/${TAG}_soak.jsonl
echo "== soak"
timeout 600 python scripts/soak.py 4096 36000 0 2 2>&1 | grep -E '^(\{|outlier)' | tee -a gpurun_out/${TAG}_soak.jsonl | cut -c1-300
for mc in "0 3" "3 3" "1 2" "2 1" "4 2"; do
  timeout 600 python scripts/soak.py 4096 12000 $mc 2>&1 | grep -E '^(\{|outlier)' | tee -a gpurun_out/${TAG}_soak.jsonl | cut -c1-300
done
