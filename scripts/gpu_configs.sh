#!/bin/bash
# BASELINE.json configs 3-5 at their per-GPU sizes (parity-test cases, not the headline): CUDA arm + reference arm each.
# usage: scripts/gpu_configs.sh TAG
TAG=${1:-cfg}; OUT=gpurun_out/${TAG}_configs.jsonl; : > $OUT
run() {  # name mode cars_per_team arenas
  echo "== $1"
  timeout 300 python bench.py --impl reference --mode $2 --cars-per-team $3 --arenas $4 --steps 3 --warmup 1 2>&1 | tail -1 | sed "s/^{/{\"case\": \"$1\", /" >> $OUT
  timeout 600 python bench.py --mode $2 --cars-per-team $3 --arenas $4 --steps 5 --warmup 3 --no-cpu-baseline --long-window-ticks 0 2>&1 | tail -1 | sed "s/^{/{\"case\": \"$1\", /" >> $OUT
}
run soccar_2v2_8192 0 2 8192       # per-GPU share of the north-star size (65,536 arenas on 8 GPUs)
run soccar_3v3_8192 0 3 8192       # per-GPU share of config 3 (65,536 on 8 GPUs)
run hoops_2v2_16384 1 2 16384      # config 4
run heatseeker_1v1_16384 2 1 16384 # config 4
run snowday_3v3_4096 3 3 4096      # config 5 at 8 GPUs (per-GPU share of 32,768)
run snowday_3v3_32768 3 3 32768    # config 5 on one GPU
run dropshot_2v2_4096 4 2 4096     # Dropshot (not a BASELINE config; same shape as the headline)
echo "== config 1 (THE_VOID 1v1, 1 M ticks)"
timeout 900 python scripts/config1_void.py 1000000 4 2>&1 | tail -1 > gpurun_out/${TAG}_config1_void.json; cut -c1-400 gpurun_out/${TAG}_config1_void.json
python - <<PY
import json
for l in open("$OUT"):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:200]); continue
    print(d.get("case"), d.get("impl", "cuda"), f"{d['value']/1e6:.2f} M arena-ticks/s", "e2e %.2f M" % (d['e2e']['value']/1e6), "err", d.get("device_error_flags"))
PY
