#!/bin/bash
# BASELINE.json configs 3-5 at their per-GPU sizes (parity-test cases, not the headline): CUDA arm + reference arm each.
# usage: scripts/gpu_configs.sh TAG
TAG=${1:-cfg}; OUT=gpurun_out/${TAG}_configs.jsonl; : > $OUT
run() {  # name mode cars_per_team arenas
  echo "== $1"
  timeout 300 python bench.py --impl reference --mode $2 --cars-per-team $3 --arenas $4 --steps 3 --warmup 1 2>&1 | tail -1 | sed "s/^{/{\"case\": \"$1\", /" >> $OUT
  timeout 600 python bench.py --mode $2 --cars-per-team $3 --arenas $4 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | sed "s/^{/{\"case\": \"$1\", /" >> $OUT
}
run soccar_3v3_8192 0 3 8192
run hoops_2v2_16384 1 2 16384
run heatseeker_1v1_16384 2 1 16384
run snowday_3v3_4096 3 3 4096
run snowday_3v3_32768 3 3 32768
run soccar_2v2_65536 0 2 65536
python - <<PY
import json
for l in open("$OUT"):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:200]); continue
    print(d.get("case"), d.get("impl", "cuda"), f"{d['value']/1e6:.2f} M arena-ticks/s", "e2e %.2f M" % (d['e2e']['value']/1e6), "err", d.get("device_error_flags"))
PY
