#!/bin/bash
# Multi-GPU lines of BASELINE.json's configs, produced by bench.py itself under torchrun (one rank per GPU, weak scaling, no
# collective on the data path).  usage: scripts/gpu_multi.sh N TAG "name:mode:cpt:arenas_per_gpu ..."   (run under gpurun --gpus N)
N=$1; TAG=${2:-multi}; CASES=$3
OUT=gpurun_out/${TAG}_n${N}.jsonl; : > $OUT
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c > gpurun_out/${TAG}_n${N}_gpus.txt; echo "host cores: $(nproc)" >> gpurun_out/${TAG}_n${N}_gpus.txt
for c in $CASES; do
  IFS=: read name mode cpt arenas <<< "$c"
  echo "== $name (N=$N, $arenas arenas/GPU)"
  timeout 300 python bench.py --impl reference --gpus $N --mode $mode --cars-per-team $cpt --arenas $arenas --steps 3 --warmup 1 2>&1 | tail -1 | sed "s/^{/{\"case\": \"$name\", /" >> $OUT
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --mode $mode --cars-per-team $cpt --arenas $arenas --steps 6 --warmup 3 --no-cpu-baseline --long-window-ticks 0 2>&1 | tail -1 | sed "s/^{/{\"case\": \"$name\", /" >> $OUT
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --mode $mode --cars-per-team $cpt --arenas $arenas --steps 6 --warmup 3 --no-cpu-baseline --long-window-ticks 0 2>&1 | grep '^{' | tail -1 | sed "s/^{/{\"case\": \"$name\", /" >> $OUT
  fi
done
python - <<PY
import json
for l in open("$OUT"):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:200]); continue
    print(d.get("case"), d.get("impl", "cuda"), "n_gpus", d.get("n_gpus"), f"{d['value']/1e6:.2f} M arena-ticks/s", "e2e %.2f M" % (d['e2e']['value']/1e6), "err", d.get("device_error_flags"), (d.get("cpu_baseline") or {}).get("cores"))
PY
