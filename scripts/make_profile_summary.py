"""Turns the raw GPU-box artefacts of one round (gpurun_out/) into the tracked summaries under profiles/.
usage: python scripts/make_profile_summary.py TAG ROUND   (reads gpurun_out/TAG_*, writes profiles/rNN_*)"""
import csv, json, os, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
os.makedirs(P, exist_ok=True)
out = {}
# 1. launch list (ncu --metrics gpu__time_duration.sum): per-kernel totals and the tick kernel's share
rows = [r for r in csv.reader(open(os.path.join(G, f"{tag}_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
per = {}
for r in rows:
    name = r[4].split("(")[0]
    per.setdefault(name, []).append(float(r[-1]))
tot = sum(sum(v) for v in per.values())
out["launch_list"] = {k: {"launches": len(v), "total_ms": sum(v) / 1e6, "share": sum(v) / tot} for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1]))}
with open(os.path.join(P, f"{rnd}_launches.csv"), "w") as f:
    f.write(open(os.path.join(G, f"{tag}_launches.csv")).read())
# 2. ncu --set full of one tick-kernel launch
rep = os.path.join(G, f"{tag}_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, vals = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
        "lts__t_bytes.sum", "sass__inst_executed_local_loads", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__cycles_elapsed.max"]
m = {}
for i, h in enumerate(hdr):
    if h in want:
        m[h] = {"value": vals[i], "unit": units[i]}
stalls = {}
for i, h in enumerate(hdr):
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try: stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(vals[i])
        except ValueError: pass
ts = sum(stalls.values()) or 1
out["ncu_full"] = {"kernel": "arena_step_kernel", "metrics": m, "stall_share": {k: round(v / ts, 4) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]}}
def num(k): return float(m[k]["value"].replace(",", "")) if k in m else None
unit_scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tr = 0.0
for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    tr += num(k) * unit_scale.get(m[k]["unit"], 1)
out["ncu_full"]["dram_traffic_bytes_per_launch"] = tr
# 3. bench lines of the same call
for name in ("bench", "bench_ref"):
    p = os.path.join(G, f"{tag}_{name}.json")
    if os.path.exists(p):
        lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
        if lines: out[name] = json.loads(lines[-1])
json.dump(out, open(os.path.join(P, f"{rnd}_summary.json"), "w"), indent=1)
for extra in ("phase.log", "rollout_divergence.jsonl", "gpu.txt"):
    p = os.path.join(G, f"{tag}_{extra}")
    if os.path.exists(p):
        open(os.path.join(P, f"{rnd}_{extra}"), "w").write(open(p).read())
print(json.dumps({k: out[k] for k in ("launch_list",)}, indent=1)[:1500])
print("traffic/launch", tr, "kernel ms", m.get("gpu__time_duration.sum"))
