"""Developer tool: join an ncu SASS source-page CSV with nvdisasm line info and rank source lines /
functions by warp-stall samples.  usage: ncu_hotspots.py REP.ncu-rep LIB.so [kernel-substring]"""
import csv, os, re, subprocess, sys, tempfile, collections, bisect, glob
rep, lib = sys.argv[1], sys.argv[2]
kern = os.environ.get("RSB_HOTSPOT_KERNEL", "arena_step_kernelILi0E")  # the Soccar instantiation (the bench workload); one .text section per game mode
focus = sys.argv[3] if len(sys.argv) > 3 else None  # only count instructions whose inline chain passes through this function
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout.splitlines()
# instruction offsets -> innermost source (file, line) for the kernel's text section
infn, cur, offs, fresh, chain = False, None, {}, True, []
for ln in dis:
    if ln.startswith(".text."):
        infn = kern in ln
        continue
    if ln.startswith("\t.section") or ln.startswith(".section"):
        infn = False
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        fr = (os.path.relpath(m.group(1), ROOT), int(m.group(2)))
        if fresh:  # frames are listed innermost first
            cur = fr
            chain = [fr]
            fresh = False
        elif fr != chain[-1]:
            chain.append(fr)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        offs[int(m.group(1), 16)] = (cur, m.group(2), tuple(chain))
        fresh = True
# function attribution: nearest preceding 'RSB_HD ... name(' line in the file
def func_of(f, line):
    if f not in _src_cache:
        try:
            _src_cache[f] = open(os.path.join(ROOT, f)).read().splitlines()
        except Exception:
            _src_cache[f] = None
    src = _src_cache[f]
    if src is None:
        return "?"
    for i in range(min(line, len(src)) - 1, -1, -1):
        m = re.match(r"^(?:RSB_HD(?:_NOINLINE)?|static|template.*|__global__|__device__|inline)\s.*?(\w+)\s*\(", src[i])
        if m and not src[i].startswith("\t"):
            return m.group(1)
    return "?"
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdr = rows[1]
ia, isamp, iinst, ithr = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
by_line, by_file = collections.Counter(), collections.Counter()
incl, incl_inst = collections.Counter(), collections.Counter()
_src_cache = {}
inst_line, stall_kind = collections.Counter(), collections.defaultdict(collections.Counter)
tot = 0; tot_inst = 0
for r in rows[2:]:
    if len(r) <= isamp or not r[ia].startswith("0x"):
        continue
    addr = int(r[ia], 16)
    if base is None:
        base = addr
    ent = offs.get(addr - base, (None, "?", ()))
    src = ent[0] or ("?", 0)
    if focus and not any(func_of(fr[0], fr[1]) == focus for fr in ent[2]):
        continue
    n = int(r[isamp] or 0); tot += n
    seen = set()
    for fr in ent[2]:
        key = (fr[0], func_of(fr[0], fr[1]))
        if key not in seen:
            seen.add(key); incl[key] += n; incl_inst[key] += int(r[iinst] or 0)
    by_line[src] += n; by_file[src[0]] += n
    inst_line[src] += int(r[iinst] or 0); tot_inst += int(r[iinst] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            stall_kind[src][hdr[i]] += v
print(f"total samples {tot}, warp instructions {tot_inst}")
_st = collections.Counter()
for k, c in stall_kind.items():
    for name, v in c.items():
        _st[name] += v
_stt = sum(_st.values())
print("== stall reasons (share of samples): " + ", ".join(f"{k[6:]} {100 * v / max(_stt, 1):.1f}%" for k, v in _st.most_common(12)))
print("== by file"); [print(f"  {100*n/tot:5.1f} %  {f}") for f, n in by_file.most_common(12)]
by_fn, inst_fn = collections.Counter(), collections.Counter()
for (f, l), n in by_line.items():
    fn = func_of(f, l); by_fn[(f, fn)] += n; inst_fn[(f, fn)] += inst_line[(f, l)]
print("== by function (samples %, warp-instr %)")
for (f, fn), n in by_fn.most_common(40):
    print(f"  {100*n/tot:5.1f} %  {100*inst_fn[(f, fn)]/max(tot_inst,1):5.1f} %  {f}:{fn}")
print("== inclusive by function (samples %, warp-instr %)")
for (f, fn), n in incl.most_common(45):
    print(f"  {100*n/tot:5.1f} %  {100*incl_inst[(f, fn)]/max(tot_inst,1):5.1f} %  {f}:{fn}")
print("== top lines")
for (f, l), n in by_line.most_common(40):
    top = ", ".join(f"{k[6:]}={v}" for k, v in stall_kind[(f, l)].most_common(3))
    try:
        text = open(os.path.join(ROOT, f)).read().splitlines()[l - 1].strip()[:90]
    except Exception:
        text = ""
    print(f"  {100*n/tot:5.1f} %  {f}:{l}  [{top}]  {text}")
