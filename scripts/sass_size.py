"""Developer tool: static SASS size of the tick kernel attributed to source functions (inclusive over the inline
chain) -- finds what inlining blows the instruction footprint up.  usage: sass_size.py LIB.so [kernel-substring]"""
import os, re, subprocess, sys, tempfile, collections, glob
lib = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else "arena_step_kernelILi0E"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout.splitlines()
_src = {}
def func_of(f, line):
    if f not in _src:
        try: _src[f] = open(os.path.join(ROOT, f)).read().splitlines()
        except Exception: _src[f] = None
    s = _src[f]
    if s is None: return "?"
    for i in range(min(line, len(s)) - 1, -1, -1):
        m = re.match(r"^(?:RSB_HD(?:_NOINLINE)?|static|template.*|__global__|__device__|inline)\s.*?(\w+)\s*\(", s[i])
        if m and not s[i].startswith("\t"): return m.group(1)
    return "?"
sections = collections.Counter(); incl = collections.Counter(); excl = collections.Counter()
sec = None; chain = []; fresh = True
for ln in dis:
    if ln.startswith(".text."):
        sec = ln.strip().rstrip(":"); continue
    if ln.startswith("\t.section") or ln.startswith(".section"):
        sec = None
    if sec is None: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        fr = (os.path.relpath(m.group(1), ROOT), int(m.group(2)))
        if fresh: chain = [fr]; fresh = False
        elif fr != chain[-1]: chain.append(fr)
        continue
    if re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", ln):
        sections[sec] += 1
        if kern in sec or "rsb" in sec:
            seen = set()
            for k, fr in enumerate(chain):
                fn = func_of(*fr)
                if k == 0: excl[(sec[:40], fn)] += 1
                if fn not in seen: seen.add(fn); incl[fn] += 1
        fresh = True
print("== sections (instructions)")
for s, n in sections.most_common(25): print(f"  {n:8d}  {s[:110]}")
print("== inclusive by function")
for f, n in incl.most_common(60): print(f"  {n:8d}  {f}")
