"""Per-CUDA-line profile: joins the SASS page of an .ncu-rep (samples, executed instructions, in program order)
with nvdisasm's line info for the same kernel.
usage: ncu_by_line.py <rep> <object-or-so> <mangled-kernel-name> [top]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, obj, fun = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
dis = ""
for c in cubin:
    d = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, c)], capture_output=True, text=True).stdout
    if f".text.{fun}:" in d:
        dis = d.split(f".text.{fun}:")[1]
        break
lines_of = []          # per instruction (program order): (addr, line)
cur = None
for ln in dis.splitlines():
    m = re.search(r'//## File ".*?", line (\d+)', ln)
    if m:
        cur = int(m.group(1)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        lines_of.append((int(m.group(1), 16), cur))
    if ln.startswith("//-----") or ln.startswith("\t.section"):
        if lines_of: break
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; ins = []
for r in rows:
    if hdr is None:
        if "Source" in r and "# Samples" in r: hdr = r
        continue
    if len(r) < len(hdr): continue
    try: ins.append((int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]), r[hdr.index("Source")]))
    except ValueError: pass
n = min(len(ins), len(lines_of))
print(f"# {len(ins)} profiled SASS instructions, {len(lines_of)} disassembled")
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
for (smp, ex, src), (addr, line) in zip(ins[:n], lines_of[:n]):
    a = agg[line]; a[0] += smp; a[1] += ex
    op = src.split()[1] if src.startswith("@") and len(src.split()) > 1 else src.split()[0]
    a[2][op.split(".")[0]] += ex
ts = sum(a[0] for a in agg.values()) or 1; te = sum(a[1] for a in agg.values()) or 1
srcfile = None
m = re.search(r'//## File "(.*?)"', dis)
text = open(m.group(1)).read().splitlines() if m and os.path.exists(m.group(1)) else []
for line, (smp, ex, ops) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    code = text[line - 1].strip()[:70] if line and line <= len(text) else ""
    mix = " ".join(f"{k}:{100*v//max(ex,1)}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
    print(f"{100*smp/ts:5.1f}% smp {100*ex/te:5.1f}% inst  L{line}  {code}   [{mix}]")
