"""Hottest CUDA source lines of an .ncu-rep (needs -lineinfo and --import-source on): stall samples and executed
warp instructions per line.  usage: ncu_cuda_lines.py rep [kernel-index] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
lines = []
for r in rows:
    if hdr is None:
        if "Source" in r and "# Samples" in r:
            hdr = r
        continue
    if len(r) < len(hdr):
        continue
    try:
        smp = int(r[hdr.index("# Samples")]); ex = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    ln = r[hdr.index("#")] if "#" in hdr else ""
    lines.append((smp, ex, ln, r[hdr.index("Source")][:110]))
ts = sum(l[0] for l in lines) or 1
te = sum(l[1] for l in lines) or 1
print(f"total samples {ts}, warp instructions {te}")
for smp, ex, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100*smp/ts:5.1f}% smp {100*ex/te:5.1f}% inst  L{ln:>4}  {src}")
