"""Per-source-line stall samples / instruction counts of an .ncu-rep (needs -lineinfo).  usage: ncu_lines.py rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fpath = ""
agg = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": h = r; si = 6; ei = 7; continue
    if len(r) > 8 and r[0] not in ("", "Line No"):
        try: agg.append((int(r[6]), int(r[7]), fpath, int(r[0]), r[1].strip()[:110]))
        except ValueError: pass
ts = sum(a[0] for a in agg) or 1; ti = sum(a[1] for a in agg) or 1
print(f"total samples {ts}, warp instructions {ti}")
for s, n, f, ln, src in sorted(agg, reverse=True)[:top]:
    print(f"{100*s/ts:5.1f}% smp {100*n/ti:5.1f}% inst  {f}:{ln}  {src}")
