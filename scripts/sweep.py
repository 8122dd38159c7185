"""Times one kernel family for every variants/lib*.so (development aid).  usage: sweep.py <what> [name-filter]
what: fedvr | banded | banded3 | eval | asm3 | density"""
import os, subprocess, sys, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
what = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
libs = [os.path.join(ROOT, "compactbases_jl_b200", "libcompactbases_cuda.so")] + sorted(glob.glob(os.path.join(ROOT, "variants", "lib*.so")))
for lib in libs:
    if flt and flt not in os.path.basename(lib) and "variants" in lib:
        continue
    env = dict(os.environ, CB_CUDA_LIB=lib, PROF_TIME="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "prof_one.py"), what, "30"], env=env, capture_output=True, text=True, timeout=300)
    last = (r.stdout.strip().splitlines() or ["(no output) " + r.stderr[-300:]])[-1]
    print(f"{os.path.basename(lib):40s} {last}", flush=True)
