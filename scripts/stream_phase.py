"""Phase clocks of the streaming assembly kernel (needs a -DSTREAM_PROF build: scripts/build_variant.sh)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import compactbases_jl_b200 as cb
from compactbases_jl_b200 import _lib
from compactbases_jl_b200.operators import _ptr, _stream
lib = _lib.load()
B = cb.BSpline(cb.ExpKnotSet(10, -3.0, 2.0, 200_000))
V = torch.empty((B.nf, 19), dtype=torch.float64, device="cuda")
out = (C.c_ulonglong * 8)()
for _ in range(3):
    _lib.call("cb_bspline_assemble", B.handle, B.handle, 1, 1, None, 0, B.nf, 0, B.nf, 9, 9, _ptr(V), _stream())
torch.cuda.synchronize()
lib.cb_debug_stream_prof(out, 1)
_lib.call("cb_bspline_assemble", B.handle, B.handle, 1, 1, None, 0, B.nf, 0, B.nf, 9, 9, _ptr(V), _stream())
torch.cuda.synchronize()
lib.cb_debug_stream_prof(out, 0)
steps = out[2]
print("CTA-steps", steps)
print("thread 0 per step [cycles]: segA own %.0f, segA wait %.0f, segB own %.0f, segB wait %.0f" %
      (out[3] / steps, out[0] / steps, out[4] / steps, out[1] / steps))
