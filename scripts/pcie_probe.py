"""PCIe ceiling for the e2e arm: H2D alone, D2H alone, both at once (pinned 737 MB each, the C2 batch)."""
import torch, time
n = 90001 * 1024
h1 = torch.empty(n, dtype=torch.float64, pin_memory=True); h2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
d1 = torch.empty(n, dtype=torch.float64, device="cuda"); d2 = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
gb = n * 8 / 1e9
for name, fn in (("H2D", h2d), ("D2H", d2h), ("both", both)):
    s = t(fn)
    print(f"{name}: {s*1e3:.2f} ms, {gb/s:.1f} GB/s per direction", flush=True)
# chunked: 16 chunks alternating on 3 streams (what cb_fedvr_apply_host does, without the kernel)
for nchunk, nstream in ((23, 3), (23, 4), (46, 3), (12, 3), (23, 2)):
    streams = [torch.cuda.Stream() for _ in range(nstream)]
    per = (n + nchunk - 1) // nchunk
    bufs = [torch.empty(per, dtype=torch.float64, device="cuda") for _ in range(nstream)]
    def chunked():
        for c in range(nchunk):
            lo, hi = c * per, min(n, (c + 1) * per)
            s = streams[c % nstream]; b = bufs[c % nstream]
            with torch.cuda.stream(s):
                b[:hi - lo].copy_(h1[lo:hi], non_blocking=True)
                b[:hi - lo].mul_(1.0)
                h2[lo:hi].copy_(b[:hi - lo], non_blocking=True)
    s = t(chunked)
    print(f"chunked {nchunk} chunks on {nstream} streams (H2D -> kernel -> D2H per chunk): {s*1e3:.2f} ms, {gb/s:.1f} GB/s per direction", flush=True)
