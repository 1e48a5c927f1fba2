#!/usr/bin/env python
"""pcie_probe.py -- raw pinned-memory H2D / D2H / concurrent rates on this box (torch as plumbing)."""
import time
import torch
n = 1572864000 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
d2 = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): d.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    ms = t(fn)
    print("%s: %.2f ms  %.1f GB/s per direction" % (name, ms, 1.572864 / ms * 1e3))
