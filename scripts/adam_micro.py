"""Steady-state microbenchmark of the table Adam pass (every row touched): ms and GB/s of the 34 B/param algorithmic traffic."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seald_nerf_b200 import _lib
from seald_nerf_b200._lib import ptr
d = torch.device("cuda", 0)
n = 12239728
p = torch.randn(n, device=d); m = torch.randn(n, device=d) * 1e-3; v = torch.rand(n, device=d) * 1e-6
g0 = torch.randn(n, device=d)
g = g0.clone()
p16 = torch.empty(n, dtype=torch.float16, device=d)
step = torch.zeros(1, dtype=torch.int32, device=d); scale = torch.ones(1, device=d); found = torch.zeros(1, dtype=torch.int32, device=d)
cap = [0]
def run():
    _lib.call("seald_adam_step_ex", ptr(p), ptr(g), ptr(m), ptr(v), n, 1e-3, 0.9, 0.99, 1e-15, 1, ptr(step), ptr(scale), ptr(found), ptr(p16), 1, cap[0],
              _lib.stream())
for c in (0, 444, 592, 888, 1776, 2368, 4736):
  cap[0] = c
  print("grid cap", c)
  for frac in (1.0,):
    ts = []
    for it in range(12):
        g.copy_(g0)
        if frac < 1.0:
            g[int(n * frac):] = 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    print("  ", "adam table pass, %.0f%% of the gradient non-zero (all moments non-zero): %.4f ms, %.0f GB/s algorithmic (34 B/param)" % (100 * frac, ms, 34.0 * n / ms / 1e6))
