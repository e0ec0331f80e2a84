#!/usr/bin/env python3
# coding: utf-8
"""In-stream time of each phase of the distance rules (development aid)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from byzantinemomentum_b200 import engine
dev = torch.device("cuda", 0)
n, f, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
gen = torch.Generator(device=dev).manual_seed(3)
stacks = [[torch.randn(d, device=dev, generator=gen) for _ in range(n)] for _ in range(4)]
def timeit(fn, K=100):
  for k in range(5): fn(k)
  torch.cuda.synchronize()
  best = None
  for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(K): fn(k)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / K * 1e3
    best = t if best is None else min(best, t)
  return best
m = n - f - 2
part = engine.pairdist_partial(stacks[0]).unsqueeze(0).contiguous()
order = engine.krum_select(part, n, f)
print("pairdist_partial (K2 + reduce)", timeit(lambda k: engine.pairdist_partial(stacks[k % 4])))
print("krum_select (K5, 1 part)      ", timeit(lambda k: engine.krum_select(part, n, f)))
print("average_selected (K3, m=%d)    " % m, timeit(lambda k: engine.average_selected(stacks[k % 4], order, m)))
fb = min(f, (n - 3) // 4)
ob, st = engine.bulyan_select(part, n, fb, n - fb - 2)
print("bulyan_reduce (K4)            ", timeit(lambda k: engine.bulyan_reduce(stacks[k % 4], fb, n - fb - 2, ob, st)))
print("rowdist_partial (K2')         ", timeit(lambda k: engine.rowdist_partial(stacks[k % 4])))
pn = engine.rowdist_partial(stacks[0]).unsqueeze(0).contiguous()
print("rowdist_select                ", timeit(lambda k: engine.rowdist_select(pn, n, True)))
print("krum whole                    ", timeit(lambda k: engine.krum(stacks[k % 4], f, m)))
print("empty torch.empty + ctypes est", timeit(lambda k: torch.empty(16, device=dev)))
