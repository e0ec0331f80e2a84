#!/usr/bin/env python3
# coding: utf-8
"""krum / bulyan with f aliased Byzantine rows (the reference's attack pattern) vs distinct rows."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import byzantinemomentum_b200 as bz
dev = torch.device("cuda", 0)
bz.config.strict_status = False
n, f, d = 25, 5, 1310922
def stacks(alias):
  out = []
  for s in range(4):
    gen = torch.Generator(device=dev).manual_seed(s)
    honest = [torch.randn(d, device=dev, generator=gen) for _ in range(n - f)]
    if alias:
      byz = torch.stack(honest).mean(dim=0).mul(-1.1)
      out.append(honest + [byz] * f)
    else:
      out.append(honest + [torch.randn(d, device=dev, generator=gen) for _ in range(f)])
  return out
for alias in (False, True):
  st = stacks(alias)
  for gar in ("krum", "bulyan"):
    plans = [bz.Plan(gar, rows, f=f) for rows in st]
    for k in range(5): plans[k % 4]()
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for k in range(100): plans[k % 4]()
      b.record(); torch.cuda.synchronize()
      t = a.elapsed_time(b) / 100 * 1e3
      best = t if best is None else min(best, t)
    print(f"{gar:7s} n={n} f={f} aliased={alias}: {best:.1f} us", flush=True)
