#!/usr/bin/env python3
# coding: utf-8
"""Back-to-back timing of the distance-based rules (development aid)."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import byzantinemomentum_b200 as bz
dev = torch.device("cuda", 0)
bz.config.strict_status = False
for n, f, d in ((25, 5, 1310922), (11, 3, 1310922), (11, 3, 79510), (51, 12, 79510), (51, 12, 1310922), (51, 12, 4568373), (25, 5, 36489290)):
  sets = max(1, min(6, -(-3 * 126 * 2**20 // (n * d * 4))))
  gen = torch.Generator(device=dev).manual_seed(3)
  stacks = [[torch.randn(d, device=dev, generator=gen) for _ in range(n)] for _ in range(sets)]
  line = []
  for gar in ("krum", "bulyan", "brute", "cge", "aksel"):
    if gar == "brute" and n > 11: continue
    ff = min(f, (n - 3) // 4) if gar == "bulyan" else f
    plans = [bz.Plan(gar, rows, f=ff) for rows in stacks]
    for k in range(5): plans[k % sets]()
    torch.cuda.synchronize()
    K = 100 if d < 5e6 else 20
    best = None
    for rep in range(3):
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for k in range(K): plans[k % sets]()
      b.record(); torch.cuda.synchronize()
      t = a.elapsed_time(b) / K * 1e3
      best = t if best is None else min(best, t)
    if d < 2e6:
      # the same call captured once in a CUDA graph (Plan.graph()): launch gaps gone
      replays = [pl.graph() for pl in plans]
      for k in range(5): replays[k % sets]()
      torch.cuda.synchronize()
      gbest = None
      for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k in range(K): replays[k % sets]()
        b.record(); torch.cuda.synchronize()
        t = a.elapsed_time(b) / K * 1e3
        gbest = t if gbest is None else min(gbest, t)
      line.append("%s %.1f (graph %.1f)" % (gar, best, gbest))
    else:
      line.append("%s %.1f" % (gar, best))
  print("n=%d f=%d d=%d : " % (n, f, d) + "  ".join(line), flush=True)
  del stacks, plans
  torch.cuda.empty_cache()
