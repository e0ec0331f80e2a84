#!/usr/bin/env python3
# coding: utf-8
"""Timing + parity of Multi-Krum for many rows (n > 36); run once per K2 variant
(env BYZAGG_K2_BLOCKED=1 selects the row-blocked kernel).  Development aid."""
import os, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import byzantinemomentum_b200 as bz
dev = torch.device("cuda", 0)
tag = "blocked" if os.environ.get("BYZAGG_K2_BLOCKED") else "current"
CASES = ((51, 12, 4568373), (64, 15, 4568373), (40, 9, 1000003)) if os.environ.get("BYZAGG_K2_BLOCKED") else ((25, 5, 1310922), (11, 3, 1310922), (25, 5, 36489290), (30, 7, 4568373), (36, 8, 4568373), (18, 4, 1310922))
for n, f, d in CASES:
  gen = torch.Generator(device=dev).manual_seed(3)
  rows = [torch.randn(d, device=dev, generator=gen) for _ in range(n)]
  part = bz.engine.pairdist_partial(rows)
  x = torch.stack(rows).double()
  ref = torch.cdist(x[:, :200000], x[:, :200000]) if False else None
  # exact check on a column sample: squared distances in fp64
  sq = torch.zeros(n, n, dtype=torch.float64, device=dev)
  for lo in range(0, d, 1 << 20):
    blk = x[:, lo:lo + (1 << 20)]
    g = blk @ blk.T
    nrm = (blk * blk).sum(1)
    sq += nrm[:, None] + nrm[None, :] - 2 * g
  iu = torch.triu_indices(n, n, 1, device=dev)
  err = ((part[iu[0], iu[1]] - sq[iu[0], iu[1]]).abs() / sq[iu[0], iu[1]]).max().item()
  del x, sq
  plan = bz.Plan("krum", rows, f=f)
  for _ in range(3): plan()
  torch.cuda.synchronize()
  best = None
  for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(10): plan()
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 10 * 1e3
    best = t if best is None else min(best, t)
  ws_t = []
  for rep in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); bz.engine.pairdist_partial(rows); b.record(); torch.cuda.synchronize()
    ws_t.append(a.elapsed_time(b) * 1e3)
  print("%s n=%d f=%d d=%d krum %.1f us  pairdist_partial %.1f us  max rel err %.2e" % (tag, n, f, d, best, min(ws_t), err), flush=True)
  del rows, plan
  torch.cuda.empty_cache()
