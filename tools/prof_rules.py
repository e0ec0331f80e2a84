#!/usr/bin/env python3
# coding: utf-8
"""Run every rule a few times on one shape (to be wrapped in `ncu --metrics gpu__time_duration.sum`)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import byzantinemomentum_b200 as bz
n, f, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rules = sys.argv[4].split(",") if len(sys.argv) > 4 else ["median", "trmean", "phocas", "krum", "bulyan", "aksel", "cge", "average"]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
rows = [torch.randn(d, device=dev, generator=gen) for _ in range(n)]
bz.config.strict_status = False
for gar in rules:
  ff = min(f, (n - 3) // 4) if gar == "bulyan" else min(f, (n - 3) // 2) if gar == "krum" else f
  plan = bz.Plan(gar, rows, f=ff)
  for _ in range(3):
    plan()
torch.cuda.synchronize()
