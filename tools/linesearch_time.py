#!/usr/bin/env python3
# coding: utf-8
"""Per-evaluation time of the attacks' line search (attacks/identical.py:68-77): 16 evaluations of
the rule on the same honest rows with a new Byzantine row each, with and without distance reuse."""
import json, pathlib, sys
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import byzantinemomentum_b200 as bz
from byzantinemomentum_b200 import engine
dev = torch.device("cuda", 0)
bz.config.strict_status = False
out = []
QUICK = "--quick" in sys.argv
REPS = 5
CASES = (("krum", 25, 5, 5, 1_310_922),) if QUICK else (("krum", 25, 5, 5, 1_310_922), ("bulyan", 25, 5, 5, 1_310_922), ("krum", 51, 12, 12, 1_310_922), ("krum", 11, 3, 3, 1_310_922),
                         ("krum", 25, 5, 5, 79_510), ("krum", 51, 12, 12, 79_510))
for gar, n, nb, f, d in CASES:
  gen = torch.Generator(device=dev).manual_seed(1)
  honest = [torch.randn(d, device=dev, generator=gen) for _ in range(n - nb)]
  attacks = [torch.randn(d, device=dev, generator=gen) for _ in range(17)]
  rec = dict(gar=gar, n=n, f=f, d=d)
  for reuse in (False, True):
    engine.config.reuse_distances = reuse
    best = None
    for rep in range(REPS):
      bz.gars[gar].unchecked(gradients=honest + [attacks[16]] * nb, f=f)          # the step's first call (fills the table)
      torch.cuda.synchronize()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for k in range(16):
        bz.gars[gar].unchecked(gradients=honest + [attacks[k]] * nb, f=f)
      b.record(); torch.cuda.synchronize()
      t = a.elapsed_time(b) / 16 * 1e3
      best = t if best is None else min(best, t)
    rec["reuse_us" if reuse else "full_us"] = best
  rec["ratio"] = rec["reuse_us"] / rec["full_us"]
  out.append(rec)
  print(json.dumps(rec), flush=True)
engine.config.reuse_distances = True
if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
  pathlib.Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
