#!/usr/bin/env python3
# coding: utf-8
"""Where the time of a host-tensor call goes on THIS box (development aid): each piece timed alone."""
import pathlib, sys, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import byzantinemomentum_b200 as bz
dev = torch.device("cuda", 0)
n, d = 25, 1_310_922
rows = [torch.randn(d).pin_memory() for _ in range(n)]
def tm(label, fn, k=10):
  fn(); torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(k): fn()
  torch.cuda.synchronize()
  print("%-44s %8.3f ms" % (label, (time.perf_counter() - t0) / k * 1e3), flush=True)
tm("is_pinned() x 25", lambda: [r.is_pinned() for r in rows])
buf = torch.empty((n, (d + 63) // 64 * 64), device=dev)
tm("25 copy_ (1 stream)", lambda: [buf[i, :d].copy_(r, non_blocking=True) for i, r in enumerate(rows)])
out = torch.empty(d, device=dev)
cached = torch.empty(d).pin_memory()
tm("D2H into cached pinned + sync", lambda: (cached.copy_(out, non_blocking=True), torch.cuda.current_stream().synchronize()))
tm("clone() of a 5 MB CPU tensor", lambda: cached.clone())
tm("torch.empty(d, pin_memory=True)", lambda: torch.empty(d, dtype=torch.float32, pin_memory=True))
tm("torch.empty(d) pageable + fill", lambda: torch.empty(d).zero_())
tm("gars['trmean'] on the 25 host rows (whole call)", lambda: bz.gars["trmean"].unchecked(gradients=rows, f=10), 12)
print(bz.engine.host_path_report(0))
