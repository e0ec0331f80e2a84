#!/usr/bin/env python3
# coding: utf-8
"""Where does the end-to-end (host tensors) path spend its time? (development aid)"""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import byzantinemomentum_b200 as bz
dev = torch.device("cuda", 0)
n, d = 25, 1310922
def tm(fn, k=10):
  fn(); torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(k): fn()
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / k * 1e3
big = torch.empty(n * d, dtype=torch.float32).pin_memory()
dst = torch.empty(n * d, device=dev)
print("one 131 MB pinned copy        : %.2f ms" % tm(lambda: dst.copy_(big, non_blocking=True)))
rows_view = [big[i * d:(i + 1) * d] for i in range(n)]
buf = torch.empty((n, (d + 63) // 64 * 64), device=dev)
def copy_rows(rows):
  for k, g in enumerate(rows):
    buf[k, :d].copy_(g, non_blocking=True)
print("25 copies, views of one block : %.2f ms" % tm(lambda: copy_rows(rows_view)))
rows_sep = [torch.randn(d).pin_memory() for _ in range(n)]
print("25 copies, separate pinned    : %.2f ms" % tm(lambda: copy_rows(rows_sep)))
rows_pageable = [torch.randn(d) for _ in range(n)]
print("25 copies, pageable           : %.2f ms" % tm(lambda: copy_rows(rows_pageable)))
print("gars call, separate pinned    : %.2f ms" % tm(lambda: bz.gars["trmean"].unchecked(gradients=rows_sep, f=10)))
print("gars call, views of one block : %.2f ms" % tm(lambda: bz.gars["trmean"].unchecked(gradients=rows_view, f=10)))
out = torch.empty(d, device=dev); pin = torch.empty(d).pin_memory()
print("D2H 5 MB pinned + clone       : %.2f ms" % tm(lambda: (pin.copy_(out, non_blocking=True), torch.cuda.current_stream().synchronize(), pin.clone())))
