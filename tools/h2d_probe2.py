#!/usr/bin/env python3
# coding: utf-8
"""Where do pinned host buffers have to live, and how should they be copied, for the host->device
leg of the e2e path to run at the box's PCIe rate?  (development aid; one GPU)
For every NUMA node (thread bound to the node's CPUs while the pinned buffers are allocated and first
touched), for the NVML 'ideal affinity' binding and for no binding: H2D GB/s of 25 x 5.2 MB rows with
1, 2 and 4 copy streams, and of one contiguous 131 MB copy."""
import json, os, pathlib, sys, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from byzantinemomentum_b200 import hostmem
dev = torch.device("cuda", 0)
n, d = 25, 1_310_922

def nodes():
  out = {}
  base = pathlib.Path("/sys/devices/system/node")
  for p in sorted(base.glob("node[0-9]*")):
    cpus = set()
    for part in (p / "cpulist").read_text().strip().split(","):
      if not part: continue
      a, _, b = part.partition("-")
      cpus.update(range(int(a), int(b or a) + 1))
    out[p.name] = cpus
  return out

def rate(fn, nbytes, reps=5):
  fn(); torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(reps): fn()
  torch.cuda.synchronize()
  return nbytes * reps / (time.perf_counter() - t0) / 1e9

dst = torch.empty((n, (d + 63) // 64 * 64), device=dev)
streams = [torch.cuda.Stream() for _ in range(4)]
def measure(tag):
  rows = [torch.randn(d).pin_memory() for _ in range(n)]
  big = torch.empty(n * d).pin_memory()
  res = {}
  def copy_rows(k):
    def fn():
      if k == 1:
        for i, r in enumerate(rows): dst[i, :d].copy_(r, non_blocking=True)
      else:
        for s in streams[:k]: s.wait_stream(torch.cuda.current_stream())
        for i, r in enumerate(rows):
          with torch.cuda.stream(streams[i % k]): dst[i, :d].copy_(r, non_blocking=True)
        for s in streams[:k]: torch.cuda.current_stream().wait_stream(s)
    return fn
  for k in (1, 2, 4):
    res[f"rows_{k}streams"] = round(rate(copy_rows(k), n * d * 4), 1)
  flat = dst.view(-1)[:n * d]
  res["contiguous"] = round(rate(lambda: flat.copy_(big, non_blocking=True), n * d * 4), 1)
  back = torch.empty(d).pin_memory()
  res["d2h_5MB"] = round(rate(lambda: back.copy_(dst[0, :d], non_blocking=True), d * 4, 20), 1)
  print(tag, json.dumps(res), flush=True)
  return res

saved = os.sched_getaffinity(0)
print("allowed cpus:", len(saved), "nodes:", {k: len(v & saved) for k, v in nodes().items()}, flush=True)
out = {"default": measure("default")}
with hostmem.gpu_local_cpus(0) as ok:
  print("nvml ideal affinity applied:", ok, "cpus now:", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)), flush=True)
  out["nvml"] = measure("nvml-affinity")
for name, cpus in nodes().items():
  use = cpus & saved
  if not use: continue
  os.sched_setaffinity(0, use)
  out[name] = measure(name)
  os.sched_setaffinity(0, saved)
if len(sys.argv) > 1:
  pathlib.Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
