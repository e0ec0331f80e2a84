#!/usr/bin/env python3
# coding: utf-8
""" Host rows in, host vector out (the e2e step of bench.py): ms per call of every host path, and of the
pipelined one (`bz_coordinate_host`) for several chunk counts.   python tools/e2e_pipeline_ab.py [n d f] """
import pathlib
import sys
import time

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from byzantinemomentum_b200 import engine, hostmem

def main():
  n, d, f = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (25, 1310922, 10)
  with hostmem.gpu_local_cpus(0) as local:
    rows = [torch.randn(d).pin_memory() for _ in range(n)]
  print("pinned rows allocated on the GPU-local NUMA node:", local)
  def timed(label):
    for _ in range(3):
      engine.trmean(rows, f)
    times = []
    for _ in range(20):
      t0 = time.perf_counter()
      engine.trmean(rows, f)
      times.append(time.perf_counter() - t0)
    times.sort()
    print(f"  {label:22s} median {times[10] * 1e3:7.3f} ms   min {times[0] * 1e3:7.3f} ms")
  print(f"trmean n={n} f={f} d={d}: {n * d * 4 / 1e6:.1f} MB in, {d * 4 / 1e6:.1f} MB out per call")
  for mode in ("lane", "lanes"):
    engine.forced_host_path = mode
    engine._host_paths.clear()
    timed(mode)
  engine.forced_host_path = "pipeline"
  for chunks in (1, 2, 4, 8, 16, 32):
    engine._PIPELINE_CHUNKS = chunks
    engine._host_paths.clear()
    timed(f"pipeline chunks={chunks}")

if __name__ == "__main__":
  main()
