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
  # the same 16 evaluations through the C ABI with every argument prepared (no Python bookkeeping in
  # the timed loop): what the library itself gains
  if gar == "krum":
    import ctypes
    from byzantinemomentum_b200 import _lib
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    ws = engine._workspace(dev, st)
    outb = torch.empty(d, device=dev); order = torch.empty(n, dtype=torch.int32, device=dev)
    tabs = [torch.empty(64 * 64, dtype=torch.float64, device=dev) for _ in range(2)]
    lists = [honest + [attacks[k]] * nb for k in range(17)]
    ptrs = [(ctypes.c_void_p * n)(*[g.data_ptr() for g in l]) for l in lists]
    old = (ctypes.c_int32 * n)(*([i for i in range(n - nb)] + [-1] * nb))
    mode = ctypes.c_int(0)
    m = n - f - 2
    def full(k):
      return lib.bz_krum(ptrs[k], n, f, m, d, outb.data_ptr(), order.data_ptr(), ws.data_ptr(), ws.numel(), st)
    def reuse(k, cur):
      return lib.bz_krum_reuse(ptrs[k], n, f, m, d, outb.data_ptr(), order.data_ptr(), old, tabs[cur].data_ptr(), n - nb + 1,
                               tabs[cur ^ 1].data_ptr(), ctypes.byref(mode), ws.data_ptr(), ws.numel(), st)
    for name in ("abi_full_us", "abi_reuse_us"):
      best = None
      for rep in range(5):
        lib.bz_krum_reuse(ptrs[16], n, f, m, d, outb.data_ptr(), order.data_ptr(), None, None, 0, tabs[0].data_ptr(), ctypes.byref(mode),
                          ws.data_ptr(), ws.numel(), st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        cur = 0
        for k in range(16):
          if name == "abi_full_us": full(k)
          else:
            reuse(k, cur); cur ^= 1
        b.record(); torch.cuda.synchronize()
        t = a.elapsed_time(b) / 16 * 1e3
        best = t if best is None else min(best, t)
      rec[name] = best
    rec["abi_ratio"] = rec["abi_reuse_us"] / rec["abi_full_us"]
    rec["abi_last_mode"] = mode.value
  out.append(rec)
  print(json.dumps(rec), flush=True)
engine.config.reuse_distances = True
if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
  pathlib.Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
