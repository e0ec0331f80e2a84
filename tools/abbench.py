#!/usr/bin/env python3
# coding: utf-8
"""A/B timing of library variants on ONE box (development aid).
  python tools/abbench.py libA.so libB.so ...   -> per variant, back-to-back Plan() calls
Each variant runs in its own subprocess (BYZAGG_LIBRARY), twice, interleaved."""
import json, os, pathlib, subprocess, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent

CHILD = r"""
import json, sys, torch
sys.path.insert(0, %r)
import byzantinemomentum_b200 as bz
dev = torch.device("cuda", 0)
out = {}
cases = json.loads(sys.argv[1])
for gar, n, f, d in cases:
  sets = max(1, min(6, -(-3 * 126 * 2**20 // (n * d * 4))))
  gen = torch.Generator(device=dev).manual_seed(3)
  stacks = [[torch.randn(d, device=dev, generator=gen) for _ in range(n)] for _ in range(sets)]
  plans = [bz.Plan(gar, rows, f=f) for rows in stacks]
  for k in range(5): plans[k %% sets]()
  torch.cuda.synchronize()
  K = 200 if d < 5e6 else 30
  best = None
  for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(K): plans[k %% sets]()
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / K * 1e3
    best = t if best is None else min(best, t)
  out["%%s n=%%d f=%%d d=%%d" %% (gar, n, f, d)] = best
  del stacks, plans
  torch.cuda.empty_cache()
print("RESULT " + json.dumps(out))
""" % str(ROOT)

def main():
  libs = [a for a in sys.argv[1:] if not a.startswith("--")]
  if "--dist" in sys.argv:      # the distance-based rules (K2 experiments)
    cases = [("krum", 25, 5, 1310922), ("bulyan", 25, 5, 1310922), ("krum", 25, 5, 36489290), ("cge", 25, 5, 1310922), ("krum", 51, 12, 1310922)]
  elif "--wide" in sys.argv:    # clusters of 16-warp CTAs (n > 35)
    cases = [("krum", 51, 12, 4568373), ("bulyan", 51, 12, 4568373), ("krum", 51, 12, 1310922), ("krum", 40, 9, 1310922), ("krum", 45, 10, 1310922),
             ("krum", 64, 15, 1310922), ("krum", 51, 12, 36546980), ("krum", 35, 8, 1310922)]
  elif "--bulyan" in sys.argv:  # K4 experiments
    cases = [("bulyan", 25, 5, 1310922), ("bulyan", 11, 2, 1310922), ("bulyan", 51, 12, 4568373), ("bulyan", 25, 5, 36489290), ("bulyan", 15, 3, 1310922), ("krum", 25, 5, 1310922)]
  else:
   cases = [("average", 25, 10, 1310922), ("median", 25, 10, 1310922), ("trmean", 25, 10, 1310922), ("trmean", 25, 7, 1310922),
           ("median", 25, 10, 36489290), ("trmean", 25, 10, 36489290), ("trmean", 25, 7, 36489290),
           ("median", 51, 12, 4568373), ("trmean", 51, 12, 4568373), ("median", 11, 5, 1310922), ("phocas", 25, 10, 1310922)]
  results = {}
  for rep in range(2):
    for lib in libs:
      # a variant is `library.so` or `library.so@NAME=VALUE` (an environment switch of the same library)
      path, _, setting = lib.partition("@")
      env = dict(os.environ, BYZAGG_LIBRARY=str(pathlib.Path(path).resolve()))
      if setting:
        env[setting.split("=", 1)[0]] = setting.split("=", 1)[1]
      proc = subprocess.run([sys.executable, "-c", CHILD, json.dumps(cases)], env=env, capture_output=True, text=True)
      line = [l for l in proc.stdout.splitlines() if l.startswith("RESULT ")]
      if not line:
        print(lib, "FAILED", proc.stderr[-1500:])
        continue
      for k, v in json.loads(line[0][7:]).items():
        results.setdefault(k, {}).setdefault(lib, []).append(v)
  names = [pathlib.Path(l.partition("@")[0]).name[:12] + ("@" + l.partition("@")[2] if "@" in l else "") for l in libs]
  print("%-36s" % "case (us per call, best of 3; two runs)" + "".join("%30s" % n for n in names))
  for k, per in results.items():
    print("%-36s" % k + "".join("%30s" % " / ".join("%.1f" % x for x in per.get(l, [])) for l in libs))

if __name__ == "__main__":
  main()
