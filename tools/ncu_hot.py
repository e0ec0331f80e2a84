#!/usr/bin/env python3
# coding: utf-8
"""Summarise the `--page source --csv` export of an ncu report: instructions executed and stall
samples per SASS opcode, per stall reason, and the hottest instructions.
    ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [top]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
def num(r, name):
  try: return float(r[col[name]])
  except Exception: return 0.
def opcode(src):
  s = re.sub(r"^\s*(@!?U?P\d+\s+)?", "", src.strip())
  return s.split()[0].split(".")[0] if s else "?"
tot_inst = sum(num(r, "Instructions Executed") for r in data)
tot_samp = sum(num(r, "# Samples") for r in data)
by = collections.defaultdict(lambda: [0., 0.])
for r in data:
  o = opcode(r[col["Source"]]); by[o][0] += num(r, "Instructions Executed"); by[o][1] += num(r, "# Samples")
print(f"total warp instructions {tot_inst:.0f}, samples {tot_samp:.0f}")
print("opcode          inst%  samples%")
for o, (i, s) in sorted(by.items(), key=lambda kv: -kv[1][1])[:top]:
  print(f"{o:14s} {100*i/tot_inst:6.2f} {100*s/tot_samp:8.2f}")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("\nstall reason  samples%")
for h in sorted(reasons, key=lambda h: -sum(num(r, h) for r in data)):
  v = sum(num(r, h) for r in data)
  if v: print(f"{h:24s} {100*v/tot_samp:6.2f}")
print("\nhottest instructions (samples%, exec, top reasons)")
for idx, r in sorted(enumerate(data), key=lambda t: -num(t[1], "# Samples"))[:top]:
  rs = sorted(((num(r, h), h) for h in reasons), reverse=True)[:3]
  print(f"{100*num(r, '# Samples')/tot_samp:5.2f}% #{idx:5d} x{num(r, 'Instructions Executed'):9.0f} {r[col['Source']].strip()[:60]:60s} " + " ".join(f"{h[6:]}={v:.0f}" for v, h in rs if v))
