#!/usr/bin/env python3
# coding: utf-8
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = OrderedDict()
for r in rows[hdr + 1:]:
  if len(r) <= mv: continue
  name = r[kn].split("(")[0].replace("void ", "")[:60]
  try: v = float(r[mv].replace(",", ""))
  except ValueError: continue
  a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
print("%-62s %6s %12s %12s %7s" % ("kernel", "calls", "total us", "avg us", "share"))
for k, v in agg.items():
  print("%-62s %6d %12.1f %12.1f %6.1f%%" % (k, v[0], v[1] / 1e3, v[1] / v[0] / 1e3, 100 * v[1] / tot))
