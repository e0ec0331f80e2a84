# coding: utf-8
"""Comparator-network construction and analysis (build-time tool, pure Python).

Used by `tools/gen_networks.py` to emit `byzantinemomentum_b200/csrc/networks_gen.cuh`.
A network is a list of comparators (a, b), a < b: after the comparator, wire a holds the
min and wire b the max.  A sorting network leaves wire k holding the rank-k value.
"""

import itertools

def merge_exchange(n):
  """ Batcher's merge-exchange sort (Knuth TAOCP vol. 3, Algorithm 5.2.2M), any n >= 1. """
  net = []
  if n < 2:
    return net
  t = (n - 1).bit_length()
  p = 1 << (t - 1)
  while p > 0:
    q = 1 << (t - 1)
    r = 0
    d = p
    while True:
      for i in range(n - d):
        if (i & p) == r:
          net.append((i, i + d))
      if q == p:
        break
      d = q - p
      q >>= 1
      r = p
    p >>= 1
  return net

def bitonic_free_oddeven(n):
  """ Odd-even transposition sort: n rounds, n(n-1)/2 comparators (baseline for tiny n). """
  net = []
  for rnd in range(n):
    for i in range(rnd & 1, n - 1, 2):
      net.append((i, i + 1))
  return net

def prune(net, n, outputs):
  """ Backward liveness: keep only what the wires in `outputs` (after the network) depend on.
  Returns a list of (a, b, need_min, need_max, index in net). """
  live = set(outputs)
  kept = []
  for k in range(len(net) - 1, -1, -1):
    a, b = net[k]
    need_min = a in live
    need_max = b in live
    if need_min or need_max:
      kept.append((a, b, need_min, need_max, k))
      live.add(a)
      live.add(b)
  kept.reverse()
  return kept

def count_ops(pruned):
  return sum(int(x[2]) + int(x[3]) for x in pruned)

def depth(net, n):
  lvl = [0] * n
  for c in net:
    a, b = c[0], c[1]
    t = max(lvl[a], lvl[b]) + 1
    lvl[a] = lvl[b] = t
  return max(lvl) if lvl else 0

def apply(net, vals):
  v = list(vals)
  for c in net:
    a, b = c[0], c[1]
    if v[a] > v[b]:
      v[a], v[b] = v[b], v[a]
  return v

def check_sorts_01(net, n, outputs=None):
  """ 0-1 principle, bit-parallel over all 2**n inputs (n <= 26 or so). Checks that the
  wires in `outputs` (default: all) carry the right rank for every 0-1 input. """
  import numpy as np
  assert n <= 26
  N = 1 << n
  idx = np.arange(N, dtype=np.uint32)
  wires = [((idx >> k) & 1).astype(np.bool_) for k in range(n)]
  for c in net:
    a, b = c[0], c[1]
    lo = wires[a] & wires[b]
    hi = wires[a] | wires[b]
    wires[a], wires[b] = lo, hi
  ones = np.zeros(N, dtype=np.uint8)
  for k in range(n):
    ones += ((idx >> k) & 1).astype(np.uint8)
  outs = range(n) if outputs is None else outputs
  for k in outs:
    # sorted ascending: wire k is 1 iff number of ones > n-1-k
    expect = ones > (n - 1 - k)
    if not np.array_equal(wires[k], expect):
      return False
  return True

def remove_redundant_01(net, n):
  """ Drop comparators that never exchange on any 0-1 input (exhaustive, n <= 26). """
  import numpy as np
  N = 1 << n
  idx = np.arange(N, dtype=np.uint32)
  wires = [((idx >> k) & 1).astype(np.bool_) for k in range(n)]
  kept = []
  for (a, b) in net:
    swap = wires[a] & ~wires[b]
    if swap.any():
      lo = wires[a] & wires[b]
      hi = wires[a] | wires[b]
      wires[a], wires[b] = lo, hi
      kept.append((a, b))
  return kept
