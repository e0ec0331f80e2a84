# coding: utf-8
"""Build-time tooling: the generated sorting networks are current, sort, and their ALU/FMA
masks only ever pick comparators whose two outputs are live.  CPU-only."""

import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
import netlib            # noqa: E402
import gen_networks      # noqa: E402

def test_generated_header_is_current():
  proc = subprocess.run([sys.executable, str(ROOT / "tools" / "gen_networks.py"), "--check"], capture_output=True, text=True)
  assert proc.returncode == 0, proc.stdout + proc.stderr

def test_networks_sort_exhaustively_up_to_16():
  for n in range(1, 17):
    assert netlib.check_sorts_01(netlib.merge_exchange(n), n)

def test_pruned_networks_still_select_the_requested_ranks():
  import random
  rng = random.Random(3)
  for n, outs in ((25, [12]), (25, list(range(10, 15))), (51, [25]), (51, list(range(12, 39))), (11, [5])):
    net = netlib.merge_exchange(n)
    kept = [(a, b) for (a, b, lo, hi, k) in netlib.prune(net, n, outs)]
    for _ in range(200):
      vals = [rng.random() for _ in range(n)]
      got = netlib.apply(kept, vals)
      want = sorted(vals)
      assert all(got[r] == want[r] for r in outs)

def test_mix_masks_pick_only_full_comparators_and_balance_the_pipes():
  for n, outs in ((25, list(range(25))), (25, [12]), (51, list(range(12, 39))), (64, list(range(64)))):
    net = netlib.merge_exchange(n)
    words, alu, fma = gen_networks.mix_mask(net, n, outs, n)
    live = {k: (lo, hi) for (a, b, lo, hi, k) in netlib.prune(net, n, outs)}
    for k in range(len(net)):
      if (words[k >> 6] >> (k & 63)) & 1:
        assert live.get(k) == (True, True)
    assert abs(alu - fma) <= 3
