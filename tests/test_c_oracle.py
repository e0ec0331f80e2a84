# coding: utf-8
"""The C restatement (`oracle/c`) against the NumPy oracle and, through it, against the
reference's golden outputs.  CPU-only."""

import numpy as np
import pytest

from conftest import load_goldens
import parity
from oracle import byzoracle as orc
from oracle import corc

pytestmark = pytest.mark.skipif(not corc.available(), reason="oracle/c/libbyzoracle.so not built (run __graft_entry__.build())")

@pytest.mark.parametrize("g", load_goldens(), ids=lambda g: g.name)
def test_c_oracle_matches_numpy_oracle_on_goldens(g):
  rows = [g.rows[i] for i in range(g.n)]
  parity.assert_bit_exact(corc.average(rows), orc.average(rows), "average")
  parity.assert_bit_exact(corc.median(rows), orc.median(rows), "median")
  for f in sorted({1, (g.n - 1) // 4, (g.n - 1) // 2} - {0}):
    if g.n < 2 * f + 1:
      continue
    parity.assert_bit_exact(corc.trmean(rows, f), orc.trmean(rows, f), f"trmean f={f}")
    parity.assert_bit_exact(corc.phocas(rows, f), orc.phocas(rows, f), f"phocas f={f}")
    parity.assert_bit_exact(corc.meamed(rows, f), orc.meamed(rows, f), f"meamed f={f}")
  if g.n >= 2:
    D = orc.pairwise_distances(rows)
    assert np.array_equal(corc.pairwise_distances(rows), D)
    order = list(range(g.n))[::-1]
    parity.assert_bit_exact(corc.average_selected(rows, order[:max(1, g.n // 2)]), orc._avg_rows(g.rows, order[:max(1, g.n // 2)]), "avg selected")

def test_c_oracle_large_random():
  rng = np.random.default_rng(5)
  x = rng.standard_normal((25, 50_000)).astype(np.float32)
  rows = [x[i] for i in range(25)]
  parity.assert_bit_exact(corc.trmean(rows, 10), orc.trmean(rows, 10), "trmean")
  parity.assert_bit_exact(corc.median(rows), orc.median(rows), "median")
  np.testing.assert_allclose(np.sqrt(corc.rowdist_sq(rows)).astype(np.float32), orc.row_norms(rows).astype(np.float32), rtol=0, atol=0)

@pytest.mark.parametrize("seed", range(40))
def test_c_oracle_randomised_against_numpy_oracle(seed):
  """ The C restatement is the checker at BASELINE's full sizes (tests/test_cuda_fullsize.py), so
  it is itself cross-checked over random shapes: every n class (1..64), ragged d, ties, NaN / inf
  entries, aliased rows, every f the rules accept. """
  rng = np.random.default_rng(1000 + seed)
  n = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 11, 16, 17, 25, 31, 32, 33, 51, 64]))
  d = int(rng.choice([1, 2, 15, 16, 31, 32, 33, 100, 257, 1000]))
  kind = seed % 4
  x = rng.standard_normal((n, d)).astype(np.float32)
  if kind == 1:                                             # heavy ties
    x = np.round(x * 2).astype(np.float32) / 2
  if kind == 2 and n > 1:                                   # non-finite entries
    x[rng.integers(n), rng.integers(d)] = np.nan
    x[rng.integers(n), rng.integers(d)] = np.inf
    x[rng.integers(n), rng.integers(d)] = -np.inf
  rows = [x[i] for i in range(n)]
  if kind == 3 and n > 2:                                   # aliased Byzantine rows
    rows = rows[:n - 2] + [rows[-1], rows[-1]]
  parity.assert_bit_exact(corc.average(rows), orc.average(rows), "average")
  parity.assert_bit_exact(corc.median(rows), orc.median(rows), "median")
  for f in range(1, (n - 1) // 2 + 1):
    parity.assert_bit_exact(corc.trmean(rows, f), orc.trmean(rows, f), f"trmean n={n} f={f}")
    if kind != 1:                                           # exact key ties: the order of the closest set is unspecified
      parity.assert_bit_exact(corc.phocas(rows, f), orc.phocas(rows, f), f"phocas n={n} f={f}")
      parity.assert_bit_exact(corc.meamed(rows, f), orc.meamed(rows, f), f"meamed n={n} f={f}")
  if n >= 2:
    got, want = corc.pairwise_distances(rows), orc.pairwise_distances(rows)
    assert np.array_equal(got, want, equal_nan=True)
    sel = [int(i) for i in rng.permutation(n)[:max(1, n // 2)]]
    parity.assert_bit_exact(corc.average_selected(rows, sel), orc._avg_rows(np.stack(rows), sel), "avg selected")
