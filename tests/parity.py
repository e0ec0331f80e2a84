# coding: utf-8
"""Parity rules shared by the oracle-vs-golden (CPU) and CUDA-vs-oracle/golden (GPU) tests.

Bars (BASELINE.json north_star):
  * selection indices / orders: bit-exact (aliased Byzantine rows are one object in the
    reference, so indices >= n_honest are equivalent);
  * values that are a fixed-order fp32 reduction of selected rows (average, krum, brute,
    aksel, cge, bulyan stage 1) and values that are an input element (median): bit-exact
    (value equality: -0.0 == +0.0, NaN == NaN);
  * trmean: bit-exact, except the trailing partial 32-column block (`d mod 32` columns)
    where ATen's vectorised `mean(dim=0)` switches to another summation order on the
    AVX-512 box the goldens were generated on (<= 1e-6 relative there);
  * closest-m means (phocas, meamed, bulyan stage 2): the reference sums in `topk`'s
    unspecified order -> 1e-6 relative to the magnitude of the summed values, and any
    valid resolution of an exact key tie across the selection boundary is accepted.
"""

import numpy as np

RTOL = 1e-6   # the tolerance BASELINE.json's north_star states for float reductions

def equal_values(a, b):
  """ Element-wise value equality with NaN == NaN (and -0.0 == +0.0). """
  a = np.asarray(a)
  b = np.asarray(b)
  return (a == b) | (np.isnan(a) & np.isnan(b))

def assert_bit_exact(got, ref, what=""):
  got = np.asarray(got)
  ref = np.asarray(ref)
  assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
  ok = equal_values(got, ref)
  if not ok.all():
    bad = np.flatnonzero(~ok)
    j = int(bad[0])
    raise AssertionError(f"{what}: {bad.size}/{ok.size} values differ, first at {j}: got {got[j]!r} ref {ref[j]!r}")

def assert_close_scaled(got, ref, scale, what="", rtol=RTOL, exempt=None):
  """ |got - ref| <= rtol * max(|ref|, scale) wherever not exempt; NaN/inf must match. """
  got = np.asarray(got, dtype=np.float64)
  ref = np.asarray(ref, dtype=np.float64)
  scale = np.broadcast_to(np.asarray(scale, dtype=np.float64), ref.shape)
  fin = np.isfinite(ref) & np.isfinite(got)
  with np.errstate(all="ignore"):
    ok = np.where(fin, np.abs(got - ref) <= rtol * np.maximum(np.abs(ref), scale), equal_values(got, ref))
  if exempt is not None:
    ok = ok | exempt
  if not ok.all():
    bad = np.flatnonzero(~ok)
    j = int(bad[0])
    raise AssertionError(f"{what}: {bad.size}/{ok.size} values out of tolerance, first at {j}: got {got[j]!r} ref {ref[j]!r} scale {scale[j]!r}")

def column_scale(rows):
  """ Magnitude of the summed values: mean |x| over the finite entries of each column. """
  rows = np.asarray(rows, dtype=np.float64)
  fin = np.isfinite(rows)
  cnt = np.maximum(fin.sum(axis=0), 1)
  return np.where(fin, np.abs(rows), 0.).sum(axis=0) / cnt

def assert_trmean(got, ref, rows, what=""):
  """ rows: the [n, d] inputs (tolerance on the tail columns is relative to the magnitude
  of the summed values, as the mean may cancel). """
  d = ref.shape[0]
  body = d - d % 32
  assert_bit_exact(got[:body], ref[:body], what + " [body]")
  if body < d:
    assert_close_scaled(got[body:], ref[body:], column_scale(np.asarray(rows)[:, body:]), what + " [ATen tail columns]")

def closest_ambiguous(rows, m, center):
  """ Coordinates where the m-closest choice is not unique (exact key tie across the
  boundary) or involves non-finite keys: any valid resolution is accepted there. """
  rows = np.asarray(rows, dtype=np.float32)
  n = rows.shape[0]
  with np.errstate(all="ignore"):
    key = np.abs((rows - center[None, :]).astype(np.float32))
  nonfinite = ~np.isfinite(key)
  key = np.where(np.isnan(key), np.inf, key)
  sk = np.sort(key, axis=0)
  amb = nonfinite.any(axis=0)
  if m < n:
    amb |= sk[m - 1] == sk[m]
  return amb

def assert_in_hull(got, rows, mask, what=""):
  """ On exempt coordinates a mean of column entries must still lie within the column's
  finite range (or be non-finite when the column holds non-finite entries). """
  rows = np.asarray(rows, dtype=np.float64)
  got = np.asarray(got, dtype=np.float64)
  idx = np.flatnonzero(mask)
  for j in idx:
    col = rows[:, j]
    fin = col[np.isfinite(col)]
    if not np.isfinite(got[j]):
      assert fin.size < col.size, f"{what}: non-finite result at {j} from a finite column"
    elif fin.size:
      lo, hi = fin.min(), fin.max()
      tol = RTOL * max(abs(lo), abs(hi), 1e-30)
      assert lo - tol <= got[j] <= hi + tol, f"{what}: result {got[j]} at {j} outside column range [{lo}, {hi}]"
