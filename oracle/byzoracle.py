# coding: utf-8
"""CPU oracle for the Byzantine-robust aggregation rules (GARs).

TEST INFRASTRUCTURE ONLY.  This module is the *checker* for the CUDA path: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import it.  Nothing under `byzantinemomentum_b200/` or `native/` imports it,
and the product path has no CPU fallback.

It restates, in NumPy with explicit fp32 rounding, the algorithm of the reference's
`aggregators/*.py` (citations are relative to the reference root):

  average   aggregators/average.py:21-29
  median    aggregators/median.py:31-39
  trmean    aggregators/trmean.py:24-33,69-79     (ATen `mean(dim=0)` cascade order)
  phocas    aggregators/trmean.py:35-50,81-94
  meamed    aggregators/trmean.py:35-50,96-109
  krum      aggregators/krum.py:31-80             (pair layout: tools/misc.py:519-529)
  bulyan    aggregators/bulyan.py:31-84           (scores are never updated: :74-76 is dead)
  brute     aggregators/brute.py:32-80
  aksel     aggregators/aksel.py:24-64
  cge       aggregators/cge.py:28-57
  influence average.py:42-49, krum.py:126-150, brute.py:118-140, aksel.py:83-105, cge.py:72-93

Parity pin: the reference has no golden vectors or tests of its own for this path
(it has no test suite at all).  The oracle is therefore pinned against OUTPUTS OF THE
REFERENCE ITSELF, generated in the build container by `tests/golden/make_golden.py`
(which imports the unmodified reference from /root/reference) and committed as
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function here against
them.

Arithmetic conventions (all verified against the reference, see tests):
  * data is fp32; every add/sub/div that feeds an output is a single IEEE fp32 operation
    (NumPy float32 scalars/arrays give exactly that), division is a true division by the
    fp32 count, never a multiplication by a reciprocal;
  * Python `sum()` starts from integer 0, so the first operation is `0 + g0`
    (turns -0.0 into +0.0); `cge` starts from a clone instead;
  * pairwise distances are `fl32(||fl32(x - y)||_2)`.  The reference gets them from ATen's
    fp32 `norm()`, whose summation order is build specific (8 serial lanes on the survey
    box, relative bias ~ -1.8e-5 at d = 1.3M); the oracle uses the correctly rounded value
    (sum of squares in fp64, sqrt, one rounding to fp32).  Selections therefore agree with
    the reference whenever the decision margins exceed that noise; `margins` in the
    returned info lets tests tell "noise" from "bug";
  * scores are Python-double sums, sorts are stable.
"""

import itertools
import math

import numpy as np

__all__ = [
  "average", "median", "trmean", "phocas", "meamed", "krum", "bulyan", "brute", "aksel", "cge",
  "pairwise_distances", "krum_order", "bulyan_order", "brute_selection", "aksel_order", "cge_order",
  "influence", "closest_mean", "aten_mean_dim0", "as_matrix", "GARS", "compute_avg_dev_max"]

F32 = np.float32

# ---------------------------------------------------------------------------- #
# Helpers

def as_matrix(gradients):
  """ Stack a list of n fp32 vectors of length d into an [n, d] fp32 array (copy).
  Mirrors the `torch.stack(gradients)` of median.py:39, trmean.py:79. """
  g = np.stack([np.asarray(x, dtype=F32).reshape(-1) for x in gradients])
  if g.ndim != 2:
    raise ValueError("expected a list of 1-D vectors")
  return g

def _seq_sum_rows(g, order, zero_init=True):
  """ fp32 left-to-right sum of rows `order` of g (Python `sum()` semantics when zero_init). """
  acc = (F32(0) + g[order[0]]) if zero_init else g[order[0]].copy()
  for k in order[1:]:
    acc = acc + g[k]
  return acc.astype(F32, copy=False)

def _avg_rows(g, order, zero_init=True):
  """ `sum(rows in order).div_(len(order))`: krum.py:80, brute.py:80, aksel.py:64, cge.py:53-56. """
  with np.errstate(all="ignore"):
    return (_seq_sum_rows(g, order, zero_init) / F32(len(order))).astype(F32)

def aten_mean_dim0(s):
  """ fp32 `Tensor.mean(dim=0)` of an [R, d] array, in ATen's CPU summation order:
  cascade sum with 16-row blocks (level k holds up to 16**k rows), partial levels combined
  as ((acc0 + acc1) + acc2) + acc3, then one division by R.  Purely sequential for R <= 16.
  Follows the order the reference gets from `values[f:-f].mean(dim=0)` (trmean.py:33).
  (ATen's trailing partial 32-column block, `d mod 32` columns, uses another order on the
  AVX-512 build the goldens come from, differing by <= ~6e-8 abs: not modelled.) """
  s = np.asarray(s, dtype=F32)
  R = s.shape[0]
  d = s.shape[1]
  acc = [np.zeros(d, dtype=F32) for _ in range(4)]
  with np.errstate(all="ignore"):
    for r in range(R):
      acc[0] = acc[0] + s[r]
      c = r + 1
      # Fold full blocks upward: level j is folded into level j+1 every 16**(j+1) rows
      for lvl in range(3):
        if c % (16 ** (lvl + 1)) == 0:
          acc[lvl + 1] = acc[lvl + 1] + acc[lvl]
          acc[lvl] = np.zeros(d, dtype=F32)
        else:
          break
    total = ((acc[0] + acc[1]) + acc[2]) + acc[3]
    return (total / F32(R)).astype(F32)

def _sort_nan_last(g):
  """ `Tensor.sort(dim=0)`: ascending, NaN last (NumPy does the same). """
  return np.sort(g, axis=0, kind="stable")

# ---------------------------------------------------------------------------- #
# Coordinate-wise rules

def average(gradients, **kwargs):
  """ average.py:29 — `sum(gradients) / len(gradients)`. """
  g = as_matrix(gradients)
  return _avg_rows(g, list(range(g.shape[0])))

def median(gradients, **kwargs):
  """ median.py:39 — lower median per coordinate (rank (n-1)//2); a NaN anywhere in the
  column gives NaN (torch >= 1.7 `median(dim)` propagates NaN). """
  g = as_matrix(gradients)
  n = g.shape[0]
  s = _sort_nan_last(g)
  out = s[(n - 1) // 2].copy()
  out[np.isnan(g).any(axis=0)] = np.nan
  return out

def trmean(gradients, f, **kwargs):
  """ trmean.py:33 — `g.sort(dim=0).values[f:-f].mean(dim=0)`. """
  g = as_matrix(gradients)
  n = g.shape[0]
  s = _sort_nan_last(g)
  return aten_mean_dim0(s[f:n - f])

def closest_mean(g, m, c, return_info=False):
  """ trmean.py:35-50 / bulyan.py:78-84 — per coordinate, mean of the m entries of g with
  the smallest fl32|fl32(x - c)| (NaN keys count as largest: `topk(largest=False)`).
  The reference sums in topk's (unspecified) output order; the oracle sums the chosen
  entries in ascending row order, sequentially, then divides by m.
  With return_info, also returns `ambiguous`: a boolean [d] mask of coordinates where the
  choice is not unique (a key tie across the boundary, or a NaN key inside the chosen set),
  i.e. where any valid tie resolution must be accepted. """
  g = np.asarray(g, dtype=F32)
  n, d = g.shape
  with np.errstate(all="ignore"):
    key = np.abs((g - c[None, :]).astype(F32))
  # Stable: ties resolved towards the lower row index; NaN keys after +inf keys
  rank_key = np.where(np.isnan(key), np.inf, key).astype(np.float64)
  order = np.lexsort((np.arange(n)[:, None].repeat(d, 1), np.isnan(key), rank_key), axis=0)
  chosen = np.sort(order[:m], axis=0)
  vals = np.take_along_axis(g, chosen, axis=0)
  with np.errstate(all="ignore"):
    acc = vals[0].copy()
    for k in range(1, m):
      acc = acc + vals[k]
    out = (acc / F32(m)).astype(F32)
  if not return_info:
    return out
  sk = np.take_along_axis(rank_key, order, axis=0)
  nk = np.take_along_axis(np.isnan(key), order, axis=0)
  ambiguous = np.zeros(d, dtype=bool)
  if m < n:
    ambiguous |= (sk[m - 1] == sk[m]) & (nk[m - 1] == nk[m])
  ambiguous |= nk[:m].any(axis=0)
  return out, ambiguous

def phocas(gradients, f, **kwargs):
  """ trmean.py:81-94 — closest(g, f, trmean(g, f)), m = n - f. """
  g = as_matrix(gradients)
  c = trmean(g, f)
  return closest_mean(g, g.shape[0] - f, c)

def meamed(gradients, f, **kwargs):
  """ trmean.py:96-109 — closest(g, f, median(g)), m = n - f. """
  g = as_matrix(gradients)
  c = median(g)
  return closest_mean(g, g.shape[0] - f, c)

# ---------------------------------------------------------------------------- #
# Distances

def pairwise_distances(gradients, map_nonfinite=True):
  """ n x n symmetric matrix (float64 holding fp32 values) of `x.sub(y).norm().item()`
  (krum.py:45, bulyan.py:50, brute.py:45); diagonal = 0.  Non-finite -> +inf when
  map_nonfinite (krum.py:46-47, bulyan.py:51-52; brute.py keeps the raw value). """
  g = as_matrix(gradients)
  n = g.shape[0]
  D = np.zeros((n, n), dtype=np.float64)
  with np.errstate(all="ignore"):
    for x in range(n - 1):
      diff = (g[x + 1:] - g[x][None, :]).astype(F32).astype(np.float64)
      sq = np.einsum("ij,ij->i", diff, diff)
      dist = np.sqrt(sq).astype(F32).astype(np.float64)
      if map_nonfinite:
        dist = np.where(np.isfinite(dist), dist, np.inf)
      D[x, x + 1:] = dist
      D[x + 1:, x] = dist
  return D

def row_norms(gradients):
  """ cge.py:36-37 — `grad.norm().item()`, non-finite -> +inf. """
  g = as_matrix(gradients).astype(np.float64)
  with np.errstate(all="ignore"):
    nrm = np.sqrt(np.einsum("ij,ij->i", g, g)).astype(F32).astype(np.float64)
  return np.where(np.isfinite(nrm), nrm, np.inf)

def _stable_argsort(values):
  """ Python `list.sort(key=...)`: stable, ascending; NaN keys are not supported here. """
  return sorted(range(len(values)), key=lambda i: values[i])

def _margin(sorted_scores, m):
  """ Relative gap between the last selected and the first rejected score (inf when none). """
  if m >= len(sorted_scores):
    return math.inf
  a, b = sorted_scores[m - 1], sorted_scores[m]
  if not (math.isfinite(a) and math.isfinite(b)):
    return math.inf if a != b else 0.
  return (b - a) / max(abs(b), 1e-300)

def _min_adjacent_gap(sorted_scores, upto):
  """ Smallest relative gap between consecutive sorted scores among the first `upto`+1. """
  gap = math.inf
  for k in range(min(upto, len(sorted_scores) - 1)):
    a, b = sorted_scores[k], sorted_scores[k + 1]
    if math.isfinite(a) and math.isfinite(b):
      gap = min(gap, (b - a) / max(abs(b), 1e-300))
    elif a == b:
      gap = 0.
  return gap

# ---------------------------------------------------------------------------- #
# Multi-Krum

def krum_order(D, f):
  """ krum.py:52-62 — score_i = double sum of the n-f-1 smallest distances from i to the
  others (ascending, left to right); returns (order, scores) with order = stable argsort. """
  n = D.shape[0]
  scores = []
  for i in range(n):
    dists = sorted(float(D[i, j]) for j in range(n) if j != i)
    scores.append(sum(dists[:n - f - 1]))
  return _stable_argsort(scores), scores

def krum(gradients, f, m=None, return_info=False, **kwargs):
  """ krum.py:65-80 — average of the m best-scored gradients, summed in score order. """
  g = as_matrix(gradients)
  n = g.shape[0]
  if m is None:
    m = n - f - 2
  D = pairwise_distances(g)
  order, scores = krum_order(D, f)
  out = _avg_rows(g, order[:m])
  if return_info:
    ss = [scores[i] for i in order]
    return out, dict(selection=order[:m], order=order, scores=scores, distances=D,
                     margin=min(_margin(ss, m), _min_adjacent_gap(ss, m)))
  return out

# ---------------------------------------------------------------------------- #
# Bulyan

def bulyan_order(D, f, m):
  """ bulyan.py:56-62 — row i of the distance table has +inf on its diagonal; score_i =
  double sum of the m smallest entries of the row (ascending, the +inf included if m = n);
  stable argsort.  The scores are never updated afterwards (bulyan.py:74-76 is dead code). """
  n = D.shape[0]
  scores = []
  for i in range(n):
    row = [float(D[i, j]) if j != i else math.inf for j in range(n)]
    row.sort()
    scores.append(sum(row[:m]))
  return _stable_argsort(scores), scores

def bulyan_stage1(g, order, f, m, scores=None):
  """ bulyan.py:64-73 — theta = n-2f-2 rows; row i = mean of the gradients at sorted
  positions i .. i+m_i-1 (m_i = min(m, m_max - i)), summed left to right.
  Degenerate input: a pruned entry is `(inf, None)` and the sort is stable, so from the
  second iteration on, pruned entries precede every row whose score is +inf; if fewer than
  m_i finite-score rows remain, the reference indexes `gradients[None]` and raises
  TypeError (bulyan.py:70).  Mirrored here when `scores` is given. """
  n = g.shape[0]
  m_max = n - f - 2
  theta = n - 2 * f - 2
  finite = n if scores is None else sum(1 for s in scores if math.isfinite(s))
  sel = np.empty((theta, g.shape[1]), dtype=F32)
  for i in range(theta):
    m = min(m, m_max - i)
    if i >= 1 and finite - i < m:
      raise TypeError("bulyan: too many non-finite scores (the reference fails on gradients[None])")
    sel[i] = _avg_rows(g, order[i:i + m])
  return sel

def bulyan(gradients, f, m=None, return_info=False, **kwargs):
  """ bulyan.py:31-84. """
  g = as_matrix(gradients)
  n = g.shape[0]
  m_max = n - f - 2
  if m is None:
    m = m_max
  D = pairwise_distances(g)
  order, scores = bulyan_order(D, f, m)
  sel = bulyan_stage1(g, order, f, m, scores)
  theta = sel.shape[0]
  beta = theta - 2 * f
  med = median(sel)
  if return_info:
    out, ambiguous = closest_mean(sel, beta, med, return_info=True)
    ss = [scores[i] for i in order]
    return out, dict(order=order, scores=scores, distances=D, stage1=sel, ambiguous=ambiguous,
                     margin=_min_adjacent_gap(ss, m_max))
  return closest_mean(sel, beta, med)

# ---------------------------------------------------------------------------- #
# Brute

def brute_selection(D_raw, f, return_info=False):
  """ brute.py:47-68 — subsets of size n-f in lexicographic order; a subset with a
  non-finite pair distance is skipped; strictly smaller diameter wins (first minimum). """
  n = D_raw.shape[0]
  best, best_diam, second = None, None, math.inf
  for cur in itertools.combinations(range(n), n - f):
    diam = 0.
    ok = True
    for a in range(len(cur) - 1):
      x = cur[a]
      for b in range(a + 1, len(cur)):
        dist = float(D_raw[x, cur[b]])
        if not math.isfinite(dist):
          ok = False
          break
        if dist > diam:
          diam = dist
      if not ok:
        break
    if not ok:
      continue
    if best is None or diam < best_diam:
      if best is not None:
        second = min(second, best_diam)
      best, best_diam = cur, diam
    else:
      second = min(second, diam)
  if best is None:
    raise AssertionError("Too many non-finite gradients")  # brute.py:67
  if return_info:
    margin = (second - best_diam) / max(second, 1e-300) if math.isfinite(second) else math.inf
    return list(best), dict(diameter=best_diam, margin=margin)
  return list(best)

def brute(gradients, f, return_info=False, **kwargs):
  """ brute.py:70-80. """
  g = as_matrix(gradients)
  D = pairwise_distances(g, map_nonfinite=False)
  if return_info:
    sel, info = brute_selection(D, f, return_info=True)
    info.update(selection=sel, distances=D)
    return _avg_rows(g, sel), info
  sel = brute_selection(D, f)
  return _avg_rows(g, sel)

# ---------------------------------------------------------------------------- #
# Aksel

def aksel_order(g):
  """ aksel.py:39-49 — squared distance of every row to the coordinate-wise median,
  `(x - m).pow_(2).sum().item()`; correctly rounded fp32 of the exact sum of the fp32
  squares (ATen's own order is thread-count dependent); stable sort. """
  med = median(g)
  with np.errstate(all="ignore"):
    diff = (g - med[None, :]).astype(F32)
    sq = (diff * diff).astype(F32).astype(np.float64)
    dist = sq.sum(axis=1).astype(F32).astype(np.float64)
  dl = [float(x) for x in dist]
  if any(math.isnan(x) for x in dl):
    # sort order with NaN keys is undefined in the reference: keep the input order
    return list(range(g.shape[0])), dl
  return _stable_argsort(dl), dl

def aksel(gradients, f, mode="mid", return_info=False, **kwargs):
  """ aksel.py:52-64. """
  g = as_matrix(gradients)
  n = g.shape[0]
  if mode == "mid":
    c = (n + 1) // 2
  elif mode == "n-f":
    c = n - f
  else:
    raise NotImplementedError
  order, dist = aksel_order(g)
  out = _avg_rows(g, order[:c])
  if return_info:
    ss = [dist[i] for i in order]
    return out, dict(selection=order[:c], order=order, distances=dist,
                     margin=min(_margin(ss, c), _min_adjacent_gap(ss, c)))
  return out

# ---------------------------------------------------------------------------- #
# CGE

def cge_order(g):
  """ cge.py:28-38 — stable sort by fp32 norm (non-finite -> inf). """
  nrm = [float(x) for x in row_norms(g)]
  return _stable_argsort(nrm), nrm

def cge(gradients, f, return_info=False, **kwargs):
  """ cge.py:40-57 — clone of the smallest-norm gradient, add_ the next m-1, div_(m). """
  g = as_matrix(gradients)
  n = g.shape[0]
  m = n - f
  order, nrm = cge_order(g)
  # `res = clone(first); for g in normed[1:m]: res.add_(g); res.div_(m)` — cge.py:53-56;
  # `check` does not validate f (cge.py:59-70), so m <= 0 divides the first row by m
  picked = [order[0]] + order[1:m]
  with np.errstate(all="ignore"):
    out = (_seq_sum_rows(g, picked, zero_init=False) / F32(m)).astype(F32)
  if return_info:
    ss = [nrm[i] for i in order]
    return out, dict(selection=picked, order=order, norms=nrm,
                     margin=min(_margin(ss, m), _min_adjacent_gap(ss, m)))
  return out

# ---------------------------------------------------------------------------- #
# Study metrics

def compute_avg_dev_max(samples):
  """ tools/pytorch.py:97-125 — (average, norm of the average, norm standard deviation, max
  absolute coordinate of the average).  The average follows the reference's fp32 op order
  (clone, add_ in list order, div_); the three scalars are reductions whose fp32 summation order
  the reference leaves to ATen (`norm`, `dot`), so they are restated in fp64 and compared within
  the float tolerance. """
  if len(samples) == 0:
    return None, math.nan, math.nan, math.nan                              # :105-106
  g = as_matrix(samples)
  n = g.shape[0]
  with np.errstate(all="ignore"):
    avg = (_seq_sum_rows(g, list(range(n)), zero_init=False) / F32(n)).astype(F32)   # :108-111
    a64 = avg.astype(np.float64)
    norm_avg = math.sqrt(float(np.sum(a64 * a64)))                         # :112
    absa = np.abs(avg)
    norm_max = float("nan") if np.isnan(absa).any() else float(absa.max()) if absa.size else 0.  # :113
    if n >= 2:
      norm_var = 0.
      for i in range(n):                                                   # :116-119
        diff = (g[i] - avg).astype(F32).astype(np.float64)
        norm_var += float(np.sum(diff * diff))
      norm_dev = math.sqrt(norm_var / (n - 1))                             # :120-121
    else:
      norm_dev = math.nan                                                  # :122-123
  return avg, norm_avg, norm_dev, norm_max

# ---------------------------------------------------------------------------- #
# Influence (ratio of accepted Byzantine gradients)

def influence(name, honests, attacks, f=None, m=None, mode="mid", **kwargs):
  """ average.py:42-49, krum.py:126-150, brute.py:118-140, aksel.py:83-105, cge.py:72-93.
  The reference tests object identity (`gradient is attack`); with honests first and
  attacks last this is `selected index >= len(honests)`. """
  nh, na = len(honests), len(attacks)
  n = nh + na
  if name == "average":
    return na / n
  g = as_matrix(list(honests) + list(attacks))
  if name == "krum":
    if m is None:
      m = n - f - 2
    order, _ = krum_order(pairwise_distances(g), f)
    sel = order[:m]
  elif name == "brute":
    sel = brute_selection(pairwise_distances(g, map_nonfinite=False), f)
  elif name == "aksel":
    c = (n + 1) // 2 if mode == "mid" else n - f
    sel = aksel_order(g)[0][:c]
  elif name == "cge":
    sel = cge_order(g)[0][:n - f]
  else:
    raise KeyError(f"no influence for {name!r}")
  return sum(1 for i in sel if i >= nh) / len(sel)

GARS = dict(average=average, median=median, trmean=trmean, phocas=phocas, meamed=meamed,
            krum=krum, bulyan=bulyan, brute=brute, aksel=aksel, cge=cge)
