# coding: utf-8
"""Study metrics (`tools.compute_avg_dev_max`, tools/pytorch.py:97-125).

CPU: the oracle against what the reference itself returned (tests/golden/study_avg_dev_max.npz,
made by tests/golden/make_study_golden.py).  GPU: `bz_avg_dev_max` through the C ABI against the
same goldens and against the oracle at ragged / large sizes.

Bars: the average is a fixed-order fp32 computation -> bit-exact.  The three scalars are fp32
reductions whose order ATen chooses (norm, dot); they are compared at SCALAR_RTOL relative.
"""

import json
import math
import pathlib

import numpy as np
import pytest

from oracle import byzoracle as orc
from parity import assert_bit_exact

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden" / "study_avg_dev_max.npz"
SCALAR_RTOL = 2e-6     # fp32-accumulated reference vs fp64-accumulated restatement / device

def _cases():
  z = np.load(GOLDEN, allow_pickle=False)
  manifest = json.loads(str(z["manifest"]))
  return [pytest.param(c, id=c["tag"]) for c in manifest["cases"]]

def _load(case):
  z = np.load(GOLDEN, allow_pickle=False)
  rows = np.load(GOLDEN.parent / case["fixture"], allow_pickle=False)["rows"][case["lo"]:case["hi"]]
  avg = z[case["tag"] + "/avg"] if case["has_avg"] else None
  return rows, avg, float(case["norm_avg"]), float(case["norm_dev"]), float(case["norm_max"])

def _same_scalar(got, want, what):
  if math.isnan(want):
    assert math.isnan(got), f"{what}: {got} instead of nan"
  elif math.isinf(want):
    assert got == want, f"{what}: {got} instead of {want}"
  else:
    assert abs(got - want) <= SCALAR_RTOL * abs(want) + 1e-30, f"{what}: {got} vs {want}"

def _check(result, avg, norm_avg, norm_dev, norm_max):
  got_avg, got_norm, got_dev, got_max = result
  if avg is None:
    assert got_avg is None
  else:
    assert_bit_exact(np.asarray(got_avg), avg)
  _same_scalar(got_norm, norm_avg, "norm_avg")
  _same_scalar(got_dev, norm_dev, "norm_dev")
  _same_scalar(got_max, norm_max, "norm_max")

@pytest.mark.parametrize("case", _cases())
def test_oracle_matches_reference(case):
  rows, avg, norm_avg, norm_dev, norm_max = _load(case)
  _check(orc.compute_avg_dev_max(list(rows)), avg, norm_avg, norm_dev, norm_max)

def test_oracle_no_sample():
  avg, a, b, c = orc.compute_avg_dev_max([])
  assert avg is None and math.isnan(a) and math.isnan(b) and math.isnan(c)

# ---------------------------------------------------------------------------- #
# CUDA

def _to_cuda(rows):
  import torch
  return [torch.from_numpy(np.ascontiguousarray(r)).cuda() for r in rows]

@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases())
def test_cuda_matches_reference(case):
  import byzantinemomentum_b200 as bz
  rows, avg, norm_avg, norm_dev, norm_max = _load(case)
  got = bz.compute_avg_dev_max(_to_cuda(rows))
  if got[0] is not None:
    got = (got[0].cpu().numpy(),) + got[1:]
  _check(got, avg, norm_avg, norm_dev, norm_max)

@pytest.mark.gpu
@pytest.mark.parametrize("n,d", [(1, 1), (2, 3), (25, 4099), (25, 1310922), (51, 100003), (64, 8191)])
def test_cuda_matches_oracle(n, d):
  import torch
  import byzantinemomentum_b200 as bz
  rng = np.random.default_rng(1000 * n + d % 997)
  rows = (rng.standard_normal((n, d)) * rng.uniform(0.1, 10., (n, 1))).astype(np.float32)
  base = torch.from_numpy(rows).cuda()
  for offset in (0, 1):                      # 16-byte aligned rows and unaligned views
    if offset:
      flat = torch.empty(n * (d + 4) + 1, dtype=torch.float32, device="cuda")
      samples = [flat[1 + i * (d + 4):1 + i * (d + 4) + d] for i in range(n)]
      for s, b in zip(samples, base):
        s.copy_(b)
    else:
      samples = list(base)
    avg, norm_avg, norm_dev, norm_max = bz.compute_avg_dev_max(samples)
    want = orc.compute_avg_dev_max(list(rows))
    _check((avg.cpu().numpy(), norm_avg, norm_dev, norm_max), *want)

@pytest.mark.gpu
def test_cuda_non_finite_and_async():
  import torch
  import byzantinemomentum_b200 as bz
  rows = np.random.default_rng(5).standard_normal((7, 1000)).astype(np.float32)
  rows[2, 17] = np.nan
  rows[4, 400] = np.inf
  got = bz.compute_avg_dev_max(_to_cuda(rows))
  want = orc.compute_avg_dev_max(list(rows))
  _check((got[0].cpu().numpy(),) + got[1:], *want)
  avg, stats = bz.engine.avg_dev_max_async(_to_cuda(rows[:, :16]))
  assert stats.shape == (9,) and stats.dtype == torch.float64 and stats.is_cuda
  assert bz.compute_avg_dev_max([]) [0] is None

def test_install_tools_keeps_cpu_samples_on_the_reference_function():
  import types
  import torch
  import byzantinemomentum_b200 as bz
  calls = []
  def stock(samples):
    calls.append(len(samples))
    return "stock"
  tools = types.SimpleNamespace(compute_avg_dev_max=stock)
  assert bz.plugin.install_tools(tools) is stock
  assert tools.compute_avg_dev_max([torch.zeros(3)]) == "stock"
  assert tools.compute_avg_dev_max([]) == "stock"
  assert calls == [1, 0]

@pytest.mark.parametrize("case", [c for c in _cases() if c.values[0]["has_avg"]][:12])
def test_refcost_sequence_is_the_reference_sequence(case):
  """ oracle/refcost.py's timed operator sequence returns what the reference returned. """
  import torch
  from oracle import refcost
  rows, avg, norm_avg, norm_dev, norm_max = _load(case)
  torch.set_num_threads(1)
  mean, length, spread, largest = refcost.study_metrics([torch.from_numpy(r.copy()) for r in rows])
  assert_bit_exact(mean.numpy(), avg)
  _same_scalar(length, norm_avg, "norm_avg")
  _same_scalar(largest, norm_max, "norm_max")
  if len(rows) >= 2:
    _same_scalar(math.sqrt(spread / (len(rows) - 1)), norm_dev, "norm_dev")
