# coding: utf-8
""" SURVEY.md §8(f) row 3 (the study step, attack.py:846-866): `bz_rowdots` (one vector against
many, one pass) and `engine.study_step` (the whole block with ONE host read) against a plain
restatement of the reference's statements with library calls. """

import math

import pytest
import torch

import byzantinemomentum_b200 as bz
from byzantinemomentum_b200 import engine

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

def _rows(n, d, seed, offset=0):
  gen = torch.Generator(device="cpu").manual_seed(seed)
  out = []
  for i in range(n):
    buf = torch.randn(d + 8, generator=gen).to(DEV)
    out.append(buf[(offset * i) % 4:(offset * i) % 4 + d] if offset else buf[:d].clone())
  return out

@pytest.mark.parametrize("n,d,offset", [(1, 1, 0), (3, 7, 0), (5, 1000, 1), (25, 79510, 0), (33, 4099, 0), (64, 20001, 1), (7, 1310922, 0)])
def test_rowdots_matches_fp64_dot_products(n, d, offset):
  rows = _rows(n, d, 100 + n, offset)
  center = torch.randn(d, device=DEV)
  got = engine.rowdots_async(center, rows).tolist()
  for i, row in enumerate(rows):
    ref = torch.dot(row.double(), center.double()).item()
    scale = torch.dot(row.double().abs(), center.double().abs()).item()
    assert abs(got[i] - ref) <= 2e-7 * scale + 1e-30, (i, got[i], ref)

def test_rowdots_more_rows_than_one_launch_holds_and_deterministic():
  rows = _rows(70, 513, 5)
  center = torch.randn(513, device=DEV)
  a = engine.rowdots_async(center, rows)
  b = engine.rowdots_async(center, rows)
  assert a.shape == (70,) and torch.equal(a, b)
  ref = torch.stack([torch.dot(r.double(), center.double()) for r in rows])
  assert torch.allclose(a, ref, rtol=0, atol=1e-4)

def test_rowdots_rejects_a_center_of_another_shape():
  with pytest.raises(ValueError):
    engine.rowdots_async(torch.randn(9, device=DEV), _rows(2, 8, 1))

def _study_reference(sampleds, honests, attacks, defense, pasts, momentum):
  """ attack.py:846-866 with library calls (fp32 tensors, fp32 `.item()`s). """
  def avg_dev_max(samples):      # tools/pytorch.py:97-125
    if len(samples) == 0:
      return None, math.nan, math.nan, math.nan
    avg = torch.stack(samples).double().mean(dim=0).float()
    norm = avg.norm().item()
    dev = math.sqrt(sum((s - avg).double().pow(2).sum().item() for s in samples) / (len(samples) - 1)) if len(samples) > 1 else math.nan
    return avg, norm, dev, avg.abs().max().item()
  s_avg, s_norm, s_dev, s_max = avg_dev_max(sampleds)
  h_avg, h_norm, h_dev, h_max = avg_dev_max(honests)
  a_avg, a_norm, a_dev, a_max = avg_dev_max(attacks)
  d_norm = defense.norm().item()
  cos = lambda a, b, na, nb: math.nan if a is None or b is None else torch.dot(a.double(), b.double()).item() / na / nb
  out = dict(sampled_norm_avg=s_norm, sampled_norm_dev=s_dev, sampled_norm_max=s_max, honest_norm_avg=h_norm, honest_norm_dev=h_dev,
             honest_norm_max=h_max, attack_norm_avg=a_norm, attack_norm_dev=a_dev, attack_norm_max=a_max,
             defense_norm_avg=d_norm, defense_norm_max=defense.abs().max().item(),
             cosin_splhon=cos(s_avg, h_avg, s_norm, h_norm), cosin_splatt=cos(s_avg, a_avg, s_norm, a_norm),
             cosin_spldef=cos(s_avg, defense, s_norm, d_norm), cosin_honatt=cos(h_avg, a_avg, h_norm, a_norm),
             cosin_hondef=cos(h_avg, defense, h_norm, d_norm), cosin_attdef=cos(a_avg, defense, a_norm, d_norm))
  if pasts:
    out["cosin_sampled"] = cos(s_avg, pasts[0][0], s_norm, pasts[0][1])
    out["curv_sampled"] = momentum * sum(momentum ** i * torch.dot(s_avg.double(), g.double()).item() for i, (g, _) in enumerate(pasts))
  else:
    out["cosin_sampled"] = out["curv_sampled"] = math.nan
  return out, s_avg

@pytest.mark.parametrize("ns,nh,na,npast,d,shared", [(8, 8, 3, 2, 79510, True), (8, 8, 3, 0, 79510, False), (8, 6, 0, 3, 4099, False),
                                                     (1, 1, 1, 1, 33, True), (20, 20, 5, 5, 1310922, False)])
def test_study_step_matches_the_statements_of_attack_py(ns, nh, na, npast, d, shared):
  sampleds = _rows(ns, d, 11)
  honests = sampleds[:nh] if shared else _rows(nh, d, 12)
  attacks = _rows(na, d, 13)
  defense = torch.randn(d, device=DEV) * 0.3
  pasts = []
  for g in _rows(npast, d, 14):
    pasts.append((g, g.norm().item()))
  want, want_avg = _study_reference(sampleds, honests, attacks, defense, pasts, 0.9)
  before = torch.cuda.current_stream().query()
  got = engine.study_step(sampleds, honests, attacks, defense, pasts, 0.9)
  assert torch.allclose(got["sampled_grad_avg"], want_avg, rtol=1e-6, atol=1e-7)
  assert (got["attack_grad_avg"] is None) == (na == 0)
  for key, ref in want.items():
    val = got[key]
    if math.isnan(ref):
      assert math.isnan(val), key
    else:
      assert abs(val - ref) <= 2e-5 * abs(ref) + 3e-6, (key, val, ref)
  assert set(want) <= set(got)

def test_study_step_zero_vectors_follow_ieee_division():
  d = 257
  zeros = [torch.zeros(d, device=DEV) for _ in range(3)]
  got = engine.study_step(zeros, zeros, [], torch.zeros(d, device=DEV), [], 0.9)
  assert got["sampled_norm_avg"] == 0. and math.isnan(got["cosin_splhon"]) and math.isnan(got["cosin_splatt"]) and math.isnan(got["cosin_sampled"])
