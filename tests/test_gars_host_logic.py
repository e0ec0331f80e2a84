# coding: utf-8
"""Host logic of `byzantinemomentum_b200.gars` with the engine replaced by a NumPy stand-in
(CPU suite): the selection computed by an aggregation is reused by `influence()` on the same
tensors (attack.py:821 is always followed by :822) and recomputed when anything changed; the
influence values equal the oracle's (= the reference's, see the goldens)."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import importlib
gars_module = importlib.import_module("byzantinemomentum_b200.gars")   # the package attribute `gars` is the registry dict
from oracle import byzoracle as orc

class CountingEngine:
  """ Oracle-backed engine recording how often each rule ran. """
  def __init__(self):
    self.calls = []
  @staticmethod
  def _np(rows):
    return [r.numpy() for r in rows]
  def krum(self, rows, f, m):
    self.calls.append("krum")
    out, info = orc.krum(self._np(rows), f, m=m, return_info=True)
    return torch.from_numpy(out), torch.tensor(info["order"], dtype=torch.int32)
  def brute(self, rows, f):
    self.calls.append("brute")
    out, info = orc.brute(self._np(rows), f, return_info=True)
    return torch.from_numpy(out), torch.tensor(info["selection"], dtype=torch.int32)
  def aksel(self, rows, f, mode="mid"):
    self.calls.append("aksel")
    out, info = orc.aksel(self._np(rows), f, mode=mode, return_info=True)
    return torch.from_numpy(out), torch.tensor(info["order"], dtype=torch.int32)
  def cge(self, rows, f):
    self.calls.append("cge")
    out, info = orc.cge(self._np(rows), f, return_info=True)
    return torch.from_numpy(out), torch.tensor(info["order"], dtype=torch.int32)

@pytest.fixture
def engine(monkeypatch):
  fake = CountingEngine()
  monkeypatch.setattr(gars_module, "engine", fake)
  monkeypatch.setattr(gars_module, "_last", gars_module._Selection())
  return fake

def _stack(n=11, nb=3, d=257, seed=5):
  gen = torch.Generator().manual_seed(seed)
  mu = torch.randn(d, generator=gen)
  honests = [mu + (0.5 + i / n) * torch.randn(d, generator=gen) for i in range(n - nb)]
  attack = torch.stack(honests).mean(dim=0).mul(-1.1)
  return honests, [attack] * nb

@pytest.mark.parametrize("name,params", [("krum", dict(f=3)), ("krum", dict(f=3, m=2)), ("brute", dict(f=3)),
                                         ("aksel", dict(f=3)), ("aksel", dict(f=3, mode="n-f")), ("cge", dict(f=3))])
def test_influence_reuses_the_selection_of_the_aggregation(engine, name, params):
  honests, attacks = _stack()
  rule = gars_module.gars[name]
  rule.unchecked(gradients=honests + attacks, **params)
  assert engine.calls == [name]
  value = rule.influence(honests, attacks, **params)
  assert engine.calls == [name], "influence() recomputed the selection of the call just made"
  reference = orc.influence(name, [h.numpy() for h in honests], [a.numpy() for a in attacks], **params)
  assert value == pytest.approx(reference, abs=0)
  assert gars_module.last_selection() is not None

def test_influence_recomputes_when_the_inputs_changed(engine):
  honests, attacks = _stack()
  rule = gars_module.gars["krum"]
  rule.unchecked(gradients=honests + attacks, f=3)
  honests[2].mul_(1.5)                                    # in-place update bumps the tensor version
  rule.influence(honests, attacks, f=3)
  assert engine.calls == ["krum", "krum"]
  rule.unchecked(gradients=honests + attacks, f=3)
  rule.influence(honests, attacks, f=2)                   # other parameters: not the cached selection
  assert engine.calls == ["krum"] * 4
  other_h, other_a = _stack(seed=6)
  rule.influence(other_h, other_a, f=3)                   # other tensors
  assert engine.calls == ["krum"] * 5

def test_influence_without_a_previous_aggregation(engine):
  honests, attacks = _stack()
  value = gars_module.gars["cge"].influence(honests, attacks, f=3)
  assert engine.calls == ["cge"]
  assert 0. <= value <= 1.

def test_average_influence_and_rules_without_influence():
  assert gars_module.gars["average"].influence([1, 2, 3], [4]) == 0.25
  for name in ("median", "trmean", "phocas", "meamed", "bulyan"):
    assert gars_module.gars[name].influence is None      # as in the reference

def test_checked_raises_the_user_exception_with_the_rule_name():
  with pytest.raises(gars_module.UserException) as err:
    gars_module.gars["krum"].checked(gradients=[torch.zeros(4)] * 5, f=2)
  assert "krum" in str(err.value) and "f = 2" in str(err.value)
