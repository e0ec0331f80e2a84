# coding: utf-8
"""GPU parity, part 3: the phases of the d-sharded path (A: partial distances, B: selection
from gathered blocks, C: local reduce) composed by hand over R shards on ONE GPU must give the
single-device result — what `byzantinemomentum_b200.sharded.aggregate` does across ranks with
one all-gather in the middle."""

import numpy as np
import pytest

import parity
from oracle import byzoracle as orc

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
DEV = "cuda:0"

def _inputs(n, nb, d, seed):
  gen = torch.Generator().manual_seed(seed)
  mu = torch.randn(d, generator=gen)
  honest = mu[None, :] + torch.linspace(0.5, 1.5, n - nb)[:, None] * torch.randn(n - nb, d, generator=gen)
  byz = honest.mean(dim=0).mul(-1.1)
  rows = [honest[i] for i in range(n - nb)] + [byz] * nb
  return rows

@pytest.mark.parametrize("n,nb,f,d,R", [(11, 3, 3, 4001, 2), (25, 5, 5, 10007, 4), (25, 5, 5, 10007, 8), (51, 12, 12, 3001, 3)])
def test_phases_compose_to_the_single_device_result(n, nb, f, d, R):
  from byzantinemomentum_b200 import engine
  host = _inputs(n, nb, d, 77 + n)
  np_rows = [r.numpy() for r in host]
  full = [r.to(DEV) for r in host]
  per = (d + R - 1) // R
  shards = [[r[k * per:min(d, (k + 1) * per)] for r in full] for k in range(R)]    # unaligned views on purpose
  def cat(outs):
    return torch.cat(outs).cpu().numpy()
  # krum
  parts = torch.stack([engine.pairdist_partial(s) for s in shards])
  order = engine.krum_select(parts, n, f)
  ref, info = orc.krum(np_rows, f, return_info=True)
  m = n - f - 2
  assert [min(i, n - nb) for i in order.cpu().tolist()[:m]] == [min(i, n - nb) for i in info["selection"]]
  parity.assert_bit_exact(cat([engine.average_selected(s, order, m) for s in shards]), ref, "krum sharded")
  # phase B reading the R blocks in place through a pointer per "peer" (here: R buffers on one GPU)
  blocks = [parts[k].contiguous() for k in range(R)]
  order_p = engine.krum_select_peers([b.data_ptr() for b in blocks], n, f, full[0].device)
  assert order_p.cpu().tolist() == order.cpu().tolist()
  # same selection as the single-device call
  out1, order1 = engine.krum(full, f, m)
  assert order1.cpu().tolist()[:m] == order.cpu().tolist()[:m]
  parity.assert_bit_exact(out1.cpu().numpy(), ref, "krum single")
  # bulyan
  if n >= 4 * f + 3:
    order_b, status = engine.bulyan_select(parts, n, f, m)
    assert int(status.item()) == 0
    order_bp, status_p = engine.bulyan_select_peers([b.data_ptr() for b in blocks], n, f, m, full[0].device)
    assert order_bp.cpu().tolist() == order_b.cpu().tolist() and int(status_p.item()) == 0
    refb, infob = orc.bulyan(np_rows, f, return_info=True)
    got = cat([engine.bulyan_reduce(s, f, m, order_b, status) for s in shards])
    parity.assert_close_scaled(got, refb, parity.column_scale(infob["stage1"]), "bulyan sharded", exempt=infob["ambiguous"])
  # brute (small n only)
  if n <= 11:
    sel, status = engine.brute_select(parts, n, f)
    sel_p, _ = engine.brute_select_peers([b.data_ptr() for b in blocks], n, f, full[0].device)
    assert sel_p.cpu().tolist() == sel.cpu().tolist()
    refr, infor = orc.brute(np_rows, f, return_info=True)
    assert int(status.item()) == 0 and [min(i, n - nb) for i in sel.cpu().tolist()] == [min(i, n - nb) for i in infor["selection"]]
    parity.assert_bit_exact(cat([engine.average_selected(s, sel, n - f, status=status) for s in shards]), refr, "brute sharded")
  # cge
  pn = torch.stack([engine.rowdist_partial(s) for s in shards])
  order_c = engine.rowdist_select(pn, n, True)
  assert engine.rowdist_select_peers([pn[k].contiguous().data_ptr() for k in range(R)], n, True, full[0].device).cpu().tolist() == order_c.cpu().tolist() if R <= 16 else True
  refc = orc.cge(np_rows, f)
  parity.assert_bit_exact(cat([engine.average_selected(s, order_c, n - f, zero_init=False) for s in shards]), refc, "cge sharded")
  # aksel: the median is coordinate-local
  meds = [engine.median(s) for s in shards]
  pa = torch.stack([engine.rowdist_partial(s, c) for s, c in zip(shards, meds)])
  order_a = engine.rowdist_select(pa, n, False)
  refa, infoa = orc.aksel(np_rows, f, return_info=True)
  c = (n + 1) // 2
  assert [min(i, n - nb) for i in order_a.cpu().tolist()[:c]] == [min(i, n - nb) for i in infoa["selection"]]
  parity.assert_bit_exact(cat([engine.average_selected(s, order_a, c) for s in shards]), refa, "aksel sharded")

def test_sharded_aggregate_world_size_one():
  """ `sharded.aggregate` without an initialised process group = one rank holding everything. """
  from byzantinemomentum_b200 import sharded
  n, nb, f, d = 11, 3, 3, 5003
  host = _inputs(n, nb, d, 5)
  np_rows = [r.numpy() for r in host]
  rows = [r.to(DEV) for r in host]
  for gar, params in (("median", {}), ("trmean", dict(f=f)), ("krum", dict(f=f)), ("brute", dict(f=f)), ("cge", dict(f=f)), ("aksel", dict(f=f))):
    got = sharded.aggregate(gar, rows, **params).cpu().numpy()
    parity.assert_bit_exact(got, orc.GARS[gar](np_rows, **params), gar)

def test_plan_for_every_rule():
  import byzantinemomentum_b200 as bz
  n, nb, f, d = 11, 2, 2, 6007
  host = _inputs(n, nb, d, 31)
  np_rows = [r.numpy() for r in host]
  byz = host[-1].to(DEV)
  rows = [r.to(DEV) for r in host[:n - nb]] + [byz] * nb
  for gar, params in (("average", {}), ("median", {}), ("trmean", dict(f=f)), ("phocas", dict(f=f)), ("meamed", dict(f=f)),
                      ("krum", dict(f=f)), ("bulyan", dict(f=f)), ("brute", dict(f=f)), ("aksel", dict(f=f)), ("cge", dict(f=f))):
    plan = bz.Plan(gar, rows, **params)
    got = plan().cpu().numpy()
    ref = orc.GARS[gar](np_rows, **params)
    if gar in ("phocas", "meamed", "bulyan"):
      parity.assert_close_scaled(got, ref, parity.column_scale(np.stack(np_rows)), gar + " plan")
    else:
      parity.assert_bit_exact(got, ref, gar + " plan")
    again = plan().cpu().numpy()
    assert np.array_equal(got, again, equal_nan=True)      # deterministic, run to run
    if plan.status is not None:
      assert int(plan.status.item()) == 0

def test_plan_matches_plain_call_and_tracks_in_place_updates():
  import byzantinemomentum_b200 as bz
  n, f, d = 25, 10, 20003
  x = torch.randn(n, d, generator=torch.Generator().manual_seed(9))
  rows = [x[i].to(DEV) for i in range(n)]
  plan = bz.Plan("trmean", rows, f=f)
  parity.assert_bit_exact(plan().cpu().numpy(), orc.trmean([x[i].numpy() for i in range(n)], f), "plan")
  rows[3].mul_(-2.5)          # momentum buffers are updated in place between steps
  x[3].mul_(-2.5)
  parity.assert_bit_exact(plan().cpu().numpy(), orc.trmean([x[i].numpy() for i in range(n)], f), "plan after update")
  kp = bz.Plan("krum", rows, f=5)
  out = kp().cpu().numpy()
  ref, info = orc.krum([x[i].numpy() for i in range(n)], 5, return_info=True)
  assert kp.selection.cpu().tolist()[:n - 7] == info["selection"]
  parity.assert_bit_exact(out, ref, "krum plan")

def test_plan_rejects_strided_rows_but_the_plain_call_packs_them():
  import byzantinemomentum_b200 as bz
  x = torch.randn(64, 7, generator=torch.Generator().manual_seed(12)).to(DEV)
  columns = [x[:, i] for i in range(7)]               # 1-D views with stride 7
  assert not columns[0].is_contiguous()
  with pytest.raises(ValueError):
    bz.Plan("median", columns)
  got = bz.gars["median"].unchecked(gradients=columns)
  parity.assert_bit_exact(got.cpu().numpy(), orc.median([c.cpu().numpy() for c in columns]), "median of strided views")

@pytest.mark.parametrize("gar,kw", [("krum", dict(f=5)), ("krum", dict(f=5, m=3)), ("bulyan", dict(f=5)), ("aksel", dict(f=5)),
                                    ("aksel", dict(f=5, mode="n-f")), ("cge", dict(f=5)), ("median", dict(f=5)), ("trmean", dict(f=5)), ("average", {})])
def test_sharded_plan_on_one_rank_equals_the_plain_call(gar, kw):
  """ `sharded.ShardedPlan` (the prepared three-phase call) without a process group = world of 1:
  the all-gather degenerates to a copy, everything else is the code path of the ranks.  Must
  reproduce the single-device rule bit for bit (same kernels, same selection), call after call,
  and follow in-place updates of the rows. """
  import byzantinemomentum_b200 as bz
  from byzantinemomentum_b200 import sharded
  n, nb, d = 25, 5, 20011
  rows = [r.to(DEV) for r in _inputs(n, nb, d, 91)]
  plan = sharded.ShardedPlan(gar, rows, **kw)
  want = bz.gars[gar](gradients=rows, **({"f": 5} | kw))
  for _ in range(3):
    got = plan()
  assert torch.equal(got, want) or gar == "bulyan" and torch.allclose(got, want, rtol=0, atol=0)
  if gar in ("krum", "aksel", "cge"):
    assert plan.selection.cpu().tolist() == bz.last_selection()
  rows[0].mul_(3.)                       # content changes, pointers do not
  want = bz.gars[gar](gradients=rows, **({"f": 5} | kw))
  assert torch.equal(plan(), want)

def test_sharded_plan_brute_and_status():
  import byzantinemomentum_b200 as bz
  from byzantinemomentum_b200 import sharded
  rows = [r.to(DEV) for r in _inputs(11, 3, 5003, 92)]
  plan = sharded.ShardedPlan("brute", rows, f=3)
  want = bz.gars["brute"](gradients=rows, f=3)
  assert torch.equal(plan(), want) and int(plan.status.item()) == 0
  assert plan.selection.cpu().tolist() == bz.last_selection()

@pytest.mark.parametrize("gar,kw", [("krum", dict(f=5)), ("bulyan", dict(f=5)), ("cge", dict(f=5)), ("aksel", dict(f=5)), ("trmean", dict(f=5))])
def test_plan_graph_replays_the_same_result(gar, kw):
  """ `Plan.graph()`: the rule captured once in a CUDA graph (memset, fused distance + scoring pass,
  PDL-launched reduce pass) gives what the direct call gives, replay after replay, and follows
  in-place updates of the rows. """
  import byzantinemomentum_b200 as bz
  rows = [r.to(DEV) for r in _inputs(25, 5, 30011, 93)]
  plan = bz.Plan(gar, rows, **kw)
  want = plan().clone()
  replay = plan.graph()
  for _ in range(3):
    assert torch.equal(replay(), want)
  rows[2].mul_(-2.)
  want = bz.gars[gar](gradients=rows, **kw)
  assert torch.equal(replay(), want)
