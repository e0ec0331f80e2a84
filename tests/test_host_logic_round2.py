# coding: utf-8
"""Host-side logic added in round 2 that needs no GPU: the measured choice of the host->device path, the
study-metrics memo, the identity/version keyed selection cache, the shard bounds of the sharded update."""

import gc
import importlib

import pytest

torch = pytest.importorskip("torch")

engine = importlib.import_module("byzantinemomentum_b200.engine")
gars = importlib.import_module("byzantinemomentum_b200.gars")
sharded = importlib.import_module("byzantinemomentum_b200.sharded")

def test_host_path_tries_every_candidate_three_times_then_keeps_the_fastest():
  path = engine._HostPath()
  seen = []
  costs = {"lanes": 3.0, "lane": 2.0, "pipeline": 2.5}
  for _ in range(9):
    mode = path.choose(True, False)
    seen.append(mode)
    path.record(True, False, mode, costs[mode])
  assert sorted(seen) == ["lane"] * 3 + ["lanes"] * 3 + ["pipeline"] * 3      # coordinate-wise rules: three candidates
  assert all(path.choose(True, False) == "lane" for _ in range(5))
  path.record(True, False, "lane", 99.)             # later samples do not reopen the decision
  assert path.choose(True, False) == "lane"
  seen = []
  for _ in range(9):                                # another kind of call is measured on its own, without the pipeline
    mode = path.choose(False, False)
    seen.append(mode)
    path.record(False, False, mode, 1.0)
  assert sorted(seen) == ["batch"] * 3 + ["lane"] * 3 + ["lanes"] * 3

def test_host_path_ranks_by_the_slower_of_the_late_samples():
  path = engine._HostPath()
  samples = {"lanes": [9., 3., 3.], "lane": [9., 2.6, 2.7], "pipeline": [9., 2.4, 3.5]}      # fast once, slow the next time
  for _ in range(9):
    mode = path.choose(True, False)
    path.record(True, False, mode, samples[mode].pop(0))
  assert path.choose(True, False) == "lane"

def test_forced_host_path_is_used_without_sampling():
  path = engine._HostPath()
  engine.forced_host_path = "pipeline"
  try:
    assert path.choose(True, False) == "pipeline"
    assert path.choose(False, False) == "lanes"     # not a candidate there: measured as usual
  finally:
    engine.forced_host_path = None

def test_study_memo_matches_only_the_same_unmodified_tensor_objects():
  memo = engine._StudyMemo()
  rows = [torch.arange(4.), torch.ones(4)]
  result = (torch.full((4,), 2.), 1., 2., 3.)
  assert memo.lookup(rows) is None
  memo.store(rows, result)
  hit = memo.lookup(list(rows))                      # another list object, the same tensors
  assert hit is not None and torch.equal(hit[0], result[0]) and hit[1:] == (1., 2., 3.)
  assert hit[0] is not result[0]                     # callers own the average they get
  assert memo.lookup(rows[:1]) is None
  assert memo.lookup([rows[0], torch.ones(4)]) is None          # equal values, another object
  rows[1].add_(1.)                                   # in-place update bumps the version
  assert memo.lookup(rows) is None

def test_selection_cache_never_serves_a_recycled_identity():
  sel = gars._Selection()
  rows = [torch.zeros(3) for _ in range(4)]
  sel.store("krum", (1,), rows, "indices")
  assert sel.lookup("krum", (1,), rows) == "indices"
  assert sel.lookup("krum", (2,), rows) is None and sel.lookup("brute", (1,), rows) is None
  rows[2].mul_(2.)
  assert sel.lookup("krum", (1,), rows) is None
  sel.store("krum", (1,), rows, "again")
  del rows
  gc.collect()
  fresh = [torch.zeros(3) for _ in range(4)]         # may reuse ids / addresses of the dead tensors, at version 0
  assert sel.lookup("krum", (1,), fresh) is None
  assert sel.refs is None                            # a dead reference dropped the entry

def test_shard_bounds_cover_the_vector_without_overlap():
  for d, world in ((601, 2), (600, 2), (36_546_980, 8), (7, 8), (1, 4)):
    spans = [sharded.shard_bounds(d, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == d
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert all(lo <= hi for lo, hi in spans)
