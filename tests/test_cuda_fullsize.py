# coding: utf-8
"""GPU parity, part 4: BASELINE.json's five configurations at FULL size (single-GPU shapes),
checked against the compiled C oracle (`oracle/c`, threaded, seconds per case) — bit-exact for
the coordinate-wise rules and the ordered-subset means, selections equal — plus
size-independent properties: permutation invariance of the worker order, idempotence on equal
rows, bounds (min <= median <= max), and agreement of the selection with distances recomputed
in fp64 on the device."""

import numpy as np
import pytest

import parity
from oracle import byzoracle as orc
from oracle import corc

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not corc.available(), reason="oracle/c/libbyzoracle.so not built")]
DEV = "cuda:0"

def _stack(n, nb, d, seed, attack="empire"):
  gen = torch.Generator(device=DEV).manual_seed(seed)
  nh = n - nb
  mu = torch.randn(d, device=DEV, generator=gen)
  honest = [mu + (0.5 + i / max(nh - 1, 1)) * torch.randn(d, device=DEV, generator=gen) for i in range(nh)]
  rows = list(honest)
  if nb:
    byz = torch.stack(honest).mean(dim=0).mul(-1.1) if attack == "empire" else torch.full((d,), float("nan"), device=DEV)
    rows += [byz] * nb
  return rows

def _host(rows):
  cache = {}
  out = []
  for r in rows:
    if id(r) not in cache:
      cache[id(r)] = r.cpu().numpy()
    out.append(cache[id(r)])
  return out

def test_c1_median_n11_f5_d79510():
  import byzantinemomentum_b200 as bz
  rows = _stack(11, 0, 79_510, 1)
  host = _host(rows)
  parity.assert_bit_exact(bz.gars["median"](gradients=rows, f=5).cpu().numpy(), corc.median(host), "C1 median")

def test_c2_trmean_n25_f10_d1310922():
  import byzantinemomentum_b200 as bz
  n, f, d = 25, 10, 1_310_922
  rows = _stack(n, 0, d, 2)
  host = _host(rows)
  got = bz.gars["trmean"](gradients=rows, f=f)
  parity.assert_bit_exact(got.cpu().numpy(), corc.trmean(host, f), "C2 trmean")
  # properties: invariant under a permutation of the workers; bounded by the column range
  perm = torch.randperm(n, generator=torch.Generator().manual_seed(0)).tolist()
  assert torch.equal(bz.gars["trmean"](gradients=[rows[i] for i in perm], f=f), got)
  stacked = torch.stack(rows)
  assert bool((got >= stacked.min(dim=0).values).all()) and bool((got <= stacked.max(dim=0).values).all())
  med = bz.gars["median"](gradients=rows, f=f)
  assert torch.equal(med, stacked.median(dim=0).values)          # library kernel as a second opinion
  # phocas / meamed at full size: 1e-6 of the summed magnitude, exact key ties exempt
  x = np.stack(host)
  for name, center in (("phocas", corc.trmean(host, f)), ("meamed", corc.median(host))):
    ref = getattr(corc, name)(host, f)
    amb = parity.closest_ambiguous(x, n - f, center)
    parity.assert_close_scaled(bz.gars[name](gradients=rows, f=f).cpu().numpy(), ref, parity.column_scale(x), "C2 " + name, exempt=amb)
  # idempotence: n equal rows aggregate to that row (trimmed mean of equal values, IEEE: x*R/R = x exactly for R <= 5)
  same = [rows[0]] * n
  assert torch.equal(bz.gars["median"](gradients=same, f=f), rows[0])

def test_c3_krum_bulyan_n25_f5_d1310922_empire():
  import byzantinemomentum_b200 as bz
  n, nb, f, d = 25, 5, 5, 1_310_922
  rows = _stack(n, nb, d, 3, "empire")
  host = _host(rows)
  D = corc.pairwise_distances(host)
  order, scores = orc.krum_order(D, f)
  m = n - f - 2
  out = bz.gars["krum"](gradients=rows, f=f)
  sel = bz.last_selection()
  canon = lambda idx: [min(int(i), n - nb) for i in idx]
  assert canon(sel[:m]) == canon(order[:m])
  parity.assert_bit_exact(out.cpu().numpy(), corc.average_selected(host, order[:m]), "C3 krum")
  assert bz.gars["krum"].influence(rows[:n - nb], rows[n - nb:], f=f) == sum(1 for i in order[:m] if i >= n - nb) / m
  # distances recomputed on the device in fp64 agree with the kernel's to fp32 rounding
  from byzantinemomentum_b200 import engine
  part = engine.pairdist_partial(rows).cpu().numpy()
  x64 = torch.stack(rows[:n - nb + 1]).double()
  ref = torch.cdist(x64, x64).cpu().numpy()
  got = np.sqrt(part[:n - nb + 1, :n - nb + 1])
  iu = np.triu_indices(n - nb + 1, 1)
  np.testing.assert_allclose(got[iu], ref[iu], rtol=5e-7)
  # bulyan: stage-1 order from the same distances, then the oracle's stage 2 on sampled columns
  outb = bz.gars["bulyan"](gradients=rows, f=f).cpu().numpy()
  cols = np.random.default_rng(0).choice(d, size=20_000, replace=False)
  sub = [h[cols] for h in host]
  border, _ = orc.bulyan_order(D, f, m)
  stage1 = orc.bulyan_stage1(orc.as_matrix(sub), border, f, m)
  theta = stage1.shape[0]
  ref_b, amb = orc.closest_mean(stage1, theta - 2 * f, orc.median(stage1), return_info=True)
  parity.assert_close_scaled(outb[cols], ref_b, parity.column_scale(stage1), "C3 bulyan (sampled columns)", exempt=amb)

def test_c4_shard_median_trmean_n51_f12_d4568373():
  import byzantinemomentum_b200 as bz
  n, f, d = 51, 12, 4_568_373      # one of 8 shards of WideResNet-28-10 (36,546,980)
  rows = _stack(n, 0, d, 4)
  host = _host(rows)
  parity.assert_bit_exact(bz.gars["median"](gradients=rows, f=f).cpu().numpy(), corc.median(host), "C4 median")
  parity.assert_bit_exact(bz.gars["trmean"](gradients=rows, f=f).cpu().numpy(), corc.trmean(host, f), "C4 trmean")   # R = 27: cascade order

def test_c5_brute_n11_f3_d1310922():
  import byzantinemomentum_b200 as bz
  n, nb, f, d = 11, 3, 3, 1_310_922
  rows = _stack(n, nb, d, 5, "empire")
  host = _host(rows)
  D = corc.pairwise_distances(host, map_nonfinite=False)
  sel = orc.brute_selection(D, f)
  out = bz.gars["brute"](gradients=rows, f=f)
  canon = lambda idx: [min(int(i), n - nb) for i in idx]
  assert canon(bz.last_selection()) == canon(sel)
  parity.assert_bit_exact(out.cpu().numpy(), corc.average_selected(host, sel), "C5 brute")
  # NaN attack: the f Byzantine rows are excluded, the result is finite
  rows_nan = _stack(n, nb, d, 5, "nan")
  out = bz.gars["brute"](gradients=rows_nan, f=f)
  assert bool(torch.isfinite(out).all()) and all(i < n - nb for i in bz.last_selection())

def test_wideresnet_size_median_trmean_n25():
  """ d = 36,489,290 (WRN-28-10, CIFAR-10), the north-star size, against the C oracle on a column sample
  and torch's own sort on the full width. """
  import byzantinemomentum_b200 as bz
  n, f, d = 25, 10, 36_489_290
  gen = torch.Generator(device=DEV).manual_seed(6)
  rows = [torch.randn(d, device=DEV, generator=gen) for _ in range(n)]
  med = bz.gars["median"](gradients=rows, f=f)
  trm = bz.gars["trmean"](gradients=rows, f=f)
  cols = torch.from_numpy(np.random.default_rng(1).choice(d, size=200_000, replace=False)).to(DEV)
  sub = [r[cols].cpu().numpy() for r in rows]
  parity.assert_bit_exact(med[cols].cpu().numpy(), corc.median(sub), "WRN median (sample)")
  parity.assert_bit_exact(trm[cols].cpu().numpy(), corc.trmean(sub, f), "WRN trmean (sample)")
  # full width, in chunks, against the library sort (values[f:-f] summed ascending = sequential for R = 5)
  for lo in range(0, d, 8_000_000):
    hi = min(d, lo + 8_000_000)
    s = torch.stack([r[lo:hi] for r in rows]).sort(dim=0).values
    assert torch.equal(med[lo:hi], s[(n - 1) // 2])
    acc = torch.zeros(hi - lo, device=DEV)
    for k in range(f, n - f):
      acc = acc + s[k]
    # (a tensor divisor: torch turns division by a Python scalar into a multiplication by 1/x on CUDA)
    assert torch.equal(trm[lo:hi], acc / torch.full_like(acc, float(n - 2 * f)))
    del s, acc
