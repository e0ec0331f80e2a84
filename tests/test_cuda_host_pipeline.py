# coding: utf-8
""" `bz_coordinate_host` (host rows in, host vector out, column chunks with the copies and the kernels
overlapped): the reference's `--device-gar` hop (attack.py:811-815, 824-827) for the coordinate-wise
rules.  Coordinates are independent, so the result must be BIT-identical to the whole-vector call on
device rows — for every rule, ragged d, fewer coordinates than chunks, repeated rows, pageable memory. """

import ctypes

import pytest
import torch

import byzantinemomentum_b200 as bz
from byzantinemomentum_b200 import _lib, engine

pytestmark = pytest.mark.gpu
RULES = [("average", None), ("median", None), ("trmean", 3), ("phocas", 3), ("meamed", 3)]

@pytest.fixture(autouse=True)
def pipeline_path():
  before = engine.forced_host_path
  engine.forced_host_path = "pipeline"
  engine._host_paths.clear()
  yield
  engine.forced_host_path = before
  engine._host_paths.clear()

def _call(name, rows, f):
  return getattr(engine, name)(rows) if f is None else getattr(engine, name)(rows, f)

def _run_until_pipeline(name, rows, f):
  outs = [_call(name, rows, f) for _ in range(2)]
  report = engine.host_path_report()
  assert report["single_pass=True,pinned=False"]["best"] == "pipeline", report
  return outs

@pytest.mark.parametrize("name,f", RULES)
@pytest.mark.parametrize("n,d,pinned", [(11, 79510, True), (25, 1310922, True), (9, 1000, False), (7, 63, True), (12, 513, True)])
def test_pipeline_equals_the_device_call_bit_for_bit(name, f, n, d, pinned):
  gen = torch.Generator().manual_seed(n * 1000 + d)
  host = [torch.randn(d, generator=gen) for _ in range(n - 2)]
  host = host + [host[0], host[3]]                             # repeated tensors are staged once
  if pinned:
    host = [h.pin_memory() if i < n - 2 else h for i, h in enumerate(host)]
    host[-2], host[-1] = host[0], host[3]
  want = _call(name, [h.cuda() for h in host], f).cpu()
  for out in _run_until_pipeline(name, host, f):
    assert out.device.type == "cpu" and out.shape == (d,)
    assert torch.equal(out.view(torch.int32), want.view(torch.int32))

def test_pipeline_propagates_nan_and_inf_like_the_device_call():
  n, d = 11, 4099
  host = [torch.randn(d).pin_memory() for _ in range(n)]
  host[2][5] = float("nan"); host[4][7] = float("inf"); host[5][7] = float("-inf")
  want = engine.trmean([h.cuda() for h in host], 3).cpu()
  got = _run_until_pipeline("trmean", host, 3)[-1]
  assert torch.equal(got.view(torch.int32), want.view(torch.int32))

def test_c_entry_serial_streams_and_one_chunk():
  """ in_stream = out_stream = stream and chunks = 1: the plain staged call, same bits. """
  n, d = 5, 1031
  host = [torch.randn(d).pin_memory() for _ in range(n)]
  want = engine.median([h.cuda() for h in host]).cpu()
  staging = torch.empty((n, 1088), device="cuda:0")
  dev_out = torch.empty(d, device="cuda:0")
  result = torch.empty(d).pin_memory()
  ptrs = (ctypes.c_void_p * n)(*[h.data_ptr() for h in host])
  st = torch.cuda.current_stream().cuda_stream
  for chunks in (1, 3, 32):
    result.zero_()
    code = _lib.lib().bz_coordinate_host(1, ptrs, n, 0, d, result.data_ptr(), staging.data_ptr(), 1088, dev_out.data_ptr(), chunks, st, st, st)
    _lib.check(code, "bz_coordinate_host")
    torch.cuda.synchronize()
    assert torch.equal(result, want)

@pytest.mark.parametrize("name,f", [("krum", 3), ("bulyan", 2), ("cge", 3), ("aksel", 3)])
def test_batched_staging_equals_the_device_call(name, f):
  """ Distance-based rules on host rows: `bz_stage_rows` (one batched copy) instead of n copies. """
  engine.forced_host_path = "batch"
  engine._host_paths.clear()
  n, d = 11, 79510
  host = [torch.randn(d).pin_memory() for _ in range(n - 1)]
  host.append(host[2])
  want = bz.gars[name].unchecked(gradients=[h.cuda() for h in host], f=f).cpu()
  for _ in range(2):
    got = bz.gars[name].unchecked(gradients=host, f=f)
    assert got.device.type == "cpu" and torch.equal(got.view(torch.int32), want.view(torch.int32))
  assert engine.host_path_report()["single_pass=False,pinned=False"]["best"] == "batch"
