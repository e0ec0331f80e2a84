# coding: utf-8
"""Host-side argument validation (CPU suite).  `engine._validate` runs before anything touches
CUDA, and the C entry points check their arguments before the first CUDA call, so both are
testable without a GPU: invalid calls must fail with the documented error, never reach a kernel."""

import ctypes

import pytest

torch = pytest.importorskip("torch")

from byzantinemomentum_b200 import _lib, engine

def test_validate_rejects_malformed_gradient_lists():
  ok = [torch.zeros(8) for _ in range(3)]
  first, contiguous = engine._validate(ok)
  assert first is ok[0] and contiguous
  with pytest.raises(ValueError):
    engine._validate([])
  with pytest.raises(ValueError):
    engine._validate("not a list")
  with pytest.raises(TypeError):
    engine._validate([1.0, 2.0])
  with pytest.raises(TypeError):
    engine._validate([torch.zeros(8, dtype=torch.float64)] * 2)
  with pytest.raises(TypeError):
    engine._validate([torch.zeros(8), torch.zeros(8, dtype=torch.float16)])
  with pytest.raises(ValueError):
    engine._validate([torch.zeros(8), torch.zeros(9)])
  with pytest.raises(ValueError):
    engine._validate([torch.zeros(2, 4)])
  with pytest.raises(ValueError):
    engine._validate([torch.zeros(4)] * (_lib.MAX_N + 1))
  strided = torch.zeros(8, 3)[:, 0]
  assert engine._validate([strided, strided])[1] is False

def test_aksel_mode_is_checked_before_the_device_is_touched():
  with pytest.raises(NotImplementedError):
    engine.aksel([torch.zeros(4)] * 5, 1, mode="zzz")

def _rows(n, address=0x1000):
  return (ctypes.c_void_p * n)(*[address + 0x100 * i for i in range(n)])

def test_c_abi_rejects_bad_arguments_without_a_gpu():
  """ Every refusal below happens in the argument checks of api.cu, before any CUDA call; the
  pointers are never dereferenced (they are fake addresses). """
  lib = _lib.lib()
  out = ctypes.c_void_p(0x9000)
  def message():
    return lib.bz_last_error().decode()
  assert lib.bz_median(None, 3, 16, out, None) == -1 and "rows is NULL" in message()
  assert lib.bz_median(_rows(3), 0, 16, out, None) == -1
  assert lib.bz_median(_rows(3), 65, 16, out, None) == -2 and "BZ_MAX_N" in message()
  assert lib.bz_median(_rows(3), 3, -1, out, None) == -1
  assert lib.bz_median(_rows(3), 3, 16, None, None) == -1 and "output" in message()
  assert lib.bz_median(_rows(3, address=0x1002), 3, 16, out, None) == -1 and "aligned" in message()
  assert lib.bz_trmean(_rows(5), 5, 3, 16, out, None) == -1 and "n - 2f" in message()
  assert lib.bz_trmean(_rows(5), 5, -1, 16, out, None) == -1
  # d == 0 is a valid no-op for the coordinate-wise rules
  assert lib.bz_median(_rows(3), 3, 0, None, None) == 0
  assert lib.bz_trmean(_rows(5), 5, 2, 0, None, None) == 0
  # workspace: too small / misaligned
  need = lib.bz_workspace_bytes(7)
  assert need >= 7 * 7 * 8
  order = ctypes.c_void_p(0xA000)
  assert lib.bz_krum(_rows(7), 7, 2, 3, 16, out, order, ctypes.c_void_p(0xB000), need - 8, None) == -4 and "workspace" in message()
  assert lib.bz_krum(_rows(7), 7, 2, 3, 16, out, order, ctypes.c_void_p(0xB004), need, None) == -4
  assert lib.bz_avg_dev_max(_rows(7), 7, 16, out, None, ctypes.c_void_p(0xB000), need, None) == -1 and "stats" in message()
  assert lib.bz_aksel(_rows(7), 7, 2, 9, 16, out, order, ctypes.c_void_p(0xB000), need, None) == -1 and "mode" in message()

def test_check_maps_codes_to_exceptions():
  lib = _lib.lib()
  code = lib.bz_median(None, 3, 16, None, None)
  with pytest.raises(ValueError):
    _lib.check(code, "bz_median")
  _lib.check(0, "anything")
