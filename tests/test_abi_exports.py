# coding: utf-8
"""The C-ABI shared library loads without a GPU and exports exactly what `include/byzagg.h`
declares; the ctypes table in `byzantinemomentum_b200/_lib.py` covers every symbol.  No compute
calls here (CPU suite)."""

import ctypes
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent

def declared_symbols():
  text = (ROOT / "include" / "byzagg.h").read_text()
  return sorted(set(re.findall(r"BZ_API\s+[\w\s\*]+?\b(bz_\w+)\s*\(", text)))

def test_header_declares_the_expected_entry_points():
  names = declared_symbols()
  for must in ("bz_median", "bz_trmean", "bz_krum", "bz_bulyan", "bz_brute", "bz_aksel", "bz_cge", "bz_average",
               "bz_phocas", "bz_meamed", "bz_pairdist_partial", "bz_krum_select", "bz_average_selected", "bz_last_error"):
    assert must in names
  assert len(names) == 36

def test_library_loads_and_exports_every_symbol():
  from byzantinemomentum_b200 import _lib
  path = _lib.library_path()
  assert path.exists(), f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`"
  handle = ctypes.CDLL(str(path))
  for name in declared_symbols():
    assert hasattr(handle, name), f"{path.name} does not export {name}"
  assert sorted(_lib.SIGNATURES) == declared_symbols()
  lib = _lib.lib()
  assert lib.bz_version() >= 100 and lib.bz_max_n() == 64
  assert lib.bz_workspace_bytes(25) >= (160 + 1) * 25 * 25 * 8

def test_product_path_fails_loudly_without_gpu_or_library(monkeypatch, tmp_path):
  import torch
  import byzantinemomentum_b200 as bz
  from byzantinemomentum_b200 import _lib
  if not torch.cuda.is_available():
    with pytest.raises(_lib.LibraryError):
      bz.gars["median"](gradients=[torch.zeros(8)], f=1)
  # a missing library is an error, never a fallback
  monkeypatch.setenv("BYZAGG_LIBRARY", str(tmp_path / "nope.so"))
  monkeypatch.setattr(_lib, "_lib", None)
  with pytest.raises(_lib.LibraryError):
    _lib.lib()

def test_native_hook_module_shape():
  """ What the reference's loader looks for: `native_name in dir(native)` and `.aggregate`. """
  import native
  for name in ("median", "krum", "bulyan", "brute"):
    assert name in dir(native)
    assert callable(getattr(native, name).aggregate)

def test_checks_and_bounds_mirror_the_reference():
  import math
  import torch
  import byzantinemomentum_b200 as bz
  rows = [torch.zeros(4) for _ in range(11)]
  g = bz.gars
  assert set(g) == {"average", "median", "trmean", "phocas", "meamed", "krum", "bulyan", "brute", "aksel", "cge"}
  assert g["krum"].check(gradients=rows, f=4) is None and g["krum"].check(gradients=rows, f=5) is not None     # n >= 2f+3
  assert g["bulyan"].check(gradients=rows, f=2) is None and g["bulyan"].check(gradients=rows, f=3) is not None  # n >= 4f+3
  assert g["trmean"].check(gradients=rows, f=5) is None and g["trmean"].check(gradients=rows, f=6) is not None  # n >= 2f+1
  assert g["krum"].check(gradients=rows, f=2, m=7) is None and g["krum"].check(gradients=rows, f=2, m=8) is not None
  assert g["aksel"].check(gradients=rows, f=2, mode="bad") is not None
  assert g["median"].check(gradients=tuple(rows)) is not None and g["cge"].check(gradients=rows, f=99) is None
  assert g["median"].upper_bound(25, 5, 10) == 1 / math.sqrt(20)
  assert g["brute"].upper_bound(25, 5, 10) == 20 / (math.sqrt(8) * 5)
  n, f = 25, 5
  assert g["krum"].upper_bound(n, f, 1) == g["bulyan"].upper_bound(n, f, 1) == 1 / math.sqrt(2 * (n - f + f * (n + f * (n - f - 2) - 2) / (n - 2 * f - 2)))
  with pytest.raises(bz.UserException):
    g["krum"].checked(gradients=rows, f=9)
  assert g["average"].influence(rows[:8], rows[8:]) == 3 / 11
  for name in g:
    assert all(hasattr(g[name], member) for member in ("check", "checked", "unchecked", "upper_bound", "influence"))

def test_coordinate_host_rejects_bad_arguments_before_touching_the_gpu():
  """ The argument checks of `bz_coordinate_host` run before any CUDA call: no GPU needed. """
  import ctypes
  from byzantinemomentum_b200 import _lib
  lib = _lib.lib()
  buf = (ctypes.c_float * 256)()
  rows = (ctypes.c_void_p * 3)(*[ctypes.addressof(buf)] * 3)
  where = ctypes.addressof(buf)
  ok = dict(rule=2, n=3, f=1, d=64, pitch=64, chunks=4)
  def call(**kw):
    a = dict(ok, **kw)
    return lib.bz_coordinate_host(a["rule"], rows, a["n"], a["f"], a["d"], where, a.get("staging", where), a["pitch"], a.get("dev_out", where), a["chunks"], None, None, None)
  assert call(rule=7) == -1 and b"unknown rule" in lib.bz_last_error()
  assert call(f=2) == -1                      # n - 2f < 1
  assert call(pitch=32) == -1 and b"pitch" in lib.bz_last_error()
  assert call(chunks=0) == -1 and call(chunks=33) == -1
  assert call(staging=None) == -1
  assert call(n=65) == -2
  assert call(d=0) == 0                       # nothing to do
