# coding: utf-8
"""The `native.<gar>.aggregate` hooks are called POSITIONALLY by the reference (median.py:49,
krum.py:96, bulyan.py:100, brute.py:91); they must forward to the engine with the reference's
defaults and return the aggregated vector only."""

import importlib

import pytest

engine = importlib.import_module("byzantinemomentum_b200.engine")

@pytest.fixture
def calls(monkeypatch):
  seen = []
  monkeypatch.setattr(engine, "median", lambda gradients: seen.append(("median", len(gradients))) or "out")
  monkeypatch.setattr(engine, "krum", lambda gradients, f, m: seen.append(("krum", f, m)) or ("out", "order"))
  monkeypatch.setattr(engine, "bulyan", lambda gradients, f, m: seen.append(("bulyan", f, m)) or ("out", "order"))
  monkeypatch.setattr(engine, "brute", lambda gradients, f: seen.append(("brute", f)) or ("out", "sel"))
  return seen

def test_hooks_forward_positionally_with_the_reference_defaults(calls):
  import native
  rows = list(range(11))
  assert native.median.aggregate(rows) == "out"
  assert native.krum.aggregate(rows, 3, None) == "out"        # m = None -> n - f - 2 (krum.py:76-77)
  assert native.krum.aggregate(rows, 3, 2) == "out"
  assert native.bulyan.aggregate(rows, 2, None) == "out"
  assert native.brute.aggregate(rows, 4) == "out"
  assert calls == [("median", 11), ("krum", 3, 6), ("krum", 3, 2), ("bulyan", 2, 7), ("brute", 4)]

def test_importing_native_does_not_touch_cuda_or_the_library():
  """ The reference drops a whole GAR module if importing `native` raises anything but
  ImportError (tools/__init__.py:295-305): the import must stay free of side effects. """
  import sys
  for name in [k for k in sys.modules if k == "native" or k.startswith("native.")]:
    del sys.modules[name]
  lib_module = importlib.import_module("byzantinemomentum_b200._lib")
  before = lib_module._lib
  lib_module._lib = None
  try:
    import native
    assert sorted(n for n in dir(native) if not n.startswith("_")) == ["brute", "bulyan", "krum", "median"]
    assert lib_module._lib is None                              # nothing was loaded
  finally:
    lib_module._lib = before
