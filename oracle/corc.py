# coding: utf-8
"""ctypes binding of the C oracle (`oracle/c/libbyzoracle.so`, built by `__graft_entry__.build()`).
Test / bench infrastructure only (see `oracle/byzoracle.py`)."""

import ctypes
import pathlib

import numpy as np

_PATH = pathlib.Path(__file__).resolve().parent / "c" / "libbyzoracle.so"
_lib = None

def available():
  return _PATH.exists()

def lib():
  global _lib
  if _lib is None:
    handle = ctypes.CDLL(str(_PATH))
    rows_t = ctypes.POINTER(ctypes.c_void_p)
    i, i64, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p
    handle.orc_average.argtypes = [rows_t, i, i64, vp]
    handle.orc_median.argtypes = [rows_t, i, i64, vp]
    handle.orc_trmean.argtypes = [rows_t, i, i, i64, vp]
    handle.orc_closest.argtypes = [rows_t, i, i, i, i64, vp]
    handle.orc_pairdist.argtypes = [rows_t, i, i64, i, vp]
    handle.orc_rowdist_sq.argtypes = [rows_t, i, vp, i64, vp]
    handle.orc_average_selected.argtypes = [rows_t, vp, i, i, ctypes.c_float, i64, vp]
    for name in ("orc_average", "orc_median", "orc_trmean", "orc_closest", "orc_pairdist", "orc_rowdist_sq", "orc_average_selected"):
      getattr(handle, name).restype = None
    _lib = handle
  return _lib

def _rows(gradients):
  arrs = [np.ascontiguousarray(g, dtype=np.float32).reshape(-1) for g in gradients]
  ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
  return arrs, ptrs

def average(gradients):
  arrs, ptrs = _rows(gradients)
  out = np.empty(arrs[0].shape[0], dtype=np.float32)
  lib().orc_average(ptrs, len(arrs), out.shape[0], out.ctypes.data)
  return out

def median(gradients):
  arrs, ptrs = _rows(gradients)
  out = np.empty(arrs[0].shape[0], dtype=np.float32)
  lib().orc_median(ptrs, len(arrs), out.shape[0], out.ctypes.data)
  return out

def trmean(gradients, f):
  arrs, ptrs = _rows(gradients)
  out = np.empty(arrs[0].shape[0], dtype=np.float32)
  lib().orc_trmean(ptrs, len(arrs), f, out.shape[0], out.ctypes.data)
  return out

def phocas(gradients, f):
  arrs, ptrs = _rows(gradients)
  out = np.empty(arrs[0].shape[0], dtype=np.float32)
  lib().orc_closest(ptrs, len(arrs), f, 0, out.shape[0], out.ctypes.data)
  return out

def meamed(gradients, f):
  arrs, ptrs = _rows(gradients)
  out = np.empty(arrs[0].shape[0], dtype=np.float32)
  lib().orc_closest(ptrs, len(arrs), f, 1, out.shape[0], out.ctypes.data)
  return out

def pairwise_distances(gradients, map_nonfinite=True):
  arrs, ptrs = _rows(gradients)
  n = len(arrs)
  D = np.zeros((n, n), dtype=np.float64)
  lib().orc_pairdist(ptrs, n, arrs[0].shape[0], 1 if map_nonfinite else 0, D.ctypes.data)
  return D

def rowdist_sq(gradients, center=None):
  arrs, ptrs = _rows(gradients)
  out = np.empty(len(arrs), dtype=np.float64)
  c = None if center is None else np.ascontiguousarray(center, dtype=np.float32)
  lib().orc_rowdist_sq(ptrs, len(arrs), None if c is None else c.ctypes.data, arrs[0].shape[0], out.ctypes.data)
  return out

def average_selected(gradients, selection, zero_init=True, divisor=None):
  arrs, ptrs = _rows(gradients)
  sel = np.ascontiguousarray(selection, dtype=np.int32)
  out = np.empty(arrs[0].shape[0], dtype=np.float32)
  lib().orc_average_selected(ptrs, sel.ctypes.data, len(sel), 1 if zero_init else 0,
                             float(len(sel) if divisor is None else divisor), out.shape[0], out.ctypes.data)
  return out
