# coding: utf-8
"""`hostmem.gpu_local_cpus`: narrows the thread's CPU affinity to the GPU-local CPUs while pinned
buffers are allocated, restores it afterwards, and is a harmless no-op where NVML is absent."""

import os
import sys
import types

import pytest

from byzantinemomentum_b200 import hostmem

def test_noop_without_nvml_and_affinity_restored():
  before = os.sched_getaffinity(0)
  with hostmem.gpu_local_cpus(0) as applied:
    assert applied in (True, False)
  assert os.sched_getaffinity(0) == before

def test_affinity_is_narrowed_then_restored(monkeypatch):
  before = os.sched_getaffinity(0)
  if len(before) < 2:
    pytest.skip("one CPU only")
  keep = {sorted(before)[0]}
  fake = types.SimpleNamespace(
    nvmlInit=lambda: None,
    nvmlDeviceGetHandleByUUID=lambda uuid: (_ for _ in ()).throw(RuntimeError("no such GPU")),
    nvmlDeviceGetHandleByIndex=lambda index: ("handle", index),
    nvmlDeviceSetCpuAffinity=lambda handle: os.sched_setaffinity(0, keep))
  monkeypatch.setitem(sys.modules, "pynvml", fake)
  with hostmem.gpu_local_cpus(0) as applied:
    assert applied is True
    assert os.sched_getaffinity(0) == keep
  assert os.sched_getaffinity(0) == before

def test_failure_inside_nvml_leaves_the_affinity_alone(monkeypatch):
  before = os.sched_getaffinity(0)
  def boom(*args):
    raise RuntimeError("NVML_ERROR_NOT_SUPPORTED")
  fake = types.SimpleNamespace(nvmlInit=lambda: None, nvmlDeviceGetHandleByUUID=boom, nvmlDeviceGetHandleByIndex=lambda i: i,
                               nvmlDeviceSetCpuAffinity=boom)
  monkeypatch.setitem(sys.modules, "pynvml", fake)
  with hostmem.gpu_local_cpus(0) as applied:
    assert applied is False
  assert os.sched_getaffinity(0) == before
