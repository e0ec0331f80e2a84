# coding: utf-8
"""Placement of pinned host buffers for the host-tensor path (`--device-gar cpu` callers, the
end-to-end leg of `bench.py`).

A host->device copy is a DMA read of host memory by the GPU.  On a two-socket host the pages of a
pinned buffer live on the NUMA node of the thread that allocated them (first touch); when that is
not the node the GPU hangs off, every copy crosses the inter-socket link and runs at a fraction of
the PCIe rate.  `gpu_local_cpus` binds the calling thread to the CPUs NVML reports as local to the
GPU while such buffers are allocated, and restores the previous affinity afterwards.
Everything here degrades to a no-op (no NVML, restricted cpuset, single node).
"""

import contextlib
import os

__all__ = ["gpu_local_cpus", "pin"]

def _nvml_handle(pynvml, device_index):
  try:
    import torch
    uuid = str(torch.cuda.get_device_properties(device_index).uuid)
    return pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
  except Exception:
    return pynvml.nvmlDeviceGetHandleByIndex(device_index)

@contextlib.contextmanager
def gpu_local_cpus(device_index=0):
  """ Context manager: the calling thread runs on the CPUs local to GPU `device_index`.
  Yields True when the affinity was narrowed, False when nothing could be done. """
  saved, applied = None, False
  try:
    import pynvml
    pynvml.nvmlInit()
    handle = _nvml_handle(pynvml, device_index)
    saved = os.sched_getaffinity(0)
    pynvml.nvmlDeviceSetCpuAffinity(handle)
    now = os.sched_getaffinity(0)
    applied = len(now) > 0 and now != saved
  except Exception:
    applied = False
  try:
    yield applied
  finally:
    if saved is not None:
      try:
        os.sched_setaffinity(0, saved)
      except OSError:
        pass

def pin(tensor, device_index=0):
  """ A pinned copy of a CPU tensor, allocated on the GPU-local NUMA node when possible. """
  with gpu_local_cpus(device_index):
    return tensor.pin_memory()
