# coding: utf-8
"""native.krum — Multi-Krum on the GPU (binds aggregators/krum.py:82-96)."""

def aggregate(gradients, f, m=None):
  from byzantinemomentum_b200 import engine
  if m is None:
    m = len(gradients) - f - 2
  return engine.krum(gradients, f, m)[0]
