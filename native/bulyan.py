# coding: utf-8
"""native.bulyan — Bulyan over Multi-Krum on the GPU (binds aggregators/bulyan.py:86-100)."""

def aggregate(gradients, f, m=None):
  from byzantinemomentum_b200 import engine
  if m is None:
    m = len(gradients) - f - 2
  return engine.bulyan(gradients, f, m)[0]
