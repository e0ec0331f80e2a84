# coding: utf-8
"""Registration of the CUDA rules inside the reference's own registry.

The reference resolves `--gar <name>` through `aggregators.gars` (attack.py:468).  Two
routes lead there without touching a reference file:

  1. `install(aggregators)` — calls the reference's `aggregators.register(name, unchecked,
     check, upper_bound, influence)` (aggregators/__init__.py:71-86) for every rule, under
     `<prefix><name>` (default prefix "b200-"), or replaces the stock entries in place with
     `override=True` (`register` refuses duplicates, :82-84, so the entry is rebuilt with the
     reference's `make_gar`).
  2. the top-level `native` package of this repository: the reference already looks for
     `native.{median,krum,bulyan,brute}.aggregate` and registers `native-<name>` when it finds
     them (median.py:80-87, krum.py:159-166, bulyan.py:137-144, brute.py:149-156).
"""

from .gars import gars

__all__ = ["install", "install_tools"]

def install(aggregators, prefix="b200-", override=False, names=None):
  """ Register the CUDA rules in the reference's `aggregators` module.
  Args:
    aggregators  The imported reference package (duck-typed: needs `gars`, `register`, `make_gar`)
    prefix       Name prefix of the new entries (ignored with override)
    override     Replace the stock rules under their own names instead
    names        Iterable of rule names to install (default: all)
  Returns:
    List of the registered names
  """
  done = []
  for name in (names or gars.keys()):
    rule = gars[name]
    if override:
      wrapped = aggregators.make_gar(rule.unchecked, rule.check, upper_bound=rule.upper_bound, influence=rule.influence)
      aggregators.gars[name] = wrapped
      setattr(aggregators, name, wrapped)
      done.append(name)
    else:
      target = prefix + name
      if target not in aggregators.gars:
        aggregators.register(target, rule.unchecked, rule.check, upper_bound=rule.upper_bound, influence=rule.influence)
      done.append(target)
  return done

def install_tools(tools):
  """ Replace `tools.compute_avg_dev_max` (tools/pytorch.py:97, used by attack.py:846-848 for the
  study metrics) with the CUDA version when the samples live on a GPU.  CPU samples keep the
  reference's own function: this is a device kernel, not a CPU reimplementation.
  Args:
    tools  The imported reference `tools` package
  Returns:
    The previous function
  """
  from . import engine, _lib
  stock = tools.compute_avg_dev_max
  def compute_avg_dev_max(samples):
    if len(samples) > 0 and samples[0].is_cuda and len(samples) <= _lib.MAX_N:
      return engine.compute_avg_dev_max(samples)
    return stock(samples)
  compute_avg_dev_max.__doc__ = stock.__doc__
  tools.compute_avg_dev_max = compute_avg_dev_max
  return stock
