# coding: utf-8
"""`native` — the compiled-backend slot the reference reserves.

`aggregators/{median,krum,bulyan,brute}.py` of LPD-EPFL/ByzantineMomentum do `import native`
and, when `native.<gar>` exists, register `native-<gar>` calling
`native.<gar>.aggregate(gradients[, f[, m]])` positionally (median.py:49, krum.py:96,
bulyan.py:100, brute.py:91).  Putting this repository's root on `sys.path` therefore turns the
unmodified reference into a client of the CUDA kernels: `--gar native-krum`.

Importing this package must raise nothing but ImportError (any other exception makes the
reference drop the whole GAR module, tools/__init__.py:295-305), so no CUDA work happens at
import: the library loads on the first `aggregate` call.
"""

try:
  from . import median, krum, bulyan, brute
except ImportError:
  raise
except Exception as err:  # pragma: no cover - keep the reference's loader alive
  raise ImportError(f"native backend unavailable: {err}") from err

__all__ = ["median", "krum", "bulyan", "brute"]
