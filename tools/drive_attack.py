#!/usr/bin/env python3
# coding: utf-8
"""Run the UNMODIFIED reference `attack.py` offline, with the CUDA rules registered.

The reference downloads its datasets (torchvision); here `experiments.make_datasets` — looked up
at call time by attack.py:530 — is replaced by a generator of synthetic batches of the right
shape (`experiments/dataset.py:187-190` accepts a generator), the rules of this repository are
registered through the reference's own `aggregators.register` under `b200-<name>`
(`plugin.install`), optionally the study metrics are swapped (`plugin.install_tools`), and
`attack.py` is executed with `runpy` exactly as `python3 attack.py <args>` would.

    python tools/drive_attack.py --reference /path/to/ByzantineMomentum --shape mnist -- \\
        --gar b200-krum --nb-workers 11 --nb-decl-byz 3 --nb-real-byz 3 --attack empire \\
        --attack-args factor:1.1 --model simples-full --nb-steps 3 --device cuda:0

Nothing in the reference is modified; no file of it is copied.
"""

import argparse
import pathlib
import runpy
import sys

SHAPES = {"mnist": ((1, 28, 28), 10), "cifar10": ((3, 32, 32), 10), "cifar100": ((3, 32, 32), 100)}

def main():
  parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
  parser.add_argument("--reference", required=True, help="root of the ByzantineMomentum checkout")
  parser.add_argument("--shape", default="mnist", choices=sorted(SHAPES), help="shape of the synthetic samples")
  parser.add_argument("--override", action="store_true", help="replace the stock rules instead of adding b200-<name>")
  parser.add_argument("--install-tools", action="store_true", help="also swap tools.compute_avg_dev_max for CUDA samples")
  parser.add_argument("rest", nargs=argparse.REMAINDER, help="arguments of attack.py (after --)")
  args = parser.parse_args()
  rest = args.rest[1:] if args.rest[:1] == ["--"] else args.rest
  ref = pathlib.Path(args.reference).resolve()
  root = pathlib.Path(__file__).resolve().parent.parent
  sys.path.insert(0, str(root))
  sys.path.insert(0, str(ref))

  import torch
  import aggregators
  import experiments
  import tools
  import byzantinemomentum_b200 as bz

  sample_shape, classes = SHAPES[args.shape]
  def make_datasets(dataset, train_batch, test_batch, **kwargs):
    def batches(size, seed):
      gen = torch.Generator().manual_seed(seed)
      while True:
        yield torch.randn((size,) + sample_shape, generator=gen), torch.randint(classes, (size,), generator=gen)
    return experiments.Dataset(batches(train_batch or 32, 1), name="synthetic-train"), \
           experiments.Dataset(batches(test_batch or 32, 2), name="synthetic-test")
  experiments.make_datasets = make_datasets

  names = bz.plugin.install(aggregators, override=args.override)
  if args.install_tools:
    bz.plugin.install_tools(tools)
  print(f"registered: {', '.join(names)}", flush=True)
  sys.argv = [str(ref / "attack.py")] + rest
  runpy.run_path(str(ref / "attack.py"), run_name="__main__")

if __name__ == "__main__":
  main()
