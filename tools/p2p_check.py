#!/usr/bin/env python3
# coding: utf-8
"""Under torchrun (N >= 2): the peer-memory exchange must select exactly what the NCCL exchange
selects, rank by rank, and produce the same shard; prints per-step times of both."""
import datetime, os, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
from byzantinemomentum_b200 import sharded
import byzantinemomentum_b200 as bz

def main():
  rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
  bz.config.strict_status = False
  n, nb, f, d = 25, 5, 5, 1_310_922
  gen = torch.Generator(device=dev).manual_seed(100 + rank)
  honest = [torch.randn(d, device=dev, generator=gen) * (0.5 + i / 19) for i in range(n - nb)]
  byz = torch.stack(honest).mean(dim=0).mul(-1.1)
  rows = honest + [byz] * nb
  ok = True
  for gar in ("krum", "bulyan", "cge", "aksel"):
    a, sa = sharded.aggregate(gar, rows, f=f, return_selection=True)
    b, sb = sharded.aggregate_p2p(gar, rows, f=f, return_selection=True)
    same = torch.equal(a, b) and torch.equal(sa, sb)
    ok = ok and same
    def timeit(fn, K=50):
      for _ in range(5): fn()
      torch.cuda.synchronize(); dist.barrier()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(K): fn()
      e1.record(); torch.cuda.synchronize()
      return e0.elapsed_time(e1) / K * 1e3
    t_nccl = timeit(lambda: sharded.aggregate(gar, rows, f=f))
    t_p2p = timeit(lambda: sharded.aggregate_p2p(gar, rows, f=f))
    # the prepared calls: same shard, same selection, as the step-by-step ones
    plan_n = sharded.ShardedPlan(gar, rows, f=f, exchange="nccl")
    plan_p = sharded.ShardedPlan(gar, rows, f=f, exchange="p2p")
    same_plan = torch.equal(plan_n(), a) and torch.equal(plan_p(), a) and torch.equal(plan_n.selection, sa[:plan_n.selection.numel()]) \
                and torch.equal(plan_p.selection, sa[:plan_p.selection.numel()])
    same_plan = same_plan and torch.equal(plan_p(), a) and torch.equal(plan_p(), a)       # both slots of the symmetric buffer
    ok = ok and same_plan
    t_plan_n, t_plan_p = timeit(plan_n, 200), timeit(plan_p, 200)
    t_plan_f = float("nan")
    if gar in ("krum", "bulyan"):
      plan_f = sharded.ShardedPlan(gar, rows, f=f, exchange="fused")
      same_f = all(torch.equal(plan_f(), a) for _ in range(4)) and torch.equal(plan_f.selection, sa[:plan_f.selection.numel()]) and int(plan_f.status.item()) == 0
      same_plan = same_plan and same_f
      ok = ok and same_f
      t_plan_f = timeit(plan_f, 200)
    single = bz.Plan(gar, rows, f=f)
    t_single = timeit(single, 200)
    if rank == 0:
      print(f"{gar:7s} N={world}: identical={same} plans={same_plan}  aggregate: nccl {t_nccl:7.1f} p2p {t_p2p:7.1f} | ShardedPlan: nccl {t_plan_n:7.1f} p2p {t_plan_p:7.1f} fused {t_plan_f:7.1f} | single-GPU Plan {t_single:7.1f} us/step", flush=True)
  flag = torch.tensor([1 if ok else 0], device=dev)
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
  if rank == 0:
    print("ALL RANKS IDENTICAL" if int(flag.item()) == 1 else "MISMATCH", flush=True)
  dist.destroy_process_group()

if __name__ == "__main__":
  main()
