set -x
mkdir -p gpurun_out
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py 2>&1 | tail -8 ) > gpurun_out/r2_p2p_n2b.log 2>&1
( timeout 300 python tools/k2_ab.py --cases 11:1310922,16:1310922,20:1310922,11:36489290,5:1310922,2:1310922 --no-alias --json gpurun_out/k2_ab_list.json 2>&1 | tail -8 ) > gpurun_out/r2_k2ab_list.log 2>&1
( timeout 600 python -m pytest tests/test_cuda_abi.py tests/test_cuda_golden.py tests/test_cuda_reuse.py -q -m gpu -x 2>&1 | tail -5 ) > gpurun_out/r2_t12.log 2>&1
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-sweep 2>gpurun_out/r2_bench_n2.err | tail -1 ) > gpurun_out/r2_bench_n2b.json
cat gpurun_out/r2_p2p_n2b.log gpurun_out/r2_k2ab_list.log; tail -4 gpurun_out/r2_t12.log; tail -3 gpurun_out/r2_bench_n2.err; python -c "
import json
l=json.loads(open('gpurun_out/r2_bench_n2b.json').read())
for r in l['sharded']['collective']['rules']: print(r)
print(l['e2e'])"
