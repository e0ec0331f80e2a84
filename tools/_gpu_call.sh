set -x
mkdir -p gpurun_out
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py 2>&1 | tail -8 ) > gpurun_out/r2_p2p_n8.log 2>&1
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 100 --warmup 5 2>gpurun_out/r2_bench_n8.err | tail -1 ) > gpurun_out/r2_bench_n8.json
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 2>>gpurun_out/r2_bench_n8.err | tail -1 ) > gpurun_out/r2_bench_n8_ref.json
cat gpurun_out/r2_p2p_n8.log; tail -3 gpurun_out/r2_bench_n8.err; python -c "
import json
l=json.loads(open('gpurun_out/r2_bench_n8.json').read())
print(l['value'], l['ms_per_step'], l['n_gpus'], l['clocks'])
print(json.dumps(l['sharded'],indent=1)[:5000]); print(l['e2e'])
print(open('gpurun_out/r2_bench_n8_ref.json').read()[:300])"
