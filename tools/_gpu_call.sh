set -x
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_cuda_sharded_phases.py -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2_t10.log 2>&1
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py 2>&1 | tail -12 ) > gpurun_out/r2_p2p_n2.log 2>&1
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 2>gpurun_out/r2_bench_n2.err | tail -1 ) > gpurun_out/r2_bench_n2.json
( timeout 300 python tools/abk2.py 2>&1 | tail -8 ) > gpurun_out/r2_rules5.log 2>&1
tail -5 gpurun_out/r2_t10.log; cat gpurun_out/r2_p2p_n2.log; tail -3 gpurun_out/r2_bench_n2.err; python -c "
import sys,json
l=json.loads(open('gpurun_out/r2_bench_n2.json').read())
print(json.dumps(l.get('sharded'),indent=1)[:4000]); print(l['value'], l['ms_per_step'], l['e2e'])"
cat gpurun_out/r2_rules5.log
