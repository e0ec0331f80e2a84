set -x
mkdir -p gpurun_out
( timeout 400 python tools/k2_ab.py --json gpurun_out/k2_ab_2.json 2>&1 | tail -40 ) > gpurun_out/r2_k2ab2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_ring -s 3 -c 1 -o gpurun_out/r2_k2ring_n25_b python tools/k2_ab.py --cases 25:1310922 --no-alias --only ring > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_ring -s 3 -c 1 -o gpurun_out/r2_k2ring_n51_b python tools/k2_ab.py --cases 51:1310922 --no-alias --only ring > gpurun_out/ncu2.log 2>&1
cat gpurun_out/r2_k2ab2.log
