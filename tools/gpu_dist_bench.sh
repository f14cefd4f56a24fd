#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist_check exit $?"
grep "dist stage1" gpurun_out/dist_check_$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
tail -n 5 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
for k in ('value','n_gpus','ms_per_step','e2e','gpu_launches','multi_gpu_stage1','clocks'): print(k, d[k])
PY
