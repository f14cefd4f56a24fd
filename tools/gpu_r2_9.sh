#!/bin/bash
# round 2, call 9 (8 GPUs): multi-GPU parity at 8 ranks, then the bench line at N=8 (configs[3] at n=65536 with in-run
# invariants and the one-GPU time of the same run, configs[4] sharded by matrix)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_smi8.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/dist_check.py large > gpurun_out/r2_dist_check_8.log 2>&1
echo "rc=$?" >> gpurun_out/r2_dist_check_8.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "rc=$?" >> gpurun_out/r2_bench_n8.err
grep -E "^dist|rc=" gpurun_out/r2_dist_check_8.log; tail -2 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
    print(json.dumps(d['config']['north_star'].get('config3_block_cyclic_stage1'), indent=1))
    print(json.dumps(d['config']['north_star'].get('config4_batched_8192x256'), indent=1))
    print('value', d['value'])
except Exception as ex:
    print('parse failed', ex)
PY
