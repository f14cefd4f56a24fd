#!/bin/bash
# round 2, call 17 (2 GPUs): distributed LQ panel (local Gram + all-reduce) parity on 2 ranks, fallback path, block-cyclic timing at n=16384 / 32768
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/dist_check.py small > gpurun_out/r2_dist_check_2s.log 2>&1
echo "rc=$?" >> gpurun_out/r2_dist_check_2s.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tests/dist_check.py large > gpurun_out/r2_dist_check_2l.log 2>&1
echo "rc=$?" >> gpurun_out/r2_dist_check_2l.log
grep -E "^dist|rc=|Error|error" gpurun_out/r2_dist_check_2s.log gpurun_out/r2_dist_check_2l.log | cut -c1-260
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 tools/dist_only.py 32768 > gpurun_out/r2_dist_only_2.log 2>&1
echo "rc=$?" >> gpurun_out/r2_dist_only_2.log
tail -22 gpurun_out/r2_dist_only_2.log
