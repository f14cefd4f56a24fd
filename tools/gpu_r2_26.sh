#!/bin/bash
# round 2, call 26 (8 GPUs): multi-GPU parity at 8 ranks with the distributed panels (elementwise vs one GPU), configs[3] at n = 65536
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_smi8c.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/dist_check.py large > gpurun_out/r2_dist_check_8b.log 2>&1
echo "rc=$?" >> gpurun_out/r2_dist_check_8b.log
grep -E "^dist|rc=" gpurun_out/r2_dist_check_8b.log | cut -c1-260
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/dist_only.py 65536 > gpurun_out/r2_dist_only_8c.log 2>&1
echo "rc=$?" >> gpurun_out/r2_dist_only_8c.log
grep -v "^\*\|OMP_NUM" gpurun_out/r2_dist_only_8c.log | tail -19
