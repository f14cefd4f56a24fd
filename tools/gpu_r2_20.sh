#!/bin/bash
# round 2, call 20 (8 GPUs): block-cyclic stage 1 at n=65536 with the distributed QR panel (raw-row broadcast), early broadcast on / off
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/dist_only.py 65536 > gpurun_out/r2_dist_only_8b.log 2>&1
echo "rc=$?" >> gpurun_out/r2_dist_only_8b.log
grep -v "^\*\|OMP_NUM" gpurun_out/r2_dist_only_8b.log | tail -20
SVDB200_DIST_EARLY_BCAST=0 SKIP_N1=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/dist_only.py 65536 > gpurun_out/r2_dist_only_8b_noearly.log 2>&1
echo "rc=$?" >> gpurun_out/r2_dist_only_8b_noearly.log
grep -E '"ms"|rc=|error' gpurun_out/r2_dist_only_8b_noearly.log
