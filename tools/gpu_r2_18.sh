#!/bin/bash
# round 2, call 18 (8 GPUs): BASELINE configs[3] shape only -- block-cyclic stage 1 at n=65536 on 8 ranks (+ the one-GPU time of the
# same run), with the look-ahead SM reservation on and off
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_smi8b.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/dist_only.py 65536 > gpurun_out/r2_dist_only_8.log 2>&1
echo "rc=$?" >> gpurun_out/r2_dist_only_8.log
grep -v "^\*\|OMP_NUM" gpurun_out/r2_dist_only_8.log | tail -20
SVDB200_RESERVE_SMS=0 SKIP_N1=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/dist_only.py 65536 > gpurun_out/r2_dist_only_8_noreserve.log 2>&1
echo "rc=$?" >> gpurun_out/r2_dist_only_8_noreserve.log
grep -E '"ms"|rc=' gpurun_out/r2_dist_only_8_noreserve.log
