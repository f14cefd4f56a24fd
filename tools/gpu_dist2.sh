#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "dist_driver" -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/t_dist1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/dist2.log 2>&1; echo "exit $?" >> gpurun_out/dist2.log
tail -n 20 gpurun_out/t_dist1.log gpurun_out/dist2.log
