#!/bin/bash
# round 2, call 30: pass 1 over all rows (no A1^T A1 product in the algebra kernel): parity, single-rank distributed driver, stage-1 time
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_dist.py -q -m gpu -k "chol or dist_driver or single_rank or vs_oracle_larger" > gpurun_out/r2_t_g.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_g.log
tail -3 gpurun_out/r2_t_g.log
REPS=3 timeout 100 python tools/stage1_only.py 16384 64 f32 2>&1 | grep "stage1 ms"
