#!/bin/bash
# round 2, call 4: stage-2 fast kernel (band 32): bit-exactness vs oracle at depth + timing vs the previous kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_large.py -x -q -m gpu -k "deep_pipeline or many_pipeline" > gpurun_out/r2_t_s2fast.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_s2fast.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stage2" >> gpurun_out/r2_t_s2fast.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_s2fast.log
for dt in f64 f32; do
  for n in 1920 3840; do
    python tools/stage2_only.py $n 32 $dt >> gpurun_out/r2_s2_timing.log 2>&1
    SVDB200_S2_FAST=0 python tools/stage2_only.py $n 32 $dt >> gpurun_out/r2_s2_timing.log 2>&1
  done
done
tail -6 gpurun_out/r2_t_s2fast.log; cat gpurun_out/r2_s2_timing.log
