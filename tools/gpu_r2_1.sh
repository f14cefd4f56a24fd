#!/bin/bash
# round 2, call 1: new large-size parity tests + baseline bench before any kernel change
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity_large.py -x -q -m gpu --durations=20 > gpurun_out/r2_t_large.log 2>&1
echo "large rc=$?" >> gpurun_out/r2_t_large.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err
echo "bench rc=$?" >> gpurun_out/r2_bench0.err
tail -5 gpurun_out/r2_t_large.log
