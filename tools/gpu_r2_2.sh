#!/bin/bash
# round 2, call 2: blocked panel kernel -- parity + timing, then the GPU test suite and a bench
mkdir -p gpurun_out
timeout 900 python tools/panel_diag.py > gpurun_out/r2_panel_diag.log 2>&1
echo "diag rc=$?" >> gpurun_out/r2_panel_diag.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc05.py -x -q -m gpu > gpurun_out/r2_t_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2_t_parity.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo "bench rc=$?" >> gpurun_out/r2_bench1.err
tail -15 gpurun_out/r2_panel_diag.log; tail -4 gpurun_out/r2_t_parity.log
