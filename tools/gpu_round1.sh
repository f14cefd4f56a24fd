#!/bin/bash
# first GPU round: parity tests by group (each under its own timeout so a hang cannot starve the rest)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 420 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$1" -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/t_$name.log; echo "exit $?" >> gpurun_out/t_$name.log; }
run stage2 "stage2"
run gemm "trailing_update"
run panel "panel_order"
run tile "tile_order"
run qr "bidiag_qr"
run chain "chain or error"
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
timeout 120 python - > gpurun_out/peaks.log 2>&1 <<'PY'
import numpy as np
from svdsolver_b200 import capi
with capi.Handle(256, 32, np.float64) as h:
    for k, name in enumerate(("DFMA", "DMMA_f64_m8n8k4", "FFMA", "TF32_mma_sync")):
        print(name, round(h.probe_peak(k), 2), "TFLOP/s")
PY
tail -n 30 gpurun_out/t_*.log gpurun_out/smoke.log gpurun_out/peaks.log
