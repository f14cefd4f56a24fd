#!/bin/bash
# round 2, call 22: list pipeline (bench sweep) with / without the Cholesky-QR panels beside the stage-2 kernels
mkdir -p gpurun_out
for V in "SVDB200_PANEL_CHOL=0" "SVDB200_PIPE_CHOL=0" "SVDB200_PIPE_CHOL=1"; do
  env $V timeout 600 python bench.py --no-big --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2_bench_v.json 2> gpurun_out/r2_bench_v.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_v.json').read().strip().splitlines()[-1])
    print('$V: value', round(d['value'], 1), 'ms_per_step', round(d['ms_per_step'], 1), 'e2e', round(d['e2e']['value'], 1))
except Exception as ex:
    print('$V: parse failed', ex)
PY
done
