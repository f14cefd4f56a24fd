#!/bin/bash
# round 2, call 21: new Cholesky-QR panel tests; bench sweep with 2 / 3 / 4 chains in the list pipeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_large.py -x -q -m gpu -k "chol or vs_oracle_larger" > gpurun_out/r2_t_chol6.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_chol6.log
tail -4 gpurun_out/r2_t_chol6.log
for L in 2 3 4; do
  SVDB200_LANES=$L timeout 600 python bench.py --no-big --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2_bench_lanes$L.json 2> gpurun_out/r2_bench_lanes$L.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_lanes$L.json').read().strip().splitlines()[-1])
    print('lanes $L: value', round(d['value'], 1), 'ms_per_step', round(d['ms_per_step'], 1), 'e2e', round(d['e2e']['value'], 1))
except Exception as ex:
    print('lanes $L: parse failed', ex)
PY
done
