#!/bin/bash
# round 2, call 28: balanced chain schedule of the list pipeline: list tests, then the bench sweep
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -q -m gpu -k "bidiagonalize_many or many_pipeline" > gpurun_out/r2_t_many.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_many.log
tail -3 gpurun_out/r2_t_many.log
timeout 300 python bench.py --no-big --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2_bench_bal.json 2> gpurun_out/r2_bench_bal.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_bal.json').read().strip().splitlines()[-1])
    print('balanced: value', round(d['value'], 1), 'ms_per_step', round(d['ms_per_step'], 1), 'e2e', round(d['e2e']['value'], 1))
except Exception as ex:
    print('parse failed', ex)
PY
