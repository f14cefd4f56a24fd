#!/bin/bash
# round 2, call 29: list pipeline, odd chains large-to-small (SVDB200_LIST_ZIGZAG=1) against ascending in every chain
mkdir -p gpurun_out
for V in "SVDB200_LIST_ZIGZAG=1" "SVDB200_LIST_ZIGZAG=0"; do
  env $V timeout 300 python bench.py --no-big --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/r2_bench_zz.json 2> gpurun_out/r2_bench_zz.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_zz.json').read().strip().splitlines()[-1])
    print('$V: value', round(d['value'], 1), 'ms_per_step', round(d['ms_per_step'], 1), 'e2e', round(d['e2e']['value'], 1))
except Exception as ex:
    print('$V: parse failed', ex)
PY
done
