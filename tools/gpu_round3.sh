#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/t3.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?" >> gpurun_out/bench.err
tail -n 15 gpurun_out/t3.log gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
for k in ('value','ms_per_step','e2e','gpu_launches','roofline','cpu_baseline','clocks'): print(k, d[k])
for k,v in d['kernel_classes_f64'].items(): print(k, v)
for r in d['detail']: print(r)
PY
