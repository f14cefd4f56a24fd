#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/t6.log
tail -n 6 gpurun_out/t6.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench exit $?"
tail -n 5 gpurun_out/bench_r1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1.json 2>gpurun_out/bench_ref_r1.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1.json'))
for k in ('value','ms_per_step','e2e','gpu_launches','roofline','north_star_shape','cpu_baseline','clocks'): print(k, d[k])
for k,v in d['kernel_classes_f64'].items(): print(k, v)
print(open('gpurun_out/bench_ref_r1.json').read()[:400])
PY
