#!/bin/bash
# round 2, call 24: full GPU suite (no -x), smoke()
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/r2_t_all4.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_all4.log
tail -16 gpurun_out/r2_t_all4.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1
tail -n 2 gpurun_out/r2_smoke.log
