#!/bin/bash
# round 2, call 27: final build sanity -- Cholesky-QR panel tests, list pipeline, single-rank distributed driver, smoke()
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_dist.py -q -m gpu -k "chol or bidiagonalize_many or dist_driver or single_rank or svdvals_chain" > gpurun_out/r2_t_final.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_final.log
tail -4 gpurun_out/r2_t_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
