#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "stage2 or bidiag_qr or chain" -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/t2.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -n 15 gpurun_out/t2.log gpurun_out/bench.err gpurun_out/bench_ref.err; cat gpurun_out/bench.json gpurun_out/bench_ref.json
