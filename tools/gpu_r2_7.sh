#!/bin/bash
# round 2, call 7 (2 GPUs): multi-GPU parity (band gather, distributed svdvals, tcgen05 update) + shifted QR + panel timing
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_smi2.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu -s > gpurun_out/r2_t_dist.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_dist.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "shifted_qr or batched_svdvals" > gpurun_out/r2_t_sqr.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_sqr.log
export SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so
timeout 600 python tools/panel_blk_timing.py 3840 32 f64 4096 64 f64 8192 64 f32 > gpurun_out/r2_blk_timing4.log 2>&1
unset SVDB200_LIB
for lanes in 2 3 4; do
  SVDB200_LANES=$lanes timeout 600 python bench.py --steps 5 --warmup 3 --no-big --no-cpu-baseline > gpurun_out/r2_bench_l$lanes.json 2> gpurun_out/r2_bench_l$lanes.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --dist-n 32768 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "rc=$?" >> gpurun_out/r2_bench_n2.err
tail -12 gpurun_out/r2_t_dist.log; tail -5 gpurun_out/r2_t_sqr.log; cat gpurun_out/r2_blk_timing4.log; tail -2 gpurun_out/r2_bench_n2.err
