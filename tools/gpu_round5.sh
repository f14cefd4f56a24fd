#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15 > gpurun_out/t5.log
tail -n 15 gpurun_out/t5.log
timeout 600 python tools/prof_classes.py 16384 64 f64 s1
timeout 300 python tools/prof_classes.py 3840 32 f64 s1
timeout 300 python tools/prof_classes.py 3840 32 f32 s1
