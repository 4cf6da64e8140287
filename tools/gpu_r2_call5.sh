#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python tools/bwd_debug.py 64 > gpurun_out/r2_bwd_debug3.log 2>&1
grep -E "dpre[18]|dd |GLOBAL|grad fine.xyz_encoding_1|Error|error" gpurun_out/r2_bwd_debug3.log | tail -16
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest5.log 2>&1
tail -8 gpurun_out/r2_pytest5.log
timeout 200 python tools/prof_train.py 1024 50 time > gpurun_out/r2_train_time3.log 2>&1
head -3 gpurun_out/r2_train_time3.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches3.csv python tools/prof_train.py 1024 2 plain > gpurun_out/r2_train_ncu3.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
tail -c 1500 gpurun_out/r2_bench5.json; tail -3 gpurun_out/r2_bench5.err
