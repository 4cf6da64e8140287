#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 200 python tools/prof_train.py 1024 50 time > gpurun_out/r2_train_time.log 2>&1
cat gpurun_out/r2_train_time.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches.csv python tools/prof_train.py 1024 2 plain > gpurun_out/r2_train_ncu.log 2>&1
tail -3 gpurun_out/r2_train_ncu.log
AB_ROUNDS=3 timeout 400 python tools/ab_speed.py nerf_pl_b200/libnerf_pl_b200.so nerf_pl_b200/variants/lib_split.so > gpurun_out/r2_ab2.log 2>&1
cat gpurun_out/r2_ab2.log
NERFB200_LIB=nerf_pl_b200/variants/lib_split.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or tile_pipeline or full_image" 2>&1 | tail -3
