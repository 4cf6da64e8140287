#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python tools/bwd_debug.py 64 > gpurun_out/r2_bwd_debug4.log 2>&1
grep -E "GLOBAL|sigma.weight|sigma.bias|xyz_encoding_[1-8].0.bias|Error|error" gpurun_out/r2_bwd_debug4.log | tail -24
timeout 600 python -m pytest tests -m gpu -q -k "training or gradients or deterministic or upstream or adam or host_buffer" 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches4.csv python tools/prof_train.py 1024 2 plain > gpurun_out/r2_train_ncu4.log 2>&1
timeout 200 python tools/prof_train.py 1024 50 time 2>&1 | head -3
