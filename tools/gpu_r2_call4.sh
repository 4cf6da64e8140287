#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python tools/bwd_debug.py 64 > gpurun_out/r2_bwd_debug2.log 2>&1
grep -E "dpre|dd |GLOBAL|grad fine.xyz_encoding_1|Error|error" gpurun_out/r2_bwd_debug2.log | tail -30
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest4.log 2>&1
tail -8 gpurun_out/r2_pytest4.log
timeout 200 python tools/prof_train.py 1024 50 time > gpurun_out/r2_train_time2.log 2>&1
cat gpurun_out/r2_train_time2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches2.csv python tools/prof_train.py 1024 2 plain > gpurun_out/r2_train_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_wgrad python tools/prof_train.py 1024 1 plain > gpurun_out/r2_ncu_wgrad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_bwd_kernel -s 7 -c 1 -f -o gpurun_out/prof_r2_chain python tools/prof_train.py 1024 1 plain > gpurun_out/r2_ncu_chain.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_rays_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_fwdsave python tools/prof_train.py 1024 1 plain > gpurun_out/r2_ncu_fwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
