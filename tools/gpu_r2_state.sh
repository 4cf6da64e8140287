#!/bin/bash
# State-of-the-repo GPU run: tests, bench (both arms), launch lists of the bench and of the training step.
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -5
timeout 400 python bench.py --steps 200 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
tail -c 3000 gpurun_out/r2_bench_1gpu.json; tail -3 gpurun_out/r2_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
tail -c 600 gpurun_out/r2_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-train > gpurun_out/r2_ncu_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train.csv python tools/prof_train.py 1024 2 plain > gpurun_out/r2_ncu_train.log 2>&1
timeout 200 python tools/prof_train.py 1024 50 time 2>&1 | head -6
