#!/bin/bash
# round-2 GPU call 2: first runs of the hand-written backward + new bench
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python tools/bwd_debug.py 64 > gpurun_out/r2_bwd_debug.log 2>&1
tail -5 gpurun_out/r2_bwd_debug.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest2.log 2>&1
tail -15 gpurun_out/r2_pytest2.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
tail -c 3000 gpurun_out/r2_bench2.json; tail -5 gpurun_out/r2_bench2.err
