#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -s 2>&1 | grep -E "grad_trained|trained_|passed|failed|Error|rel |assert|FAILED" | head -40
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_train.csv python tools/prof_train.py 1024 2 plain > gpurun_out/r2b_ncu_train.log 2>&1
timeout 200 python tools/prof_train.py 1024 50 time 2>&1 | head -6
timeout 400 python bench.py --steps 200 --warmup 5 > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err
python - <<'P'
import json
j=json.load(open('gpurun_out/r2b_bench_1gpu.json'))
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'],'kernel_ms',j['roofline']['kernel_ms'], 'train', j['train'].get('ms_per_step'), j['train'].get('error'), j['clocks'])
P
tail -3 gpurun_out/r2b_bench_1gpu.err
