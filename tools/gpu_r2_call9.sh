#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=300 -k "trained or host_buffer" -s 2>&1 | grep -E "trained_|passed|failed|Error|error|assert" | head -40
timeout 400 python bench.py --steps 200 --warmup 5 --no-train > gpurun_out/r2_bench_zc.json 2> gpurun_out/r2_bench_zc.err
python - <<'P'
import json
j=json.load(open('gpurun_out/r2_bench_zc.json'))
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e'],'kernel_ms',j['roofline']['kernel_ms'], j['clocks'])
P
NERFB200_NO_ZERO_COPY=1 timeout 400 python bench.py --steps 200 --warmup 5 --no-train > gpurun_out/r2_bench_nozc.json 2> gpurun_out/r2_bench_nozc.err
python - <<'P'
import json
j=json.load(open('gpurun_out/r2_bench_nozc.json'))
print('NOZC value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e'],'kernel_ms',j['roofline']['kernel_ms'])
P
