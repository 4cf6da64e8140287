#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -s 2>&1 | grep -E "grad_trained|passed|failed|Error|FAILED" | head -20
timeout 400 python bench.py --steps 200 --warmup 5 > gpurun_out/r2c_bench_1gpu.json 2> gpurun_out/r2c_bench_1gpu.err
python - <<'P'
import json
j=json.load(open('gpurun_out/r2c_bench_1gpu.json'))
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'],'kernel_ms',j['roofline']['kernel_ms'], 'train', j['train'].get('ms_per_step'), j['train'].get('error'), j['clocks'])
print('parity', j['parity']); print('parity_trained', j['parity_trained_weights'])
P
tail -3 gpurun_out/r2c_bench_1gpu.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_train.csv python tools/prof_train.py 1024 2 plain > gpurun_out/r2c_ncu_train.log 2>&1
for v in nostore nostage; do
  NERFB200_LIB=nerf_pl_b200/variants/lib_$v.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_train_$v.csv python tools/prof_train.py 1024 2 plain > gpurun_out/r2c_ncu_train_$v.log 2>&1
done
python - <<'P'
import csv
for tag in ('', '_nostore', '_nostage'):
    f='gpurun_out/r2c_launches_train%s.csv'%tag
    try:
        lines=[l for l in open(f) if not l.startswith('==')]
    except Exception as e:
        print(tag, e); continue
    seq=[]
    for row in csv.DictReader(lines):
        if row.get('Metric Name')!='gpu__time_duration.sum': continue
        v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
        v = v/1000 if u=='ns' else v*1000 if u=='ms' else v
        seq.append((row['Kernel Name'][:48],v))
    idx=[i for i,(n,v) in enumerate(seq) if 'pack_weights' in n]
    last=seq[idx[-1]:]
    print(tag or 'base', ' | '.join(f"{n.split('::')[-1][:18]} {v:.0f}" for n,v in last if v>15), 'sum %.0f'%sum(v for n,v in last))
P
