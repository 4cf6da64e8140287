#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
NERFB200_LIB=nerf_pl_b200/variants/lib_direct.so timeout 600 python -m pytest tests -m gpu -q --timeout=300 -k "training or gradients or deterministic or trained or upstream or adam" 2>&1 | tail -3
for v in direct; do
  NERFB200_LIB=nerf_pl_b200/variants/lib_$v.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_train_$v.csv python tools/prof_train.py 1024 2 plain > gpurun_out/r2d_ncu_train_$v.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_train.csv python tools/prof_train.py 1024 2 plain > gpurun_out/r2d_ncu_train.log 2>&1
python - <<'P'
import csv
for tag in ('', '_direct'):
    f='gpurun_out/r2d_launches_train%s.csv'%tag
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
NERFB200_LIB=nerf_pl_b200/variants/lib_direct.so timeout 200 python tools/prof_train.py 1024 50 time 2>&1 | head -2
timeout 200 python tools/prof_train.py 1024 50 time 2>&1 | head -2
