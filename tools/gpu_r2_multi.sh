#!/bin/bash
# multi-GPU validation: gpurun --gpus N -- 'bash tools/gpu_r2_multi.sh N'
set -x
N=${1:-2}
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_check.py > gpurun_out/r2_nccl_check_${N}gpu.log 2>&1
tail -4 gpurun_out/r2_nccl_check_${N}gpu.log
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -4; fi
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n = 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_scale_${N}box_1gpu.json 2> gpurun_out/r2_scale_${N}box_1gpu.err
    else
      NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_scale_${N}box_${n}gpu.json 2> gpurun_out/r2_scale_${N}box_${n}gpu.err
    fi
    python - <<P
import json
try:
    j=json.load(open('gpurun_out/r2_scale_${N}box_${n}gpu.json'))
    print('N=$n value %.4g ms %.4f e2e %.4g train %s img %s' % (j['value'], j['ms_per_step'], j['e2e']['value'], (j.get('train') or {}).get('ms_per_step'), (j.get('image_800') or {}).get('ms')))
except Exception as e: print('N=$n parse failed', e)
P
    grep -E "NVLS|nranks|Using network|via P2P|NCCL version" gpurun_out/r2_scale_${N}box_${n}gpu.err | sort | uniq -c | head -8
  fi
done
