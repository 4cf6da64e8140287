#!/bin/bash
# gpurun --gpus 8 -- 'bash tools/gpu_r2_n8.sh' : the scaling point the driver measures last, kept short (8x charge)
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_8gpu.txt 2>&1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_scale_8box_8gpu.json 2> gpurun_out/r2_scale_8box_8gpu.err
python - <<'P'
import json
try:
    j=json.load(open('gpurun_out/r2_scale_8box_8gpu.json'))
    print('N=8 value %.4g ms %.4f e2e %.4g kernel_ms %.4f train %s img %s' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['kernel_ms'], (j.get('train') or {}).get('ms_per_step'), (j.get('image_800') or {}).get('ms')))
    print(j['config']['timed_region'], '|', j['collective'], '|', j['clocks'])
except Exception as e: print('parse failed', e)
P
grep -E "NVLS|nranks|capture|eager" gpurun_out/r2_scale_8box_8gpu.err | sort | uniq -c | sort -rn | head -12
tail -3 gpurun_out/r2_scale_8box_8gpu.err
