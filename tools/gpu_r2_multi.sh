#!/bin/bash
# multi-GPU validation: gpurun --gpus N -- 'bash tools/gpu_r2_multi.sh N'
set -x
N=${1:-2}
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_check.py > gpurun_out/r2_nccl_check.log 2>&1
tail -4 gpurun_out/r2_nccl_check.log
timeout 600 python -m pytest tests -m gpu -q -k nccl 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 1200 gpurun_out/r2_bench_${N}gpu.json; tail -5 gpurun_out/r2_bench_${N}gpu.err
