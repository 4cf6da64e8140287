#!/bin/bash
# round-2 GPU call 1: MN-major descriptor probe, the two round-1 experiments (correctness + A/B), base tests
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
NERFB200_LIB=nerf_pl_b200/variants/lib_diag.so timeout 180 python tools/gpu_probe.py gemm_mn > gpurun_out/r2_gemm_mn.log 2>&1
tail -6 gpurun_out/r2_gemm_mn.log
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_base.log 2>&1
tail -3 gpurun_out/r2_pytest_base.log
for v in hoist pingpong; do
  NERFB200_LIB=nerf_pl_b200/variants/lib_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or tile_pipeline or full_image" > gpurun_out/r2_pytest_$v.log 2>&1
  tail -3 gpurun_out/r2_pytest_$v.log
done
AB_ROUNDS=3 timeout 600 python tools/ab_speed.py nerf_pl_b200/libnerf_pl_b200.so nerf_pl_b200/variants/lib_hoist.so nerf_pl_b200/variants/lib_pingpong.so > gpurun_out/r2_ab1.log 2>&1
cat gpurun_out/r2_ab1.log
