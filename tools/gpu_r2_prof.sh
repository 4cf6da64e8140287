#!/bin/bash
# ncu --set full captures (forward at two sizes, the training step's four big kernels), trained-weights run, sanitizer.
set -x
export PYTHONUNBUFFERED=1
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_rays -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_bench1024 python tools/prof_bench1024.py > gpurun_out/ncu_full_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_rays -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_render python tools/prof_run.py 32768 2 > gpurun_out/ncu_full_r.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:render_rays_kernel|chain_bwd_kernel|wgrad_kernel|head_bwd_kernel' -s 15 -c 5 -f -o gpurun_out/prof_${TAG}_train python tools/prof_train.py 1024 2 plain > gpurun_out/ncu_full_t.log 2>&1
tail -2 gpurun_out/ncu_full_t.log
timeout 600 python tools/train_sharp_weights.py 6000 gpurun_out/sharp_weights.npz 2>&1 | tail -16
if [ -f nerf_pl_b200/variants/lib_sanitize.so ]; then
  export NERFB200_LIB=nerf_pl_b200/variants/lib_sanitize.so
  for tool in racecheck memcheck synccheck; do
    timeout 300 compute-sanitizer --tool $tool python tools/prof_run.py 296 1 > gpurun_out/sanitizer_$tool.log 2>&1
    tail -2 gpurun_out/sanitizer_$tool.log
  done
  timeout 400 compute-sanitizer --tool memcheck python tools/prof_train.py 296 1 plain > gpurun_out/sanitizer_memcheck_train.log 2>&1
  tail -2 gpurun_out/sanitizer_memcheck_train.log
fi
