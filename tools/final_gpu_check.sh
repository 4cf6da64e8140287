#!/bin/bash
# One-call round-end validation on a B200 box (gpurun -- 'bash tools/final_gpu_check.sh'):
# GPU tests, bench, ncu launch list + two --set full captures, compute-sanitizer on the
# -DNERFB200_SANITIZE build (python tools/build_variants.py sanitize=-DNERFB200_SANITIZE first).
# Afterwards here: python tools/summarize_ncu.py rNN ; cp the bench / sanitizer files to profiles/.
set -x
TAG=${1:-r01}
timeout 300 python -m pytest tests -m gpu -q --timeout=100 2>&1 | tail -3
timeout 250 python bench.py --steps 300 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:render_rays -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_render python tools/prof_run.py 32768 2 > gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:render_rays -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_bench1024 python tools/prof_bench1024.py > gpurun_out/ncu_full2.log 2>&1
if [ -f nerf_pl_b200/variants/lib_sanitize.so ]; then
  export NERFB200_LIB=nerf_pl_b200/variants/lib_sanitize.so
  for tool in racecheck memcheck synccheck; do
    timeout 250 compute-sanitizer --tool $tool python tools/prof_run.py 296 1 > gpurun_out/sanitizer_$tool.log 2>&1
    tail -2 gpurun_out/sanitizer_$tool.log
  done
fi
