set -x
timeout 300 python -m pytest tests -m gpu -q --timeout=100 2>&1 | tail -3
timeout 250 python bench.py --steps 300 --warmup 5 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 200 gpurun_out/bench_r01.json
export NERFB200_LIB=nerf_pl_b200/variants/lib_sanitize.so
timeout 250 compute-sanitizer --tool racecheck python tools/prof_run.py 296 1 > gpurun_out/sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck.log
timeout 200 compute-sanitizer --tool memcheck python tools/prof_run.py 512 1 > gpurun_out/sanitizer_memcheck.log 2>&1; tail -2 gpurun_out/sanitizer_memcheck.log
timeout 200 compute-sanitizer --tool synccheck python tools/prof_run.py 512 1 > gpurun_out/sanitizer_synccheck.log 2>&1; tail -2 gpurun_out/sanitizer_synccheck.log
