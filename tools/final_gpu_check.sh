set -x
timeout 300 python -m pytest tests -m gpu -q --timeout=100 2>&1 | tail -3
timeout 250 python bench.py --steps 300 --warmup 5 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 300 gpurun_out/bench_r01.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:render_rays -s 1 -c 1 -f -o gpurun_out/prof_r01_render python tools/prof_run.py 32768 2 > gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:render_rays -s 2 -c 1 -f -o gpurun_out/prof_r01_bench1024 python tools/prof_bench1024.py > gpurun_out/ncu_full2.log 2>&1
timeout 200 compute-sanitizer --tool memcheck python tools/prof_run.py 512 1 > gpurun_out/sanitizer_memcheck.log 2>&1; tail -2 gpurun_out/sanitizer_memcheck.log
timeout 200 compute-sanitizer --tool racecheck python tools/prof_run.py 296 1 > gpurun_out/sanitizer_racecheck.log 2>&1; tail -2 gpurun_out/sanitizer_racecheck.log
ls -la gpurun_out/ | tail -12
