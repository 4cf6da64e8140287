#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
for PLAN in 0 1; do for COPY in 8192 32768; do
  NERFB200_WG_PLAN=$PLAN NERFB200_WG_COPY=$COPY timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:wgrad_kernel --csv --log-file gpurun_out/r2_wg_${PLAN}_${COPY}.csv python tools/prof_train.py 1024 2 plain > /dev/null 2>&1
  echo "plan $PLAN copy $COPY"; grep wgrad gpurun_out/r2_wg_${PLAN}_${COPY}.csv | awk -F, '{print $(NF-2), $(NF-1), $NF}' | tail -4
done; done
NERFB200_WG_PLAN=1 NERFB200_WG_COPY=32768 timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_wgrad2 python tools/prof_train.py 1024 1 plain > gpurun_out/r2_ncu_wgrad2.log 2>&1
timeout 300 python -m pytest tests -m gpu -q -k "training or gradients or deterministic or upstream" 2>&1 | tail -3
