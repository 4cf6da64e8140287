#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
for EXP in 0 1 2 3; do
  NERFB200_WG_EXP=$EXP timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wgrad_kernel --csv --log-file gpurun_out/r2_wgexp_$EXP.csv python tools/prof_train.py 1024 2 plain > /dev/null 2>&1
  echo "wg exp $EXP"; grep wgrad gpurun_out/r2_wgexp_$EXP.csv | awk -F'","' '{print $NF}' | tail -2
done
for V in nostage nostore; do
  NERFB200_LIB=nerf_pl_b200/variants/lib_$V.so timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"render_rays_kernel|chain_bwd_kernel" --csv --log-file gpurun_out/r2_st_$V.csv python tools/prof_train.py 1024 2 plain > /dev/null 2>&1
  echo "variant $V"; grep -E "render_rays|chain_bwd" gpurun_out/r2_st_$V.csv | awk -F'","' '{print $5, $NF}' | tail -3
done
timeout 300 python -m pytest tests -m gpu -q -k "adam or host_buffer or reference_style" 2>&1 | tail -3
