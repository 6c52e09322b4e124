#!/bin/bash
# Run on the GPU box (gpurun): ncu launch lists + one full capture per kernel; text/CSV extracts under gpurun_out/
# (the .ncu-rep files are too large to ship back: raw-metric and source pages are exported on the box).
set -x
mkdir -p gpurun_out
for wl in point cheetah; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$wl.csv python tools/run_iters.py $wl 4 > gpurun_out/ncu_l_$wl.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"rollout_kernel|process_fused|policy_grad|policy_hvp|meta_update" -s 25 -c 12 -o /tmp/prof_$wl python tools/run_iters.py $wl 3 > gpurun_out/ncu_f_$wl.log 2>&1
  ncu -i /tmp/prof_$wl.ncu-rep --page raw --csv > gpurun_out/raw_$wl.csv 2>/dev/null
  ncu -i /tmp/prof_$wl.ncu-rep --page details --csv > gpurun_out/details_$wl.csv 2>/dev/null
done
ncu -i /tmp/prof_point.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:policy_hvp --launch-count 1 2>/dev/null | gzip > gpurun_out/src_hvp_point.csv.gz
ncu -i /tmp/prof_point.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:policy_grad --launch-count 1 2>/dev/null | gzip > gpurun_out/src_grad_point.csv.gz
ls -la gpurun_out
