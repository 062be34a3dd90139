#!/bin/bash
# ncu --set full of the shipped fp16-table pair kernels and of the sort kernels (end of round 2)
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k 'regex:lotd_pair' -s 4 -c 2 -o gpurun_out/r2y_pair_f16 -f python scripts/prof_step.py 4 4194304 half > gpurun_out/r2y_ncu1.log 2>&1
timeout 100 ncu --set full --clock-control none -k 'regex:sort_hist|sort_scan|sort_scatter' -s 6 -c 3 -o gpurun_out/r2y_sort -f python scripts/prof_step.py 4 > gpurun_out/r2y_ncu2.log 2>&1
ls -la gpurun_out | grep r2y
