#!/bin/bash
# M2 A/B with repeats: single-pass vs two-pass marcher, merge variants (per-step times in the JSON)
mkdir -p gpurun_out
m2() { NR3D_B200_LIB=$2 NR3D_B200_MARCH_SCRATCH_GB=$3 timeout 300 python scripts/m2_bench.py --steps 4 --warmup 2 > gpurun_out/r2v_m2_$1.json 2> gpurun_out/r2v_m2_$1.err
  python -c "import json; d=json.load(open('gpurun_out/r2v_m2_$1.json')); print('$1', round(d['value'],3), round(d['ms_per_step'],3), d['step_ms'], round(d['peak_mem_gb'],2))"; }
V=$PWD/nr3d_lib_b200/lib/variants
m2 warm "" 0
m2 twopass_a "" 0
m2 single_a "" 4
m2 twopass_b "" 0
m2 single_b "" 4
m2 runs_all_a $V/runs_all.so 0
m2 any_d8_a $V/any_d8.so 0
m2 any_d16_a $V/any_d16.so 0
m2 runs_all_b $V/runs_all.so 0
m2 any_d8_b $V/any_d8.so 0
m2 any_d16_b $V/any_d16.so 0
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv
