#!/bin/bash
# one gpurun call: A/B of the backward merge variants (scripts/ab_bench.py), fast-path tests of the default build, M2 with old / new merge
mkdir -p gpurun_out
export NR3D_AB_FP32_ONLY=1
V=runs_all,runs_d8,runs_d2,any_d8,any_neigh_d8,any_neigh_d2,any_neigh_d32,any_neigh_d8_occ1280,link3_d8,link3_neigh_d8,link5_d8
timeout 900 python scripts/ab_bench.py run $V > gpurun_out/r2t_ab_merge.txt 2> gpurun_out/r2t_ab_merge.err
cat gpurun_out/r2t_ab_merge.txt | cut -c1-260
timeout 600 python -m pytest tests/test_lotd_fast_gpu.py -x -q -m gpu > gpurun_out/r2t_pytest_fast.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2t_pytest_fast.log
for n in runs_all any_neigh_d8 any_d8; do
  NR3D_B200_LIB=$PWD/nr3d_lib_b200/lib/variants/$n.so timeout 300 python scripts/m2_bench.py --steps 2 --warmup 1 > gpurun_out/r2t_m2_$n.json 2> gpurun_out/r2t_m2_$n.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2t_m2_$n.json')); print('$n', d['value'], d['ms_per_step'])"
done
