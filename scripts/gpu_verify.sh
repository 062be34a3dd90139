#!/bin/bash
# round-2 final verification: full GPU test suite, smoke, both bench arms, ncu launch lists and --set full captures of the shipped kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2x_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2x_smoke.log 2>&1; tail -2 gpurun_out/r2x_smoke.log
timeout 600 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; tail -2 gpurun_out/r2x_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2x_bench.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['whole_step']['frac'],
      'fp16', d['fp16_params']['value'], 'm2', d['m2']['value'], d['m2']['roofline']['frac'], d['gpu_launches'], d['ref_cuda_build']['value'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_bench_ref.json 2>> gpurun_out/r2x_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 2 --warmup 3 --no-m2 --no-extras --no-cpu-baseline > gpurun_out/r2x_bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2x_m2_launches.csv python scripts/m2_bench.py --steps 1 --warmup 1 > gpurun_out/r2x_m2_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:lotd_pair' -s 4 -c 4 -o gpurun_out/r2x_pair_f32 -f python scripts/prof_step.py 4 > gpurun_out/r2x_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:march_kernel|march_compact' -c 2 -o gpurun_out/r2x_march -f python scripts/pack_bench.py > gpurun_out/r2x_ncu2.log 2>&1
ls -la gpurun_out | grep r2x
