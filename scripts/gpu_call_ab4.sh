#!/bin/bash
# instruction diet of pair_geo + predicated staging, sort workspace reuse, records token: tests of everything on the fast path + timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lotd_fast_gpu.py tests/test_pipeline_gpu.py tests/test_fused_gpu.py tests/test_reference_wrappers_gpu.py tests/test_full_size_gpu.py -x -q -m gpu > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2w_pytest.log
export NR3D_AB_FP32_ONLY=1
timeout 120 python scripts/step_probe.py --tag diet > gpurun_out/r2w_step.txt 2> gpurun_out/r2w_step.err
timeout 120 python scripts/step_probe.py --tag diet_half --half >> gpurun_out/r2w_step.txt 2>> gpurun_out/r2w_step.err
python - <<'PY'
import json
for l in open('gpurun_out/r2w_step.txt'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(f"{d['tag']:18s} sort {d['sort']:.3f} fwd {d['fwd']:.3f} bwd {d['bwd']:.3f} step {d['step']:.3f} {d['msamples_per_s']:.1f}")
PY
timeout 300 python scripts/m2_bench.py --steps 4 --warmup 2 > gpurun_out/r2w_m2.json 2> gpurun_out/r2w_m2.err
python -c "import json; d=json.load(open('gpurun_out/r2w_m2.json')); print('m2', round(d['value'],3), round(d['ms_per_step'],3), d['step_ms'], round(d['peak_mem_gb'],2))"
timeout 300 python scripts/next_rows_bench.py n3train > gpurun_out/r2w_n3.txt 2>&1; tail -6 gpurun_out/r2w_n3.txt
