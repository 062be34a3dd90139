#!/bin/bash
# one gpurun call: marcher (single pass vs two passes) tests + bench, backward-merge A/B round 2, M2 per variant
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pack_march_gpu.py tests/test_reference_wrappers_gpu.py tests/test_pipeline_gpu.py tests/test_lotd_fast_gpu.py -x -q -m gpu -k "march or lotd_fast or fast" > gpurun_out/r2u_pytest_march.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest_march.log
timeout 300 python scripts/pack_bench.py > gpurun_out/r2u_pack_march_bench.txt 2> gpurun_out/r2u_pack_march_bench.err; tail -4 gpurun_out/r2u_pack_march_bench.txt
export NR3D_AB_FP32_ONLY=1
timeout 120 python scripts/step_probe.py --tag default_any_d4 > gpurun_out/r2u_ab_merge.txt 2> gpurun_out/r2u_ab_merge.err
timeout 600 python scripts/ab_bench.py run runs_all,any_d2,any_d8,any_d16 >> gpurun_out/r2u_ab_merge.txt 2>> gpurun_out/r2u_ab_merge.err
python - <<'PY'
import json
for l in open('gpurun_out/r2u_ab_merge.txt'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(f"{d['tag']:18s} fwd {d['fwd']:.3f} bwd {d['bwd']:.3f} step {d['step']:.3f} diff {d.get('grad_max_abs_diff_vs_first')}")
PY
m2() { NR3D_B200_LIB=$2 NR3D_B200_MARCH_SCRATCH_GB=$3 timeout 300 python scripts/m2_bench.py --steps 2 --warmup 1 > gpurun_out/r2u_m2_$1.json 2> gpurun_out/r2u_m2_$1.err
  python -c "import json; d=json.load(open('gpurun_out/r2u_m2_$1.json')); print('$1', round(d['value'],3), round(d['ms_per_step'],3), round(d['peak_mem_gb'],2))"; }
m2 default_single "" 4
m2 default_twopass "" 0
for n in runs_all any_d2 any_d8 any_d16; do m2 $n $PWD/nr3d_lib_b200/lib/variants/$n.so 4; done
