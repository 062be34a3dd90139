#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native LoTD hot path.

Metric (BASELINE.json): "LoTD-Hash 16L fwd+bwd Msamples/s" on configs[1]: 16-level NGP LoTD (gen_ngp_cfg defaults,
T=2^19, F=2, fp32 params; nr3d_lib/models/grid_encodings/lotd/lotd_cfg.py:48-57), 4 Mi uniform random 3-D points per GPU.

One step  = lod_fwd(need_input_grad=False) + lod_bwd(need_param_grad=True) through the reference-facing operator
            surface (nr3d_lib_b200.bindings._lotd == nr3d_lib.bindings._lotd) with its DEFAULT settings (the cell-sorted fast path is
            what a drop-in user gets) [+ one NCCL all-reduce of dL/dparams if N > 1].  Consecutive steps see DIFFERENT points (two
            point sets alternate), so every forward sorts and every backward reuses the forward's records after the on-device check.
`value`   = whole-job samples/s with x and dL_dy already resident in HBM (device-timed, max over ranks).
`e2e`     = the same step driven from HOST buffers: x comes from pinned host memory every step (H2D inside the timed
            region), dL_dy is the step's own output y (loss = |y|^2 / 2: stand-in for the decoder's backward, costs no extra pass over [N, 32]),
            and the step's result dL/dparams is read back to the host (D2H inside the timed region).
`--impl reference` times the CPU port of the reference's algorithm (oracle/lotd_port.c, plain C, fp32, OpenMP over all host threads) on a
            bounded sample of the same workload -- the reference has no CPU implementation of this path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "LoTD-Hash 16L fwd+bwd Msamples/s"
UNIT = "Msamples/s"
N_POINTS = 4 * 1024 * 1024          # per GPU (weak scaling)
# algorithmic bytes per sample, SURVEY.md section 8d (fp32 params, L=16, F=2, D=3):
#   fwd 12 (x) + 1024 (corner reads) + 128 (y write); bwd 12 (x) + 128 (dL_dy read) + 1024 (gradient scatter); + 23 (table zero-init/read)
BYTES_FWD, BYTES_BWD, BYTES_TABLE = 12 + 1024 + 128, 12 + 128 + 1024, 23
BYTES_PER_SAMPLE = BYTES_FWD + BYTES_BWD + BYTES_TABLE   # 2351
BYTES_PER_SAMPLE_F16 = (12 + 512 + 64) + (12 + 64 + 512) + 12   # 1188 (SURVEY.md 8d, fp16 parameter tables)
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the two dominant kernels at this workload, from the committed
# `ncu --set full` capture profiles/r2_pair_ncu_summary.txt (4 Mi points): the table is L2 resident, DRAM only sees x / y / dL_dy
NCU_DRAM_BYTES = {"lotd_pair_bwd_kernel (dL/dparam scatter)": 672.3e6 + 36.0e6, "lotd_pair_fwd_kernel (corner gather)": 153.6e6 + 499.6e6}


def ngp_cfg(min_res=16, n_levels=16, scale=1.382, log2_T=19, F=2):
    res = (min_res * scale ** np.arange(n_levels)).astype(int).tolist()
    types = ["Dense" if r ** 3 <= 2 ** log2_T else "Hash" for r in res]
    return (3, res, [F] * n_levels, types, 2 ** log2_T, False)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._stop_evt, self.sm_max = index, [], set(), threading.Event(), None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().splitlines()[0].split(",")
                self.samples.append(float(out[0]))
                self.sm_max = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=10)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm restated for the host (kind "port"), all threads, bounded sample
# ------------------------------------------------------------------------------------------------------------------
def cpu_fwd_bwd(steps, warmup, seed=42, budget_s=20.0):
    """The reference's algorithm for this path on the host: plain-C fp32 port of its Dense/Hash kernels (oracle/lotd_port.c, pinned by
    the golden vectors of the reference's CUDA build), OpenMP over all host threads, on a bounded sample of the workload.  The
    sample size comes from a short probe so that warmup + steps fit in `budget_s` seconds."""
    from oracle import lotd_oracle as O, lotd_port as P
    # all host threads this process may use -- asked for explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
    cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    meta = O.OracleMeta(*ngp_cfg())
    rs = np.random.RandomState(seed)
    params = ((rs.rand(meta.n_params).astype(np.float32) * 2 - 1) * 1e-4).astype(np.float32)
    grad = np.zeros(meta.n_params, dtype=np.float32)

    def one(n):
        x = np.clip(rs.rand(n, 3).astype(np.float32), 1e-6, 1 - 1e-6)
        dL_dy = (rs.randn(n, meta.n_encoded_dims) * 1e-4).astype(np.float32)
        t0 = time.perf_counter()
        grad[:] = 0.0                                  # the step's zero-init of dL/dparam is part of the algorithm
        P.fwd(meta, x, params, cores)
        P.bwd_param(meta, dL_dy, x, cores, out=grad)
        return time.perf_counter() - t0

    one(4096)                                  # page in / thread-pool start
    probe = one(65536)
    per_point = probe / 65536
    sample = int(min(N_POINTS, max(65536, budget_s / max(1, warmup + steps) / per_point)))
    sample = 1 << (sample.bit_length() - 1)     # power of two <= estimate
    times = []
    for it in range(warmup + steps):
        dt = one(sample)
        if it >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    return dict(value=sample / sec / 1e6, unit=UNIT, cores=cores, kind="port",
                sample=f"{sample} of the {N_POINTS} points per step, {steps} steps after {warmup} warm-up, fp32, plain-C port of the reference's "
                       f"Dense/Hash kernels (oracle/lotd_port.c), OpenMP with {cores} threads"), sec, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    base, sec, sample = cpu_fwd_bwd(args.steps, min(args.warmup, 2), budget_s=60.0)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus):
    return {"workload": "configs[1]: 16-level Hash LoTD (NGP gen_ngp_cfg: res 16..2049, 6 Dense + 10 Hash levels, T=2^19, F=2), "
                        "4Mi uniform points per GPU, lod_fwd + lod_bwd(dL/dparam)",
            "points_per_gpu": N_POINTS, "n_params": 12131648, "param_dtype": "f32", "parallelism": f"dp{n_gpus} (points sharded, params replicated, "
                                                                                                 "1 all-reduce of dL/dparams per step)",
            "l2_policy": "inputs larger than L2 (x 48 MB + dL_dy 512 MB + y 512 MB per step >> 126 MB); the 48.5 MB parameter table is "
                         "L2-resident by nature of the workload"}


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e leg without copy/compute overlap (single stream)")
    ap.add_argument("--no-sort", action="store_true", help="use the generic (unsorted, feature-major) kernels instead of lotd_fast.cu")
    ap.add_argument("--no-m2", action="store_true", help="skip the secondary M2 block (march + encode + composite rays/s)")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "symm", "bucketed", "allreduce"],
                    help="N > 1: how dL/dparams is summed (dist.GradReducer): one all-reduce after the scatter -- on a symmetric-memory buffer with NVSwitch "
                         "multicast (`symm`; measured 2.165 vs 2.226 ms / step at 8 GPUs but 2.46 vs 2.12 at 2) or plain NCCL (`allreduce`); `auto` (default) "
                         "takes symm from 8 ranks on --, or fine levels first "
                         "with their all-reduce overlapping the coarse levels' scatter (measured slower: 2.47 vs 2.15 ms / step at 2 GPUs, "
                         "profiles/r2_bench_2gpu_*.json)")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational legs (reference CUDA build, generic path, torch CPU baselines)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE line (the JSON): libraries that print there (e.g. "NCCL version ..." at communicator
    # creation) are sent to stderr for the duration of the run, the result goes out through the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    from nr3d_lib_b200 import _lib, dist as ndist
    from nr3d_lib_b200.bindings import _lotd

    rank, world, local = ndist.init_from_env("nccl")
    if world != args.gpus and rank == 0 and world > 1:
        print(f"[bench] note: WORLD_SIZE={world} differs from --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
    n_gpus = max(world, 1)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa = ndist.bind_to_gpu_numa(local)       # before any pinned allocation: staging buffers on the GPU's own NUMA node

    meta = _lotd.LoDMeta(*ngp_cfg())
    if args.no_sort:
        meta.c_sort_points = False             # reference strides + generic kernels (the default is the cell-sorted fast path)
    torch.manual_seed(42 + rank)
    N = N_POINTS
    # two point sets alternate between steps: a training step never sees the previous step's points, so every forward sorts
    xs = [torch.rand(N, 3, device=dev).clamp(1e-6, 1 - 1e-6) for _ in range(2)]
    gen = torch.Generator(device=dev).manual_seed(42)                      # parameters are replicated: same seed on every rank
    params = (torch.rand(meta.n_params, device=dev, generator=gen) * 2 - 1) * 1e-4
    dL_dy = torch.randn(N, meta.n_encoded_dims, device=dev) * 1e-4
    xs_host = [x.cpu().pin_memory() for x in xs]
    grad_host = torch.empty(meta.n_params, dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream(dev)
    ar_mode = args.allreduce if args.allreduce != "auto" else ("symm" if n_gpus >= 8 else "allreduce")
    reducer = ndist.GradReducer(meta, n_gpus, dev, mode=ar_mode)
    step_no = [0]

    def step_resident(m=meta, p=params, gy=dL_dy):
        x = xs[step_no[0] & 1]
        step_no[0] += 1
        y, _ = _lotd.lod_fwd(m, x, p, need_input_grad=False)
        _, g = _lotd.lod_bwd(m, gy, x, p, None, need_input_grad=False, need_param_grad=True)
        reducer.reduce(g)
        return y, g

    # e2e: the same step fed from / drained to HOST buffers through the public host-fed driver (pipeline.HostFedLoTDStep):
    # every step copies its 48 MB of points host->device and its 48.5 MB of gradients device->host; the copies of neighbouring
    # steps overlap the kernels (3 streams, double buffers).  `--e2e-serial` issues everything on one stream instead.
    from nr3d_lib_b200.pipeline import HostFedLoTDStep
    grad_host2 = [grad_host, torch.empty(meta.n_params, dtype=torch.float32).pin_memory()]
    pipe = HostFedLoTDStep(meta, params, N, dev, grad_of_y=lambda y: y, world=n_gpus, rank=rank)   # dL/dy = y (loss = |y|^2 / 2): stand-in for the decoder's backward, no extra pass

    def run_e2e(steps):
        if args.e2e_serial:
            for k in range(steps):
                xd = xs_host[k & 1].to(dev, non_blocking=True)
                y, _ = _lotd.lod_fwd(meta, xd, params, need_input_grad=False)
                _, g = _lotd.lod_bwd(meta, y, xd, params, None, need_input_grad=False, need_param_grad=True)
                out = pipe.reducer.reduce(g)
                lo, hi = pipe.reducer.slice_range(g.shape[0], rank) if n_gpus > 1 else (0, g.shape[0])
                grad_host2[k % 2][lo:hi].copy_(out[: hi - lo], non_blocking=True)
            return
        pipe.prefetch(xs_host[0])
        for k in range(steps):
            pipe.step(xs_host[(k + 1) & 1] if k + 1 < steps else None, grad_host2[k % 2])
        pipe.drain()

    def timed(fn, steps, sampler=None):
        ndist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ndist.barrier()
        return ndist.max_over_ranks(e0.elapsed_time(e1), dev)              # ms, slowest rank

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = _lib.launch_count()
    ms_total = timed(step_resident, args.steps)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = n_gpus * N / (ms_step * 1e-3) / 1e6

    reducer.close()        # ("bucketed": the level-group hook must not fire inside the e2e leg, which has its own reduce-scatter)
    run_e2e(2)
    ms_e2e = timed(lambda: run_e2e(args.steps), 1) / args.steps
    e2e_value = n_gpus * N / (ms_e2e * 1e-3) / 1e6

    # per-kernel durations for the roofline block: CUDA events on the launch stream around each call, inside this run
    def kernel_ms(fn, iters):
        fn()                                   # untimed warm-up
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))
    it = max(3, min(args.steps, 10))
    flip = [0]
    def fwd_with_sort():
        flip[0] ^= 1
        _lotd.lod_fwd(meta, xs[flip[0]], params, need_input_grad=False)          # new points: fingerprint + sort + gather
    ms_fwd = kernel_ms(fwd_with_sort, it)      # includes the per-step point sort of the fast path
    ms_bwd = kernel_ms(lambda: _lotd.lod_bwd(meta, dL_dy, xs[flip[0]], params, None, need_input_grad=False, need_param_grad=True), it)   # fingerprint (hit) + scatter
    peak, peak_src = measured_peak_gbs()
    # the two dominant KERNELS on their own (no sort, no fingerprint check): the C-ABI entry points called directly on the sorted records
    import ctypes
    lib = _lib.get_lib()
    recs, _ = _lotd._sorted_points(xs[0], expect_new=True)
    y_tmp = torch.empty(N, meta.n_encoded_dims, device=dev)
    g_tmp = torch.zeros(meta.n_params, device=dev)
    E, P_ = meta.n_encoded_dims, meta.n_pseudo_levels
    k_fwd = kernel_ms(lambda: _lib.check(lib.nr3d_lotd_fwd_sorted(ctypes.byref(meta._c), _lib.dtype_code(torch.float32), N, recs.data_ptr(), None, 1,
                                                                  params.data_ptr(), meta.n_levels, y_tmp.data_ptr(), E, 1, stream.cuda_stream)), it)
    k_bwd = kernel_ms(lambda: _lib.check(lib.nr3d_lotd_bwd_param_sorted(ctypes.byref(meta._c), _lib.dtype_code(torch.float32), N, recs.data_ptr(), None, 1,
                                                                        dL_dy.data_ptr(), E, 1, meta.n_levels, 0, P_, g_tmp.data_ptr(), stream.cuda_stream)), it)
    del y_tmp, g_tmp
    dom = "lotd_pair_bwd_kernel (dL/dparam scatter)" if k_bwd >= k_fwd else "lotd_pair_fwd_kernel (corner gather)"
    dom_ms, dom_bytes = (k_bwd, BYTES_BWD + BYTES_TABLE) if k_bwd >= k_fwd else (k_fwd, BYTES_FWD)
    achieved = N * dom_bytes / (dom_ms * 1e-3) / 1e9
    whole = value * 1e6 / n_gpus * BYTES_PER_SAMPLE / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_DRAM_BYTES[dom],
                "traffic_source": "ncu --set full capture of this kernel at this workload, profiles/r2_pair_ncu_summary.txt (bytes per launch)",
                "peak_source": peak_src, "algorithmic_bytes_per_sample": {"fwd": BYTES_FWD, "bwd": BYTES_BWD, "table": BYTES_TABLE},
                "kernel_ms": {"lotd_pair_fwd_kernel": k_fwd, "lotd_pair_bwd_kernel": k_bwd},
                "ms": {"lod_fwd (sort + gather)": ms_fwd, "lod_bwd (fingerprint check + memset + scatter)": ms_bwd},
                "note": "86 % of the algorithmic bytes are gathers / reductions that hit the L2-resident 48.5 MB table: the binding resources are the L1 line "
                        "rate (forward, 97 % of peak in the ncu capture) and the L2 reduction-unit packet rate + issue slots (backward: 88 % / 79 %), DRAM carries 0.65 - 0.71 GB per launch",
                "whole_step": {"achieved": whole, "frac": whole / peak, "bytes_per_sample": BYTES_PER_SAMPLE}}

    # secondary number of metric M1: the same step with fp16 parameter tables (y, dL_dy and dL/dparams in half)
    params_h, dL_dy_h = params.half(), dL_dy.half()
    def step_half():
        step_resident(meta, params_h, dL_dy_h)
    for _ in range(3):
        step_half()
    ms_half = timed(step_half, args.steps) / args.steps
    half_value = n_gpus * N / (ms_half * 1e-3) / 1e6
    whole_h = half_value * 1e6 / n_gpus * BYTES_PER_SAMPLE_F16 / 1e9
    fp16_block = {"value": half_value, "unit": UNIT, "ms_per_step": ms_half, "algorithmic_bytes_per_sample": BYTES_PER_SAMPLE_F16,
                  "roofline_frac_whole_step": whole_h / peak}

    # second half of BASELINE.json's metric ("full march+composite rays/s @1/2/4/8 GPU"): 1024^2 rays per GPU through
    # march -> LoTD -> density head -> alpha-composite, forward + backward, one all-reduce per step (scripts/m2_bench.py)
    m2_block = None
    if not args.no_m2:
        pass
        from scripts.m2_bench import run_m2
        # configs[2]: 1024^2 rays on one GPU; configs[4]: 4096^2 rays over 8 GPUs = 2 Mi rays per GPU
        m2_block = run_m2(dev, rank, n_gpus, rays=(4096 * 4096 // 8 if n_gpus == 8 else 1024 * 1024), steps=3, warmup=1)

    # informational legs, outside every timed region above (N = 1, rank 0): the same step (a) with the fast path switched off -- the
    # reference's strides and our generic kernels -- and (b) on the REFERENCE'S OWN CUDA KERNELS compiled for sm_100 (oracle/_ref/_lotd.so,
    # "reported in every benchmark run, same process, same box"), on the same tensors
    generic_block = ref_cuda_block = None
    if n_gpus == 1 and not args.no_extras:
        def fwd_bwd_ms(fwd, bwd, iters=5):
            for _ in range(2):
                fwd(); bwd()
            return kernel_ms(fwd, iters), kernel_ms(bwd, iters)
        meta_g = _lotd.LoDMeta(*ngp_cfg())
        meta_g.c_sort_points = False
        f_ms, b_ms = fwd_bwd_ms(lambda: _lotd.lod_fwd(meta_g, xs[0], params, need_input_grad=False),
                                lambda: _lotd.lod_bwd(meta_g, dL_dy, xs[0], params, None, need_input_grad=False, need_param_grad=True))
        generic_block = {"value": N / ((f_ms + b_ms) * 1e-3) / 1e6, "unit": UNIT, "ms": {"lod_fwd": f_ms, "lod_bwd": b_ms},
                         "note": "LoDMeta.c_sort_points = False: reference strides (feature-major y), generic kernels"}
        try:
            ref_so = os.path.join(ROOT, "oracle", "_ref", "_lotd.so")
            if os.path.exists(ref_so):
                import importlib.util
                spec = importlib.util.spec_from_file_location("_lotd", ref_so)
                ref = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(ref)
                m_r = ref.LoDMeta(*ngp_cfg())
                f_ms, b_ms = fwd_bwd_ms(lambda: ref.lod_fwd(m_r, xs[0], params, None, None, None, None, False),
                                        lambda: ref.lod_bwd(m_r, dL_dy, xs[0], params, None, None, None, None, None, False, True))
                ref_cuda_block = {"value": N / ((f_ms + b_ms) * 1e-3) / 1e6, "unit": UNIT, "ms": {"lod_fwd": f_ms, "lod_bwd": b_ms},
                                  "note": "the reference's own hash-only CUDA kernels (csrc/lotd, unmodified, nvcc 12.9 -O3 sm_100 via oracle/build_ref.py), "
                                          "same tensors, same process, CUDA events"}
        except Exception as e:      # the checker build is optional on the box
            ref_cuda_block = {"unavailable": str(e)[:200]}

    reducer.close()
    ndist.barrier()
    ndist.shutdown()
    if rank != 0:
        return 0
    cpu_base = torch_cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        cpu_base, _, _ = cpu_fwd_bwd(steps=5, warmup=1, budget_s=15.0)
        if not args.no_extras:
            # BASELINE.json north_star: "next to the reference's pure-PyTorch grid_sample / F.cumprod path timed on the box's own host
            # cores (core count stated) in the same run as a reported baseline only" -- configs[0] and the dense-batch compositing
            from oracle import torch_baselines as TB
            torch_cpu = {"grid_sample_config0": TB.time_config0(), "cumprod_composite": TB.time_composite()}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n_gpus), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(xs_host[0].numel() * 4) * n_gpus,
                    "d2h_bytes_per_step": int(grad_host.numel() * 4),
                    "note": "whole-job bytes per step: every rank feeds its own 4 Mi points from pinned host memory, dL_dy is the step's own y on the device "
                            "(loss |y|^2 / 2), the summed dL/dparams (one reduce-scatter) is read back to the host once -- each rank returns its 1/N slice; "
                            + ("single stream" if args.e2e_serial else "copies of neighbouring steps overlap the kernels (pipeline.HostFedLoTDStep)")},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_base, "fp16_params": fp16_block, "m2": m2_block,
            "collective": (None if n_gpus == 1 else {"value_leg": {"bucketed": "bucketed: NCCL all-reduce of levels 8-15 overlaps the scatter of levels 0-7",
                                                                    "allreduce": "one NCCL all-reduce (sum) of the 48.5 MB fp32 dL/dparams after the scatter, on the step's stream",
                                                                    "symm": "one all-reduce (sum) of the 48.5 MB fp32 dL/dparams after the scatter: torch symmetric memory, NVSwitch "
                                                                            "multicast / in-switch reduction (torch.ops.symm_mem.multimem_all_reduce_) after a 48.5 MB device copy into the symmetric buffer"}[reducer.mode],
                                                     "requested": args.allreduce, "fallback_reason": reducer.fallback_reason,
                                                     "e2e_leg": "reduce-scatter, each rank returns its slice"}),
            "numa": numa, "generic_path": generic_block, "ref_cuda_build": ref_cuda_block, "torch_cpu_baselines": torch_cpu}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    return 0


if __name__ == "__main__":
    sys.exit(main())
