"""Multi-GPU data parallelism for the hot path: one process per GPU, rays / points sharded, parameters replicated,
one all-reduce of dL/dparams per step (the reference has no counterpart, its DDP support is
"discarded", nr3d_lib/config.py:74-75).

Rays (and every sample pack they own) are independent through march -> encode -> composite, so the data path needs no
collective; only the LoTD parameter gradient is summed across ranks.  The functions work with any torch.distributed
backend (NCCL over NVLink/NVSwitch on the B200 box, gloo in the CPU tests).
"""
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise the default process group from torchrun's environment.  Returns (rank, world_size, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of ``n`` units (rays or points) owned by ``rank``; the first ``n % world``
    ranks get one extra unit so that every unit is owned exactly once."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_rays(rays_o: torch.Tensor, rays_d: torch.Tensor, near: torch.Tensor, far: torch.Tensor, rank: int, world: int):
    """This rank's contiguous block of rays; each rank later builds its own pack_infos starting at 0."""
    b, e = shard_range(rays_o.shape[0], rank, world)
    return rays_o[b:e], rays_d[b:e], near[b:e], far[b:e], (b, e)


def allreduce_param_grads(grad: torch.Tensor, world: Optional[int] = None, async_op: bool = False):
    """Sum dL/dparams over all ranks in place (one all-reduce of n_params values per step).  No-op for one process."""
    if not dist.is_available() or not dist.is_initialized():
        return None
    if (world or dist.get_world_size()) == 1:
        return None
    return dist.all_reduce(grad, op=dist.ReduceOp.SUM, async_op=async_op)


class GradReducer:
    """Sum of dL/dparams over the ranks, once per step, in one of three ways (`mode`):

      "allreduce"  one all-reduce of the whole table after the scatter, on the step's stream (default)
      "bucketed"   the scatter runs fine levels first; the all-reduce of their part of the table (2/3 of the bytes for the NGP ladder)
                   is issued asynchronously and overlaps the scatter of the coarse levels.  MEASURED SLOWER on B200 (2 GPUs: 2.47 vs 2.15 ms
                   per step, profiles/r2_bench_2gpu_*.json): two scatter launches re-read the records and dL_dy rows, and NCCL's CTAs take
                   SMs from the scatter they overlap -- the whole all-reduce is only 0.1 - 0.2 ms.  Kept for the record / other shapes.
      "scatter"    reduce-scatter: every rank ends with ITS contiguous 1/world slice of the summed table (what a sharded optimizer or the
                   host-fed driver needs: pipeline.HostFedLoTDStep returns each slice to the host from its own rank) -- half the NVLink bytes

      "symm"       the all-reduce runs on a torch symmetric-memory buffer with the NVSwitch multicast / in-switch reduction
                   (`torch.ops.symm_mem.multimem_all_reduce_`): 0.131 ms instead of NCCL's 0.220 ms for the 48.5 MB table on 8 B200s
                   (profiles/r2_allreduce_probe_8gpu.json).  The reducer owns the buffer; the gradient is copied into it (48.5 MB, ~15 us) --
                   letting the scatter write straight into the symmetric buffer (`symm_direct=True`, through
                   `bindings._lotd.set_param_grad_allocator`) was measured SLOWER (2 GPUs: 2.44 vs 2.12 ms per step: the L2 reductions of
                   the scatter do not like the multicast-mapped allocation).  The returned gradient IS the symmetric buffer -- consume it
                   (optimizer step) before the next reduce.  Falls back to "allreduce" when symmetric memory or multicast is not
                   available (`self.mode` then says so).

    Usage per step:  (lod_bwd runs, the hook fires)  ->  out = reducer.reduce(grad)   # "scatter": out is the rank's slice, else `grad`.
    The hooks are installed with `bindings._lotd.set_grad_bucket_hook` / `set_param_grad_allocator`; call close() to remove them."""

    def __init__(self, meta, world: int, device, mode: str = "allreduce", symm_direct: bool = False):
        from .bindings import _lotd
        self._lotd, self.meta, self.world, self.mode, self.device = _lotd, meta, int(world), mode, device
        self.pending = []
        self.active = self.world > 1 and dist.is_available() and dist.is_initialized()
        if mode not in ("allreduce", "bucketed", "scatter", "symm"):
            raise RuntimeError(f"GradReducer: unknown mode {mode!r}")
        if self.active and mode == "bucketed":
            _lotd.set_grad_bucket_hook(self._on_part)
        self._slice = None
        self._symm_buf = self._symm_op = self._symm_group = None
        self.fallback_reason = None
        if self.active and mode == "symm":
            try:
                import torch.distributed._symmetric_memory as symm
                self._symm_group = dist.group.WORLD.group_name
                n = int(meta.n_params)
                buf = symm.empty(n, dtype=torch.float32, device=torch.device(device))
                symm.rendezvous(buf, self._symm_group)
                op = torch.ops.symm_mem.multimem_all_reduce_
                buf.zero_()
                op(buf, "sum", self._symm_group)           # first use: fails here, not inside a step, when multicast is unavailable
                torch.cuda.synchronize(torch.device(device))
                self._symm_buf, self._symm_op = buf, op
                if symm_direct:
                    _lotd.set_param_grad_allocator(self._alloc)
            except Exception as e:      # no symmetric memory / no NVLS on this box: plain NCCL all-reduce
                self.fallback_reason = str(e)[:200]
                self.mode = "allreduce"

    def _alloc(self, numel, dtype, device):
        buf = self._symm_buf
        if buf is None or numel != buf.numel() or dtype != buf.dtype or torch.device(device) != buf.device:
            return None
        buf.zero_()
        return buf

    def _on_part(self, grad, lo, hi):
        # issued on NCCL's stream behind everything already queued on the current stream: overlaps what the caller launches next
        self.pending.append((lo, hi, dist.all_reduce(grad[lo:hi], op=dist.ReduceOp.SUM, async_op=True)))

    def reduce(self, grad: torch.Tensor) -> torch.Tensor:
        if not self.active:
            return grad
        if self.mode == "bucketed":
            covered = sorted((lo, hi) for lo, hi, _ in self.pending)
            pos = 0
            for lo, hi in covered:
                if lo != pos:
                    break
                pos = hi
            if pos != grad.shape[0]:        # the hook did not fire (generic path / batched call): reduce the whole table now
                for _, _, w in self.pending:
                    w.wait()
                self.pending = []
                dist.all_reduce(grad, op=dist.ReduceOp.SUM)
                return grad
            for _, _, w in self.pending:
                w.wait()                    # the current stream waits for the collective; the host does not block
            self.pending = []
            return grad
        if self.mode == "symm":
            if grad.data_ptr() != self._symm_buf.data_ptr():      # a gradient that did not come from our allocator (generic path, other sizes)
                if grad.numel() != self._symm_buf.numel() or grad.dtype != self._symm_buf.dtype:
                    dist.all_reduce(grad, op=dist.ReduceOp.SUM)
                    return grad
                self._symm_buf.copy_(grad)
            self._symm_op(self._symm_buf, "sum", self._symm_group)
            return self._symm_buf
        if self.mode == "scatter":
            n = grad.shape[0]
            per = (n + self.world - 1) // self.world
            if per * self.world != n:       # pad to a multiple of the world size (the NGP table divides evenly for 2 / 4 / 8 ranks)
                g = torch.zeros(per * self.world, dtype=grad.dtype, device=grad.device)
                g[:n] = grad
            else:
                g = grad
            out = torch.empty(per, dtype=grad.dtype, device=grad.device)
            dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM)
            return out
        dist.all_reduce(grad, op=dist.ReduceOp.SUM)
        return grad

    def slice_range(self, n: int, rank: int):
        """[begin, end) of the table that `reduce` returns on `rank` in "scatter" mode."""
        per = (n + self.world - 1) // self.world
        return rank * per, min(n, (rank + 1) * per)

    def close(self):
        if self.active and self.mode == "bucketed":
            self._lotd.set_grad_bucket_hook(None)
        if self._symm_buf is not None:
            self._lotd.set_param_grad_allocator(None)


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Run this process (and allocate its pinned host buffers) on the NUMA node the GPU hangs off: with 8 ranks on one box the default
    placement puts every rank's staging buffers on node 0 and half of the host<->device traffic crosses the socket interconnect.
    Tries, in order, CPU affinity (first-touch then allocates locally) and a MPOL_PREFERRED memory policy.  Returns what it did (for the
    bench line); never raises."""
    info = {"gpu": int(local_rank), "node": None, "cpu_affinity": None, "mempolicy": None}
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bus = None
        if all(hasattr(props, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            # torch exposes the three numbers, sysfs wants "dddd:bb:dd.f"
            bus = f"{int(props.pci_domain_id):04x}:{int(props.pci_bus_id):02x}:{int(props.pci_device_id):02x}.0"
        if bus is None or not os.path.exists(f"/sys/bus/pci/devices/{bus}/numa_node"):
            import subprocess
            bus = subprocess.run(["nvidia-smi", f"--id={local_rank}", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True,
                                 text=True, timeout=10).stdout.strip().lower()
            if len(bus.split(":")[0]) == 8:      # nvidia-smi prints an 8-digit domain, sysfs uses 4
                bus = bus[4:]
        info["pci"] = bus
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        want = cpus & allowed
        if want:
            os.sched_setaffinity(0, want)
            info["cpu_affinity"] = f"{len(want)} cpus of node {node}"
        else:
            info["cpu_affinity"] = f"node {node} cpus not in the allowed set ({len(allowed)} cpus)"
        try:
            import ctypes
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            MPOL_PREFERRED, SYS_set_mempolicy = 1, 238      # x86-64
            rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))
            info["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy failed (errno %d)" % ctypes.get_errno()
        except Exception as e:  # pragma: no cover
            info["mempolicy"] = f"unavailable ({e})"
    except Exception as e:
        info["error"] = str(e)[:200]
    return info


def max_over_ranks(value: float, device=None) -> float:
    """max of a python float over all ranks (device-side timing is reported as the slowest rank's)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    """sum of a python float over all ranks (e.g. the number of samples every rank marched this step)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def shutdown():
    """Tear the default process group down (silences NCCL's leak warning at interpreter exit)."""
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
