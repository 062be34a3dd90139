"""world_size-2 gloo tests of the N>1 host logic (sharding + the single all-reduce of dL/dparams), no GPU needed."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from nr3d_lib_b200 import dist as nd
    from oracle import lotd_oracle as O
    r, w, _ = nd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # every point is owned exactly once
    n = 1001
    b, e = nd.shard_range(n, r, w)
    owned = torch.zeros(n)
    owned[b:e] = 1
    torch.distributed.all_reduce(owned)
    assert torch.all(owned == 1)
    # data-parallel LoTD gradient: shard points, all-reduce dL/dparams, compare with the single-process gradient
    meta = O.OracleMeta(3, [8, 16, 24], [2, 2, 2], ["Dense", "Hash", "VM"], 512)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n, 3, generator=g).clamp(1e-6, 1 - 1e-6)
    p = torch.randn(meta.n_params, generator=g) * 0.1
    dLdy = torch.randn(n, meta.n_encoded_dims, generator=g)
    _, gp_local = O.bwd(meta, dLdy[b:e], x[b:e], p)
    gp_local = gp_local.clone()
    nd.allreduce_param_grads(gp_local, w)
    _, gp_full = O.bwd(meta, dLdy, x, p)
    assert torch.allclose(gp_local, gp_full, rtol=1e-12, atol=1e-12)
    # ray sharding keeps per-rank pack offsets local
    ro, rd, near, far, (rb, re_) = nd.shard_rays(torch.arange(30.).view(10, 3), torch.ones(10, 3), torch.zeros(10), torch.ones(10), r, w)
    assert ro.shape[0] == re_ - rb and float(ro[0, 0]) == 3.0 * rb
    assert nd.max_over_ranks(float(rank)) == float(world - 1)
    assert nd.sum_over_ranks(float(rank + 1)) == float(world * (world + 1) // 2)        # e.g. the samples every rank marched
    # host-fed read-back: every rank returns its own slice of the (identical) all-reduced gradient -> the host sees it exactly once
    host = torch.zeros(meta.n_params, dtype=gp_local.dtype)
    lo, hi = nd.shard_range(meta.n_params, r, w)
    host[lo:hi] = gp_local[lo:hi]
    torch.distributed.all_reduce(host)                                                  # (stands for the shared pinned host buffer)
    assert torch.equal(host, gp_local)
    # GradReducer modes on this backend: plain all-reduce, reduce-scatter (each rank keeps its slice), and the symmetric-memory mode, which
    # has no multicast here and must fall back to the plain all-reduce (and say so) instead of failing
    import types
    fake_meta = types.SimpleNamespace(n_params=meta.n_params)
    _, gp_mine = O.bwd(meta, dLdy[b:e], x[b:e], p)
    for mode in ("allreduce", "symm", "scatter"):
        red = nd.GradReducer(fake_meta, w, "cpu", mode=mode)
        out = red.reduce(gp_mine.clone())
        if mode == "scatter":
            slo, shi = red.slice_range(meta.n_params, r)
            assert torch.allclose(out[: shi - slo], gp_full[slo:shi], rtol=1e-12, atol=1e-12)
        else:
            assert torch.allclose(out, gp_full, rtol=1e-12, atol=1e-12)
        if mode == "symm":
            assert red.mode == "allreduce" and red.fallback_reason
        red.close()
    nd.barrier()
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([1]))
    torch.distributed.destroy_process_group()


def test_two_process_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}.npy") for r in range(world))


def test_shard_range_properties():
    from nr3d_lib_b200.dist import shard_range
    for n in (0, 1, 7, 8, 4194304, 16777217):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
