"""Where does the host-fed step lose time?  Times the same step with (a) no copies, (b) H2D only, (c) D2H only, (d) both (overlapped)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import ngp_cfg
from nr3d_lib_b200.bindings import _lotd

dev = torch.device("cuda:0")
meta = _lotd.LoDMeta(*ngp_cfg()); meta.c_sort_points = True
N = 4 * 1024 * 1024
torch.manual_seed(0)
x = torch.rand(N, 3, device=dev).clamp(1e-6, 1 - 1e-6)
params = (torch.rand(meta.n_params, device=dev) * 2 - 1) * 1e-4
x_host = x.cpu().pin_memory()
g_host = [torch.empty(meta.n_params).pin_memory() for _ in range(2)]
comp = torch.cuda.current_stream(dev)
h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
xbuf = [torch.empty_like(x) for _ in range(2)]


def run(steps, do_h2d, do_d2h):
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]
    for e in free: e.record(comp)
    keep = []
    def prefetch(b):
        h2d.wait_event(free[b])
        with torch.cuda.stream(h2d):
            xbuf[b].copy_(x_host, non_blocking=True)
            ready[b].record(h2d)
    if do_h2d: prefetch(0)
    for k in range(steps):
        b = k & 1
        if do_h2d:
            if k + 1 < steps: prefetch(b ^ 1)
            comp.wait_event(ready[b])
        xx = xbuf[b] if do_h2d else x
        _lotd.clear_sort_cache()
        y, _ = _lotd.lod_fwd(meta, xx, params, need_input_grad=False)
        _, g = _lotd.lod_bwd(meta, y * 1e-4, xx, params, None, need_input_grad=False, need_param_grad=True)
        free[b].record(comp)
        if do_d2h:
            done = torch.cuda.Event(); done.record(comp)
            d2h.wait_event(done)
            with torch.cuda.stream(d2h):
                g_host[b].copy_(g, non_blocking=True)
            g.record_stream(d2h)
            keep.append(g)
            keep[:] = keep[-3:]
    if do_d2h:
        e = torch.cuda.Event(); e.record(d2h); comp.wait_event(e)


for name, a, b in (("no copies", False, False), ("H2D only", True, False), ("D2H only", False, True), ("both", True, True)):
    run(3, a, b); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(comp); run(20, a, b); e1.record(comp); torch.cuda.synchronize()
    print(f"{name:10s} {e0.elapsed_time(e1) / 20:.3f} ms/step")
# raw copy bandwidths
for name, fn in (("H2D 48 MB", lambda: xbuf[0].copy_(x_host, non_blocking=True)), ("D2H 48.5 MB", lambda: g_host[0].copy_(params, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(10)]; e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10:.3f} ms per copy")
