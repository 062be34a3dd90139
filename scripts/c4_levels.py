import os, sys
sys.path.insert(0, os.getcwd())
import torch
from nr3d_lib_b200.bindings import _lotd as mine
from scripts.quick_bench import timeit
dev = torch.device("cuda:0"); N = 2*1024*1024; B = 8
torch.manual_seed(42)
x = torch.rand(N, 3, device=dev).clamp(1e-6, 1-1e-6)
bi = torch.randint(0, B, (N,), device=dev); bi[torch.rand(N, device=dev) < 0.01] = -1
bis = torch.sort(bi).values
for res, nf, tp in ((8,4,"Dense"),(16,4,"Dense"),(32,4,"VM"),(64,4,"VM"),(128,2,"CP"),(256,2,"CP")):
    meta = mine.LoDMeta(3, [res], [nf], [tp], None, False)
    E = meta.n_encoded_dims
    params = torch.randn(B*meta.n_params, device=dev)*1e-2
    g = torch.randn(N, E, device=dev)*1e-2; ddx = torch.randn(N,3,device=dev)
    for name, b in (("random bidx", bi), ("sorted bidx", bis)):
        kw = dict(batch_inds=b, batch_offsets=None, batch_data_size=None, max_level=None)
        y, dydx = mine.lod_fwd(meta, x, params, need_input_grad=True, **kw)
        t1 = timeit(lambda: mine.lod_bwd(meta, g, x, params, None, need_input_grad=False, need_param_grad=True, **kw), iters=5, warm=2)
        t2 = timeit(lambda: mine.lod_bwd_bwd_input(meta, ddx, g, x, params, dydx, need_dLdinput_ddLdoutput=False, need_dLdinput_dparams=True, need_dLdinput_dinput=False, **kw), iters=5, warm=2)
        print(f"{tp:6s} res {res:4d} F {nf} n_params {meta.n_params:7d} {name}: bwd dparam {t1:7.3f} ms | bwd_bwd dparam {t2:7.3f} ms", flush=True)
