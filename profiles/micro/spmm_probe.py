import sys, os
sys.path.insert(0, os.getcwd())
import torch, ctypes as C
from distgcn_b200 import engine as E
from profiles.micro.stream_probe import er_batch_device
dev=torch.device("cuda",0)
gp,rp,ci=er_batch_device(16384,0,dev)
ctx=E.Context(0)
b=E.DeviceBatch(ctx, graph_ptr=gp,row_ptr=rp,col_idx=ci)
z=torch.randn(rp.numel()-1,32,device=dev); y=torch.empty_like(z)
for _ in range(3): E.spmm_laplacian(ctx,b,z,y)
ctx.synchronize()
ms=C.c_double(); E.check(ctx._lib.dg_timer_start(ctx.handle))
for _ in range(10): E.spmm_laplacian(ctx,b,z,y)
E.check(ctx._lib.dg_timer_stop(ctx.handle,C.byref(ms))); print("spmm us", ms.value*100, ctx.last_kernel)
