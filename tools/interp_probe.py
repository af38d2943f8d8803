import sys, os, json, torch
sys.path.insert(0, '/root/repo')
from libcpab_b200 import Cpab, _lib, ops
from libcpab_b200.transformer import _basis
import numpy as np
def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(iters):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for tess,n,size in (([10,10],128,[512,512]),([3,3],64,[256,256]),([4,4,4],4,[128,128,128])):
    T=Cpab(tess,backend='pytorch',device='gpu'); theta=T.sample_transformation(n); grid=T.uniform_meshgrid(size)
    B,Bt=_basis(T.params,theta.device,theta.dtype); As,Tr=ops.theta_to_trels(theta,Bt,tess,50); gt=ops.forward(grid,Tr,tess,50)
    data=torch.rand(n,1,*size,device='cuda'); g2=torch.randn_like(data)
    nP=grid.shape[1]; ndim=len(tess)
    for var in range(5):
        _lib.set_tuning('interp_variant',var)
        ms=timeit(lambda: ops.interpolate_forward(data,gt,size))
        print(json.dumps(dict(kind='interp_fwd',tess=tess,variant=var,ms=ms,gbps=n*nP*(4*ndim+8)/ms/1e6)))
    ms=timeit(lambda: ops.interpolate_backward(data,gt,g2,True,False))
    print(json.dumps(dict(kind='interp_bwd',tess=tess,ms=ms,gbps=n*nP*(8*ndim+8)/ms/1e6)))
