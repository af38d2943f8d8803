import sys, torch, numpy as np
sys.path.insert(0,'/root/repo')
from libcpab_b200 import Cpab, ops, _lib
from tools.gpu_probe import timeit
T=Cpab([10,10],backend='pytorch',device='gpu',volume_perservation=True)
for n,size in ((128,[512,512]),(512,[512,512])):
    theta=T.sample_transformation(n); grid=T.uniform_meshgrid(size)
    with torch.no_grad(): gt=T.transform_grid(grid,theta)
    data=torch.rand(n,1,*size,device='cuda'); g2=torch.randn_like(data)
    byts=n*size[0]*size[1]*16
    for var in range(5):
        _lib.set_tuning("interp_variant",var)
        med,best=timeit(lambda: ops.interpolate_forward(data,gt,size))
        print(n,"fwd variant",var,"ms %.4f GB/s %.0f"%(med,byts/med/1e6), flush=True)
    _lib.set_tuning("interp_variant",2)
    med,best=timeit(lambda: ops.interpolate_backward(data,gt,g2,True,False))
    print(n,"bwd dgrid ms %.4f GB/s %.0f"%(med,n*size[0]*size[1]*24/med/1e6))
    del data,g2,gt
