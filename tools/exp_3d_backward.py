import os, sys
sys.path.insert(0, '/root/repo')
import torch
from libcpab_b200 import Cpab, _lib, ops
from libcpab_b200.transformer import _basis
from tools.gpu_probe import timeit
torch.manual_seed(1234)
tess=[4,4,4]; T=Cpab(tess,backend='pytorch',device='gpu')
theta=T.sample_transformation(4); grid=T.uniform_meshgrid([128,128,128]); nP=grid.shape[1]
B,Bt=_basis(T.params,theta.device,theta.dtype); As,Tr=ops.theta_to_trels(theta,Bt,tess,50)
gout=torch.randn(4,3,nP,device='cuda')
for seg,stage,block in ((0,-1,128),(3,0,128),(3,1,128),(5,0,128),(5,1,128),(3,0,256),(10,1,128)):
    _lib.set_tuning("bwd_seg",seg); _lib.set_tuning("bwd_stage",stage); _lib.set_tuning("bwd_block",block)
    med,best=timeit(lambda: ops.backward_theta(grid,As,B,gout,tess,50))
    print("seg",seg,"stage",stage,"block",block,"ms %.3f"%med, flush=True)
