import sys, torch, numpy as np
sys.path.insert(0,'/root/repo')
from libcpab_b200 import Cpab, ops, _lib
T=Cpab([3,3],backend='pytorch',device='gpu'); theta=T.sample_transformation(64); grid=T.uniform_meshgrid([256,256])
with torch.no_grad(): gt=T.transform_grid(grid,theta)
data=torch.rand(64,1,256,256,device='cuda')
flush=torch.empty(256*1024*1024//4,device='cuda')
def t(fn, pre=None, n=10):
    ts=[]
    for _ in range(n):
        if pre: pre()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts), min(ts)
print('warm', t(lambda: ops.interpolate_forward(data,gt,[256,256])))
print('flushed', t(lambda: ops.interpolate_forward(data,gt,[256,256]), pre=lambda: flush.add_(1.0)))
print('flushed fwd kernel', t(lambda: ops.forward(grid, ops.theta_to_trels(theta, torch.as_tensor(T.params.basis,dtype=torch.float32).cuda().t().contiguous(), [3,3], 50)[1], [3,3], 50), pre=lambda: flush.add_(1.0)))
x=torch.rand(16*1024*1024,device='cuda'); y=torch.empty_like(x)
print('copy 64MB warm', t(lambda: y.copy_(x)))
print('copy 64MB flushed', t(lambda: y.copy_(x), pre=lambda: flush.add_(1.0)))
_lib.profile_enable(True)
for _ in range(5):
    flush.add_(1.0); ops.interpolate_forward(data,gt,[256,256])
torch.cuda.synchronize(); print('lib profile flushed', _lib.profile_read('interp_fwd'))
_lib.profile_enable(True)
for _ in range(5):
    ops.interpolate_forward(data,gt,[256,256])
torch.cuda.synchronize(); print('lib profile warm', _lib.profile_read('interp_fwd'))
