import time, numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
from radae_b200 import RadeBatch
from radae_b200.batch import HostLink
from oracle.core import synth_features
S=1024
b=RadeBatch(S); b.channel_config(EbNodB=3.0, freq_offset_hz=-11.0, doppler_spread_hz=1.0, seed=5)
link=HostLink(b)
pin=lambda shape,dt: torch.empty(shape,dtype=dt).pin_memory().numpy()
base=synth_features(64,12*8,seed=1).reshape(64,8,432); fh=np.tile(base,(16,1,1))
feats=[]
for j in range(8):
    a=pin((S,432),torch.float32); a[...]=fh[:,j]; feats.append(a)
tx=pin((S,960,2),torch.float32).view(np.complex64).reshape(S,960); rx=pin((S,960,2),torch.float32).view(np.complex64).reshape(S,960)
T=np.zeros(4); 
for k in range(40):
    t0=time.perf_counter(); b.tx(feats[k%8],out=tx); t1=time.perf_counter(); b.channel(tx,out=rx); t2=time.perf_counter(); link.push(rx); t3=time.perf_counter(); link.rx(); t4=time.perf_counter()
    if k>=20: T+=np.array([t1-t0,t2-t1,t3-t2,t4-t3])
print('ms per step: tx %.3f channel %.3f push %.3f rx %.3f total %.3f'%tuple(list(T/20*1e3)+[T.sum()/20*1e3]))
b.profile_enable(True)
for k in range(20):
    b.tx(feats[k%8],out=tx); b.channel(tx,out=rx); link.push(rx); link.rx()
pr=b.profile_read(); b.profile_enable(False)
print('device ms per launch in e2e mode (kernels read/write pinned host buffers in place):')
for k,(ms,c) in pr.items(): print('  %-24s %.4f'%(k, ms/c))
print('  sum %.4f'%sum(ms/c for ms,c in pr.values()))
# raw copy speed
import ctypes
x=torch.empty(S*960*2,dtype=torch.float32).pin_memory(); d=torch.empty_like(x,device='cuda')
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(20): d.copy_(x,non_blocking=True)
torch.cuda.synchronize(); t1=time.perf_counter()
for _ in range(20): x.copy_(d,non_blocking=True)
torch.cuda.synchronize(); t2=time.perf_counter()
print('H2D GB/s',20*x.numel()*4/(t1-t0)/1e9,'D2H GB/s',20*x.numel()*4/(t2-t1)/1e9)
