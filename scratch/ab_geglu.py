import sys; sys.path.insert(0,'/root/repo')
import torch
from frido_b200.program import Program
from frido_b200 import _lib as L
dev=torch.device('cuda:0')
M,K,N=16384,384,3072
x=torch.randn(M,K,device=dev); w=torch.randn(N,K,device=dev)/20; b=torch.randn(N,device=dev)
outs={}
for act in (3,4,0):
    P=Program(dev,'ab',engine='bf16x3')
    out=torch.zeros(M,N//2 if act else N,device=dev)
    if act==4:
        # python helper computes n_out for act==3 only; emulate
        P.conv(__import__('frido_b200.program',fromlist=['Src']).Src(x,K,0,0,K,1), w, out, B=1,Hin=1,Win=M,Hout=1,Wout=M,Cout=N,bias=b,act=4,o_sp=N//2,o_sb=M*N//2)
    else:
        P.linear(x,w,out,M=M,K=K,N=N,bias=b,act=act)
    P.prepare_weights(); P.run(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): P.run()
    e1.record(); torch.cuda.synchronize()
    print('act',act,'us per launch',e0.elapsed_time(e1)/20*1e3)
    outs[act]=out
print('max diff fast vs erff', (outs[3]-outs[4]).abs().max().item())
