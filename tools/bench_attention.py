import os, sys, torch
sys.path.insert(0, "/root/repo")
import mmsam_b200
from mmsam_b200 import kernels as K
nh=16
for (Bp,Kh,Kw,bias) in ((8,64,64,True),(8,64,64,False),(8,32,32,True),(200,14,14,True),(200,14,14,False)):
    T=Kh*Kw
    qkv=torch.randn(Bp,T,3*nh*64,device="cuda").to(torch.bfloat16)
    th=K.relpos_table(torch.randn(2*Kh-1,64,device="cuda")*0.2,Kh) if bias else None
    tw=K.relpos_table(torch.randn(2*Kw-1,64,device="cuda")*0.2,Kw) if bias else None
    out=K.attention(qkv,nh,(Kh,Kw),th,tw)
    torch.cuda.synchronize()
    s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): K.attention(qkv,nh,(Kh,Kw),th,tw,out=out)
    e.record(); torch.cuda.synchronize()
    ms=s.elapsed_time(e)/5
    print(f"Bp={Bp} T={T} bias={bias}: {ms:.3f} ms {4.0*Bp*nh*T*T*64/ms/1e9:.0f} TFLOP/s")
