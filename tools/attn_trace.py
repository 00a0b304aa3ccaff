import os, sys, torch
sys.path.insert(0, "/root/repo")
import mmsam_b200
from mmsam_b200 import kernels as K
nh=16; Bp,Kh,Kw=8,64,64; T=Kh*Kw
qkv=torch.randn(Bp,T,3*nh*64,device="cuda").to(torch.bfloat16)
bias = os.environ.get("BIAS","0")=="1"
th=K.relpos_table(torch.randn(2*Kh-1,64,device="cuda")*0.2,Kh) if bias else None
tw=K.relpos_table(torch.randn(2*Kw-1,64,device="cuda")*0.2,Kw) if bias else None
out=K.attention(qkv,nh,(Kh,Kw),th,tw); torch.cuda.synchronize()
