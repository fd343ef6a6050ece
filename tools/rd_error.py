"""Error of the reduced distance rd = ||c||^2 - 2 x.c returned by the BMU kernels against fp64, relative to
||c||^2 + 2 sum|x c| (the magnitude the accumulator sees): the tensor-core variant shows the truncation bias of the
fp32 TMEM accumulator (negative, growing with the number of chained MMAs), the FFMA variant is unbiased.
usage: python tools/rd_error.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch, somcb
from somcb import ops
dev="cuda"
g=torch.Generator(device=dev).manual_seed(1)
for (n,p,k) in ((512,32,512),(4096,8,2048),(4096,4,4096)):
    d=4*p*p
    x=torch.tanh(torch.randn(n,4,32,32,generator=g,device=dev))
    pool=torch.tanh(torch.randn(max(8,k*d//4096+1),4,32,32,generator=g,device=dev))
    w=somcb.patchify(pool,(p,p)).reshape(-1,d)[:k].contiguous()
    geom=ops.geometry(x.shape,(p,p))
    cn=ops.prepare_codebook(w)
    out={}
    for name,v in (("tc",ops.SOM_BMU_TC3X),("ffma",ops.SOM_BMU_FFMA)):
        idx,rd=ops.bmu(x,geom,w,cn,want_rd=True,variant=v)
        flat=somcb.patchify(x,(p,p)).reshape(-1,d).double()
        wi=w.double()[idx]
        true=(wi*wi).sum(1)-2*(flat*wi).sum(1)
        scale=(wi*wi).sum(1)+2*(flat*wi).abs().sum(1)
        err=((rd.double()-true).abs()/scale)
        out[name]=(float(err.max()),float(err.mean()),float(((rd.double()-true)/scale).mean()))
    print("D=%d K=%d: rd error / (|c|^2 + 2 sum|x c|): tc max %.2e mean %.2e signed mean %+.2e | ffma max %.2e mean %.2e signed %+.2e"%(d,k,*out["tc"],*out["ffma"]))
