import sys; sys.path[:0]=[".","quantized-autoregression-image-generator_b200","tests"]
import torch, somcb
from somcb import ops
from oracle.step_oracle import synthetic_fmaps, trained_like_codebook
dev="cuda:0"
pd,k=(4,4),1024
w0=trained_like_codebook(k,pd,7)
trs=[]
for small in (True,False):
    cb=somcb.Codebook(patch_dim=pd,image_dim=(32,32),image_channel=4,num_embeddings=k,init_neighbour_range=k//2)
    with torch.no_grad(): cb.codebook.weight.copy_(w0)
    trs.append(somcb.SomTrainer(cb.to(dev),lr=1e-4,neighbourhood_step=200,small_step_kernel=small))
for step in range(8):
    x=synthetic_fmaps(8,123+step).to(dev)
    # BMU of the small kernel vs ops.bmu on the SAME weights (before the step)
    w=trs[0].cb.codebook.weight.data.clone()
    geom=ops.geometry(x.shape,pd)
    ref=ops.bmu(x,geom,w,variant=ops.SOM_BMU_FFMA)
    l0=float(trs[0].step(x)); l1=float(trs[1].step(x))
    b0=trs[0].last_bmu
    mism=int((b0!=ref).sum())
    wd=float((trs[0].cb.codebook.weight.data-trs[1].cb.codebook.weight.data).norm()/trs[1].cb.codebook.weight.data.norm())
    print(step,"loss",l0,l1,"rel",abs(l0-l1)/l1,"bmu mismatches vs ffma on same W",mism,"w rel",wd)
    if mism:
        bad=torch.nonzero(b0!=ref).flatten()[:5]
        print("  bad patches",bad.tolist(),"small",b0[bad].tolist(),"ffma",ref[bad].tolist())
        flat=somcb.patchify(x,pd).reshape(-1,64)[bad].double()
        da=(flat-w[b0[bad]].double()).norm(dim=1); db=(flat-w[ref[bad]].double()).norm(dim=1)
        print("  d64 small pick",da.tolist(),"ffma pick",db.tolist(),"rel gap",((da-db)/db).tolist())
        xs=flat[0].float(); 
        for u in (int(b0[bad][0]),int(ref[bad][0])):
            c=w[u]; print("   unit",u,"score fp32",float((xs*c).sum()-0.5*(c*c).sum()),"fp64",float((flat[0]*c.double()).sum()-0.5*(c.double()**2).sum()))
