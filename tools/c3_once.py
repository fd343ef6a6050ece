"""One C3-shape BMU (4096 fmaps, P=32 -> D=4096, K=512) a few times: for ncu launch lists."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch
from somcb import ops
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.tanh(torch.randn(4096, 4, 32, 32, generator=g, device="cuda"))
w = torch.tanh(torch.randn(512, 4096, generator=g, device="cuda"))
geom = ops.geometry(x.shape, (32, 32))
cn = ops.prepare_codebook(w)
for _ in range(4):
    idx = ops.bmu(x, geom, w, cn)
torch.cuda.synchronize()
print("ok", int(idx.sum()))
