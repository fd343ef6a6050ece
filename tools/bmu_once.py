"""A few BMU launches of one shape, nothing else: the command ncu wraps for launch lists / full captures.
usage: python tools/bmu_once.py <fmaps> <P> <K> [reps]      (4x32x32 fmaps, P x P patches)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
from somcb import ops  # noqa: E402

b, p, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
d = 4 * p * p
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.empty(b, 4, 32, 32, device="cuda")
for lo in range(0, b, 8192):
    hi = min(b, lo + 8192)
    x[lo:hi] = torch.tanh(torch.randn(hi - lo, 4, 32, 32, generator=g, device="cuda"))
w = torch.tanh(torch.randn(k, d, generator=g, device="cuda"))
geom = ops.geometry(x.shape, (p, p))
cn = ops.prepare_codebook(w)
for _ in range(reps):
    idx = ops.bmu(x, geom, w, cn)
torch.cuda.synchronize()
print("ok", int(idx.sum()))
