"""FFMA vs tensor-core BMU on small batches: where the static variant rule should switch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch
from somcb import ops
dev = "cuda:0"
def t(fn, reps=30):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (p, k) in ((2, 4096), (2, 512), (4, 1024), (4, 16384), (8, 2048), (32, 512)):
    d = 4 * p * p
    for fm in (1, 2, 4, 8, 16, 32, 64, 128):
        x = torch.tanh(torch.randn(fm, 4, 32, 32, device=dev))
        w = torch.tanh(torch.randn(k, d, device=dev))
        geom = ops.geometry(x.shape, (p, p)); cn = ops.prepare_codebook(w)
        n = ops.n_patches_of(geom)
        a = t(lambda: ops.bmu(x, geom, w, cn, variant=ops.SOM_BMU_FFMA))
        b = t(lambda: ops.bmu(x, geom, w, cn, variant=ops.SOM_BMU_TC3X))
        print(f"P={p} D={d} K={k} n={n}: ffma {a:.1f} us, tc {b:.1f} us  -> {'TC' if b < a else 'FFMA'}", flush=True)
