"""Config-S shapes: tensor-core BMU vs the FFMA kernel on random data (indices must agree except near-ties)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch
from somcb import ops
torch.manual_seed(0)
cases = [(4, 1, 2, 1000), (2, 2, 2, 1000), (4, 2, 2, 1000), (1, 4, 4, 1000), (3, 2, 2, 1000), (4, 1, 4, 1000),
         (4, 2, 2, 256), (4, 2, 2, 4096), (4, 1, 1, 300)]
for (c, ph, pw, k) in cases:
    d = c * ph * pw
    x = torch.tanh(torch.randn(64, c, 32, 32, device="cuda"))
    w = torch.tanh(torch.randn(k, d, device="cuda"))
    geom = ops.geometry(x.shape, (ph, pw))
    cn = ops.prepare_codebook(w)
    try:
        a = ops.bmu(x, geom, w, cn, variant=ops.SOM_BMU_TC3X)
        torch.cuda.synchronize()
        b = ops.bmu(x, geom, w, cn, variant=ops.SOM_BMU_FFMA)
        print(f"C={c} P={ph}x{pw} D={d} K={k}: diffs vs FFMA {int((a != b).sum())} of {a.numel()}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"C={c} P={ph}x{pw} D={d} K={k}: FAILED {str(e)[:100]}", flush=True)
        break
