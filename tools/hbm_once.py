"""One launch of each HBM-bound kernel of the SOM path at a working set larger than L2: the command ncu wraps
(`ncu --set full --clock-control none -k regex:... python tools/hbm_once.py`) for DRAM bytes per launch.
Shapes as in tools/bench_hbm.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import somcb  # noqa: E402,F401
from somcb import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)

n = 1 << 26
w, m, v = (torch.randn(n, device=dev, generator=g) * 0.01 for _ in range(3))
v.abs_()
gr = torch.randn(n, device=dev, generator=g)
ops.adam_step(w, m, v, gr, 1e-4, 3)
del w, m, v, gr

for k in (4096, 32768, 262144):
    idx = torch.randint(0, k, (n,), device=dev, generator=g)
    ops.histogram(idx, k)
    del idx

for fm, pd, k in ((16384, (4, 4), 16384), (39063, (2, 2), 4096)):
    d = 4 * pd[0] * pd[1]
    x = torch.tanh(torch.randn(fm, 4, 32, 32, device=dev, generator=g))
    geom = ops.geometry(x.shape, pd)
    table = torch.randn(k, d, device=dev, generator=g)
    idx = torch.randint(0, k, (ops.n_patches_of(geom),), device=dev, generator=g)
    ops.quantize(idx, table, geom)
    if pd == (4, 4):
        ops.accumulate(x, geom, idx, table, k, want_sse=True)
        ops.neighbourhood_filter(table, k // 2)
    del x, table, idx

wt = torch.randn(262144, 256, device=dev, generator=g)
ops.prepare_codebook(wt)
keep = torch.nonzero(torch.rand(262144, device=dev, generator=g) < 0.5).flatten()
ops.gather_rows(wt, keep)
del wt, keep
rd = torch.rand(8, 1 << 22, device=dev, generator=g)
ix = torch.randint(0, 262144, (8, 1 << 22), device=dev, generator=g)
ops.merge_candidates(rd, ix)
torch.cuda.synchronize()
print("ok")
