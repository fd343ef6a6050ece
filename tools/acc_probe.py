"""One eager accumulate_packed call per batch size on staged (patch-major) rows: the target of an ncu capture of the
accumulation's kernels (ncu -k regex:'seg_level|csort' ...).  usage: python tools/acc_probe.py [n_fmaps ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import somcb  # noqa: E402,F401
from somcb import ops  # noqa: E402

dev = torch.device("cuda:0")
kk, d = 16384, 64
wt = torch.randn(kk, d, device=dev)
for n_f in [int(a) for a in sys.argv[1:]] or [2048, 16384]:
    n = n_f * 64
    stage = torch.randn(n, d, device=dev)
    bmu = torch.randint(0, kk, (n,), device=dev)
    packed = torch.empty(kk * d + 4, device=dev)
    for _ in range(3):
        ops.accumulate_packed(stage, ops.flat_geometry(n, d), bmu, wt, kk, packed=packed)
    torch.cuda.synchronize()
    print(n, float(packed[:8].sum()))
