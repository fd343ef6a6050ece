"""A few eager neighbourhood-filter calls (D = 64, range 8192): the target of an ncu capture of filter_tc_kernel.
usage: python tools/filter_probe.py [K ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import somcb  # noqa: E402,F401
from somcb import ops  # noqa: E402

dev = torch.device("cuda:0")
for k in [int(a) for a in sys.argv[1:]] or [16384, 3264]:
    w = torch.randn(k, 64, device=dev)
    for _ in range(3):
        out = ops.neighbourhood_filter(w, 8192)
    torch.cuda.synchronize()
    print(k, float(out[0, 0]))
