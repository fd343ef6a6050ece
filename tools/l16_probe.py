"""FP16-split BMU kernel (bmu_tc_l16): time and issue-loop counters for a few batch sizes.
usage: python tools/l16_probe.py <P> <K> <fmaps> [<fmaps> ...]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import somcb  # noqa: E402
from somcb import ops  # noqa: E402
import bench  # noqa: E402

p, k = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
d = 4 * p * p
w = torch.tanh(torch.randn(k, d, generator=torch.Generator().manual_seed(1))).to(dev)
cn = ops.prepare_codebook(w)
lib = ctypes.CDLL(somcb._lib.LIB_PATH)
for n_f in (int(a) for a in sys.argv[3:]):
    x = bench._fmaps(n_f, 11, dev)
    geom = ops.geometry(x.shape, (p, p))
    ms = bench._timed(lambda: ops.bmu(x, geom, w, cn, variant=ops.SOM_BMU_TC_F16), 5)
    torch.cuda.synchronize()
    cyc = (ctypes.c_longlong * 5)()
    lib.som_debug_tc_l16_cycles(cyc)
    n = ops.n_patches_of(geom)
    print(f"P={p} D={d} K={k} fmaps={n_f} patches={n} tiles={-(-n // 128)}: {ms:.4f} ms, "
          f"{2.0 * k * d * n / ms / 1e9:.1f} TFLOP/s fp32-faithful; issue loop {cyc[0] / max(1, cyc[1]):.0f} cycles/tile "
          f"({cyc[1]} tiles), waiting acc {cyc[2] / max(1, cyc[0]):.2f} B {cyc[3] / max(1, cyc[0]):.2f} A {cyc[4] / max(1, cyc[0]):.2f}",
          flush=True)
