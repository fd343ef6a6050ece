"""C2 BMU search (39 063 fmaps, P = 2, K = 4096): ms per call and CTA 0's cycles per tile (clock-independent).
usage: python tools/c2_probe.py [n_fmaps]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import somcb  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda:0")
n_f = int(sys.argv[1]) if len(sys.argv) > 1 else 39063
x = bench._fmaps(n_f, 11, dev)
cb = bench._codebook(4096, (2, 2), dev)
for _ in range(3):
    idx = cb.get_patches_bmu(x, reshape=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    idx = cb.get_patches_bmu(x, reshape=True)
e1.record()
torch.cuda.synchronize()
lib = ctypes.CDLL(somcb._lib.LIB_PATH)
out = (ctypes.c_longlong * 2)()
lib.som_debug_tc_cycles(out)
print(f"C2 bmu: {e0.elapsed_time(e1) / 10:.3f} ms per call; CTA 0: {out[0]} cycles, {out[1]} tiles, "
      f"{out[0] / max(1, out[1]):.0f} cycles per tile")
