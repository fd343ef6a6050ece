"""The one-kernel training step at BASELINE config 1 (512 patches, D = 64, K = 1024): graph-replayed us per step and
CTA 0's phase boundaries (globaltimer) of the last launch.  usage: python tools/small_probe.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import somcb  # noqa: E402
from somcb import ops  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tools"))
from graph_time import graph_time_us  # noqa: E402

dev = torch.device("cuda:0")
k, d, rng = 1024, 64, 512
x = torch.randn(8, 4, 32, 32, device=dev)
geom = ops.geometry(x.shape, (4, 4))
w = torch.randn(k, d, device=dev)
m, v = torch.zeros_like(w), torch.zeros_like(w)
td = torch.zeros(2, dtype=torch.int64, device=dev)
lo = torch.empty(1, dtype=torch.float64, device=dev)
fn = lambda: ops.step_small(x, geom, w, m, v, rng, 0.0, td, want_bmu=False, loss_out=lo)  # noqa: E731
print(f"graph-replayed step: {graph_time_us(fn, reps=20, replays=20):.1f} us")
fn()
torch.cuda.synchronize()
lib = ctypes.CDLL(somcb._lib.LIB_PATH)
out = (ctypes.c_ulonglong * 8)()
lib.som_debug_step_small_ns(out)
t = list(out)
print("phase boundaries (us from the first stamp):", [round((a - t[0]) / 1e3, 1) for a in t])
