"""One-GPU probe of the C4 training step at a given local batch (the per-rank shape of an N-GPU data-parallel run):
CUDA-graph step vs eager step vs the sum of its parts, and the FP16-split BMU kernel's issue-loop counters.
usage: python tools/step_probe.py <fmaps> [reps]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import somcb  # noqa: E402
from somcb import ops  # noqa: E402
import bench  # noqa: E402

n_f = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
k, d, pd = 16384, 64, (4, 4)
n_rot = max(2, min(8, -(-(256 << 20) // (n_f * 16384))))
xs = [bench._fmaps(n_f, 5000 + b, dev) for b in range(n_rot)]
out = {"fmaps": n_f, "patches": n_f * 64, "rotation": n_rot}
for mode in ("alias", False):
    cb = bench._codebook(k, pd, dev)
    tr = somcb.SomTrainer(cb, lr=1e-4, neighbourhood_step=10 ** 9, use_cuda_graph=mode)
    for b in range(n_rot + 1):
        tr.step(xs[b % n_rot])
    it = [0]

    def step():
        tr.step(xs[it[0] % n_rot])
        it[0] += 1
    out["graph_step_ms" if mode else "eager_step_ms"] = bench._timed(step, reps, warm=3)
x = xs[0]
geom = ops.geometry(x.shape, pd)
w = cb.codebook.weight.data
rng = cb.neighbourhood_range
cn = ops.prepare_codebook(w)
wt = ops.neighbourhood_filter(w, rng)
bmu = ops.bmu(x, geom, w, cn)
packed = torch.empty(k * d + 4, dtype=torch.float32, device=dev)
ops.accumulate_packed(x, geom, bmu, wt, k, packed=packed)
grad = ops.neighbourhood_filter(packed[:k * d].view(k, d), rng)
wc, mc, vc = w.clone(), torch.zeros_like(w), torch.zeros_like(w)
tdev = torch.tensor([1, 0], dtype=torch.int64, device=dev)
lo = torch.empty(1, dtype=torch.float64, device=dev)
parts = {"filter_W": lambda: ops.neighbourhood_filter(w, rng), "norms": lambda: ops.prepare_codebook(w),
         "bmu": lambda: ops.bmu(x, geom, w, cn), "bmu_tf32": lambda: ops.bmu(x, geom, w, cn, variant=ops.SOM_BMU_TC_TF32),
         "accumulate": lambda: ops.accumulate_packed(x, geom, bmu, wt, k, packed=packed),
         "filter_Rbar": lambda: ops.neighbourhood_filter(packed[:k * d].view(k, d), rng),
         "adam": lambda: ops.adam_step_dp(wc, mc, vc, grad, d, 1e-4, tdev, packed[k * d:], loss_out=lo)}
out["parts_ms"] = {kk: bench._timed(fn, reps) for kk, fn in parts.items()}
out["sum_of_parts_ms"] = sum(v for kk, v in out["parts_ms"].items() if kk != "bmu_tf32")
ops.bmu(x, geom, w, cn)
torch.cuda.synchronize()
cyc = (ctypes.c_longlong * 5)()
lib = ctypes.CDLL(somcb._lib.LIB_PATH)
if lib.som_debug_tc_l16_cycles(cyc) == 0 and cyc[1] > 0:
    out["l16_issue_loop"] = {"cycles_per_tile": cyc[0] / cyc[1], "tiles": cyc[1], "wait_acc_empty_frac": cyc[2] / cyc[0],
                             "wait_b_frac": cyc[3] / cyc[0], "wait_a_frac": cyc[4] / cyc[0]}
print(json.dumps(out))
