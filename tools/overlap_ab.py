"""A/B inside one process: the C4 step with the W~ filter on a side stream beside the search (SomTrainer(overlap_filter=
True)) against the one-stream order, graph-replayed, alternating several times.  usage: python tools/overlap_ab.py [n_fmaps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import somcb  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda:0")
n_f = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
xs = [bench._fmaps(n_f, 7000 + b, dev) for b in range(4)]
trainers = {}
for ov in (True, False):
    cb = bench._codebook(16384, (4, 4), dev)
    tr = somcb.SomTrainer(cb, lr=1e-4, neighbourhood_step=10 ** 9, use_cuda_graph="alias", overlap_filter=ov)
    for i in range(8):
        tr.step(xs[i % 4])
    trainers[ov] = tr
torch.cuda.synchronize()
res = {True: [], False: []}
for rep in range(4):
    for ov in (True, False, False, True):
        tr = trainers[ov]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(12):
            tr.step(xs[i % 4])
        e1.record()
        torch.cuda.synchronize()
        res[ov].append(e0.elapsed_time(e1) / 12)
for ov in (True, False):
    v = sorted(res[ov])
    print(f"{n_f * 64} patches, overlap_filter={ov}: median {v[len(v) // 2]:.4f} ms, min {v[0]:.4f}, max {v[-1]:.4f}")
