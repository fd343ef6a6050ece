"""GPU time of small ops without the Python / launch overhead: the op is captured REPS times into one CUDA graph and
the replay is timed (per-op time includes the dependent-launch gaps inside the graph, as in the trainer's step graph)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402


def graph_time_us(fn, reps=20, replays=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * replays) * 1e3


if __name__ == "__main__":
    import somcb  # noqa: F401
    from somcb import ops
    import bench
    dev = torch.device("cuda:0")
    print("neighbourhood filter, D = 64, range 8192 (band 1217): GPU us per call (graph), tensor-core path / FFMA path")
    for k in (1536, 2048, 3264, 5312, 9408, 16384):
        w = torch.randn(k, 64, device=dev)
        t_tc = graph_time_us(lambda: ops.neighbourhood_filter(w, 8192))
        t_ff = graph_time_us(lambda: ops.neighbourhood_filter(w, 8192, tensor_cores=False))
        print(f"  K'={k}: tc {t_tc:.1f}  ffma {t_ff:.1f}")
    print("accumulate_packed (C4 shape), NCHW input vs patch-major staging rows: GPU us per call")
    kk, d = 16384, 64
    wt = torch.randn(kk, d, device=dev)
    for n_f in (2048, 4096, 16384):
        x = bench._fmaps(n_f, 3, dev)
        geom = ops.geometry(x.shape, (4, 4))
        n = ops.n_patches_of(geom)
        bmu = torch.randint(0, kk, (n,), device=dev)
        stage = somcb.patchify(x, (4, 4)).reshape(n, d).contiguous()
        packed = torch.empty(kk * d + 4, device=dev)
        t_a = graph_time_us(lambda: ops.accumulate_packed(x, geom, bmu, wt, kk, packed=packed), reps=5)
        t_b = graph_time_us(lambda: ops.accumulate_packed(stage, ops.flat_geometry(n, d), bmu, wt, kk, packed=packed), reps=5)
        print(f"  {n} patches: nchw {t_a:.1f}  staged {t_b:.1f}")
    w = torch.randn(kk, d, device=dev)
    print("norms us", graph_time_us(lambda: ops.prepare_codebook(w)))
    m, v, g = torch.zeros_like(w), torch.zeros_like(w), torch.randn_like(w)
    td = torch.zeros(2, dtype=torch.int64, device=dev)
    tail = torch.tensor([1.0, 0.0, 256.0, 0.0], device=dev)
    lo = torch.empty(1, dtype=torch.float64, device=dev)
    print("adam_dp us", graph_time_us(lambda: ops.adam_step_dp(w, m, v, g, d, 1e-4, td, tail, loss_out=lo)))
