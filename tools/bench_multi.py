"""Multi-GPU timings of the two shardings at BASELINE sizes (launch with torch.distributed.run, one rank per GPU):

  C4  data-parallel SOM: 2^20 patches per GPU per step (weak scaling), P=4 D=64 K=16384, one NCCL all-reduce of
      the packed accumulators per step;
  C5  unit-sharded search: 2^20 patches (replicated), D=256, 32 768 units per GPU (K = 32 768 * world),
      all-gather of (distance, index) candidates + merge, then the sharded hit histogram.

Device timing with CUDA events, barrier + synchronize on both sides, MAX over ranks; rank 0 prints one JSON line.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import somcb  # noqa: E402
from somcb import ops  # noqa: E402
from somcb.distributed import sharded_bmu, sharded_histogram  # noqa: E402


def fmaps(n, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.empty(n, 4, 32, 32, device=dev)
    for lo in range(0, n, 8192):
        hi = min(n, lo + 8192)
        x[lo:hi] = torch.tanh(torch.randn(hi - lo, 4, 32, 32, generator=g, device=dev))
    return x


def timed(fn, reps, warm, dev, world):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {"n_gpus": world}

    # ---- C4: data-parallel SOM steps ----------------------------------------------------------------
    k, pd = 16384, (4, 4)
    x = fmaps(16384, 1000 + rank, dev)
    pool = fmaps(256, 7, dev)
    w0 = somcb.patchify(pool, pd).reshape(-1, 64)[:k].contiguous()
    cb = somcb.Codebook(patch_dim=pd, image_dim=(32, 32), image_channel=4, num_embeddings=k,
                        init_neighbour_range=k // 2).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(w0)
    tr = somcb.DataParallelSom(cb, lr=1e-4, neighbourhood_step=200) if world > 1 else \
        somcb.SomTrainer(cb, lr=1e-4, neighbourhood_step=200)
    ms = timed(lambda: tr.step(x), 10, 3, dev, world)
    out["c4_dp_step"] = {"ms_per_step": ms, "patches_per_s": world * 16384 * 64 / ms * 1e3,
                         "patches_per_gpu_per_step": 16384 * 64, "allreduce_bytes": 4 * k * 64 + 8}
    del x, tr, cb
    torch.cuda.empty_cache()

    # ---- C5: unit-sharded search + histogram --------------------------------------------------------
    pd, k_local = (8, 8), 32768
    x = fmaps(65536, 123, dev)                              # same patches on every rank
    gw = torch.Generator(device=dev).manual_seed(500 + rank)
    w_shard = torch.tanh(torch.randn(k_local, 256, generator=gw, device=dev))
    geom = ops.geometry(x.shape, pd)
    cn = ops.prepare_codebook(w_shard)
    lo = rank * k_local
    state = {}

    def search():
        idx = sharded_bmu(x, geom, w_shard, lo, c_norm2=cn)
        state["counts"] = sharded_histogram(idx, lo, lo + k_local)
        state["idx"] = idx

    ms = timed(search, 3, 2, dev, world)
    n_p = ops.n_patches_of(geom)
    tot = state["counts"].sum().to(torch.int64)
    if world > 1:
        dist.all_reduce(tot)
    assert int(tot) == n_p, f"sharded histogram sums to {int(tot)}, expected {n_p}"
    out["c5_sharded_search"] = {"ms": ms, "patches_per_s": n_p / ms * 1e3, "units_total": k_local * world,
                                "patches": n_p, "unit_patch_pairs_per_s": n_p * k_local * world / ms * 1e3}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
