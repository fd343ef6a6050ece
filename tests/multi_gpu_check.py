"""N-GPU check of the multi-GPU paths over NCCL (launch with torch.distributed.run):
  * DataParallelSom on N ranks == SomTrainer on one GPU over the full batch (replicas bit-identical)
  * unit-sharded search (sharded_bmu) == unsharded search, sharded histogram concatenates to the full one
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import somcb  # noqa: E402
from somcb import ops  # noqa: E402
from somcb.distributed import sharded_histogram  # noqa: E402
from oracle.step_oracle import synthetic_fmaps, trained_like_codebook  # noqa: E402


def make_cb(w, pd, dev):
    cb = somcb.Codebook(patch_dim=pd, image_dim=(32, 32), image_channel=4, num_embeddings=w.shape[0],
                        init_neighbour_range=w.shape[0] // 2)
    with torch.no_grad():
        cb.codebook.weight.copy_(w)
    return cb.to(dev)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True

    # ---- data parallel ---------------------------------------------------------------------------
    pd, k = (4, 4), 4096
    w0 = trained_like_codebook(k, pd, 7)
    batch = 64 * world
    dp_cb = make_cb(w0, pd, dev)
    dp = somcb.DataParallelSom(dp_cb, lr=1e-4, neighbourhood_step=3)           # tail="auto": peer memory when available
    dp.broadcast_weights(0)
    dp_n_cb = make_cb(w0, pd, dev)
    dp_n = somcb.DataParallelSom(dp_n_cb, lr=1e-4, neighbourhood_step=3, tail="nccl")
    if rank == 0:
        print(f"[dp] tail of the default trainer: {dp.tail}")
    ref_cb = make_cb(w0, pd, dev)
    ref = somcb.SomTrainer(ref_cb, lr=1e-4, neighbourhood_step=3)
    dp_g_cb = make_cb(w0, pd, dev)                          # the same steps through the captured CUDA graph
    dp_g = somcb.DataParallelSom(dp_g_cb, lr=1e-4, neighbourhood_step=3, use_cuda_graph=True)
    for step in range(6):
        # steps 4 and 5 use a batch that does not divide by the world size (ragged shares)
        n_f = batch if step < 4 else batch + 1 + step
        x = synthetic_fmaps(n_f, 500 + step).to(dev)
        l_dp = dp.step(somcb.split_batch(x, world, rank).contiguous())
        l_g = dp_g.step(somcb.split_batch(x, world, rank).contiguous())
        l_n = dp_n.step(somcb.split_batch(x, world, rank).contiguous())
        l_ref = ref.step(x)
        rel_n = float((dp_n_cb.codebook.weight.data - ref_cb.codebook.weight.data).norm() / ref_cb.codebook.weight.data.norm())
        ok &= rel_n <= 1e-6 and abs(float(l_n) - float(l_ref)) <= 1e-6 * abs(float(l_ref))
        same_g = torch.equal(dp_g_cb.codebook.weight.data, dp_cb.codebook.weight.data) and float(l_g) == float(l_dp)
        if rank == 0 and not same_g:
            print(f"[dp] step {step}: graph-replayed DP step differs from the eager DP step")
        ok &= same_g
        rel_l = abs(float(l_dp) - float(l_ref)) / abs(float(l_ref))
        rel_w = float((dp_cb.codebook.weight - ref_cb.codebook.weight).norm() / ref_cb.codebook.weight.norm())
        gathered = [torch.empty_like(dp_cb.codebook.weight.data) for _ in range(world)]
        dist.all_gather(gathered, dp_cb.codebook.weight.data)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        if rank == 0:
            print(f"[dp] step {step}: loss rel {rel_l:.2e}, weights vs 1-GPU rel {rel_w:.2e}, replicas identical {same}, "
                  f"range {dp_cb.neighbourhood_range}")
        ok &= rel_l <= 1e-6 and rel_w <= 1e-6 and same and dp_cb.neighbourhood_range == ref_cb.neighbourhood_range

    # ---- in-switch all-reduce (multimem.ld_reduce / multimem.st) against the fp64 sum ----------------------
    if dp.peer is not None:
        pm = dp.peer
        n = 1 << 20
        buf, mc, _ = pm.alloc(n)
        g = torch.Generator(device=dev).manual_seed(900 + rank)
        buf.copy_(torch.randn(n, generator=g, device=dev) * (10.0 ** (rank % 3)))
        parts = [torch.empty(n, device=dev) for _ in range(world)]
        dist.all_gather(parts, buf)
        want = torch.stack(parts).double().sum(dim=0)
        torch.cuda.synchronize()
        dist.barrier()
        ops.peer_allreduce(mc, n, pm.rank, pm.world, pm.signal_ptrs, 3, dev)
        torch.cuda.synchronize()
        err = float((buf.double() - want).abs().max() / want.abs().max())
        chk = [torch.empty(n, device=dev) for _ in range(world)]
        dist.all_gather(chk, buf)
        same = all(torch.equal(chk[0], c) for c in chk)
        if rank == 0:
            print(f"[peer] in-switch all-reduce of 2^20 floats: max error {err:.2e} of the largest sum, "
                  f"identical on every rank: {same}")
        ok &= err <= 1e-6 and same

        # ---- reduce-scatter fused into the filter's read against reduce_rows + filter and the fp64 sum -------------
        kk, dd, rng_f = 4096, 64, 2048
        packed, mc_p, peers_p = pm.alloc(kk * dd + 4)
        g = torch.Generator(device=dev).manual_seed(40 + rank)
        packed[:kk * dd].copy_(torch.randn(kk * dd, generator=g, device=dev))
        packed[kk * dd:].copy_(torch.tensor([1.5 + rank, 0.25, 0.0, 100.0 + rank], device=dev))
        allp = [torch.empty(kk * dd + 4, device=dev) for _ in range(world)]
        dist.all_gather(allp, packed)
        total = torch.stack([a[:kk * dd] for a in allp]).double().sum(dim=0).view(kk, dd)
        lo_u, hi_u = somcb.shard_bounds(kk, world, rank)
        hw = ops.filter_half_width(kk, rng_f)
        g0, g1 = max(0, lo_u - hw), min(kk, hi_u + hw)
        max_rows = kk
        scratch = torch.empty(max_rows, dd, device=dev)
        fused = torch.empty(g1 - g0, dd, device=dev)
        tail_a, tail_b = torch.empty(4, device=dev), torch.empty(4, device=dev)
        rows = torch.empty(g1 - g0, dd, device=dev)
        torch.cuda.synchronize()
        dist.barrier()
        ops.peer_reduce_filter_rows(mc_p, peers_p, kk, dd, g0, g1, max_rows, rng_f, scratch, fused, tail_a, pm.rank,
                                    pm.world, pm.signal_ptrs, 3)
        ops.peer_reduce_rows(mc_p, peers_p, kk, dd, g0, g1, max_rows, rows, tail_b, pm.rank, pm.world, pm.signal_ptrs, 3)
        two_step = ops.neighbourhood_filter(rows, rng_f)
        want_f = ops.neighbourhood_filter(total[g0:g1].float().contiguous(), rng_f).double()
        torch.cuda.synchronize()
        same_f = bool(torch.equal(fused, two_step)) and bool(torch.equal(tail_a, tail_b))
        err_f = float((fused.double() - want_f).norm() / want_f.norm())
        want_tail = sum(1.75 + r for r in range(world))
        tail_ok = abs(float(tail_a[0]) + float(tail_a[1]) - want_tail) <= 1e-6 * want_tail and \
            float(tail_a[2]) * 4096 + float(tail_a[3]) == sum(100.0 + r for r in range(world))
        if rank == 0:
            print(f"[peer] filter with the in-switch reduce fused into its read: identical to reduce_rows + filter "
                  f"{same_f}, rel error vs the fp64 sum {err_f:.2e}, exact tail {tail_ok}")
        ok &= same_f and err_f <= 1e-6 and tail_ok
        dist.barrier()

    # ---- unit-sharded search + histogram ---------------------------------------------------------
    pd, k = (8, 8), 8192
    w = trained_like_codebook(k, pd, 3).to(dev)
    x = synthetic_fmaps(512, 77).to(dev)
    geom = ops.geometry(x.shape, pd)
    full = ops.bmu(x, geom, w)
    lo, hi = somcb.shard_bounds(k, world, rank)
    got = somcb.sharded_bmu(x, geom, w[lo:hi].contiguous(), lo)
    eq = bool(torch.equal(got, full))
    mine = sharded_histogram(got, lo, hi)
    parts = [torch.empty(somcb.shard_bounds(k, world, r)[1] - somcb.shard_bounds(k, world, r)[0],
                         dtype=torch.int64, device=dev) for r in range(world)]
    dist.all_gather(parts, mine)
    hist_ok = bool(torch.equal(torch.cat(parts), ops.histogram(full, k)))
    if rank == 0:
        print(f"[shard] sharded_bmu == unsharded: {eq}; sharded histogram == full: {hist_ok}")
    ok &= eq and hist_ok

    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag) else "FAIL", flush=True)
    # captured CUDA graphs hold NCCL kernels and destroy_process_group() can block behind them: release the graphs
    # and leave without tearing the communicator down
    dp_g._graphs.clear()
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
