"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: data-parallel packing + all-reduce,
unit-sharded candidate merge, sharded histogram.  Kernels are replaced by the CPU test double
(tests/_oracle_ops.py); on the GPU box the same code paths run with the CUDA ops over NCCL."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, out_dir):
    for p in (HERE, os.path.dirname(HERE),
              os.path.join(os.path.dirname(HERE), "quantized-autoregression-image-generator_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world, out_dir)
    finally:
        dist.destroy_process_group()


def _run(fn_name, tmp_path, world=2):
    mp.spawn(_worker, args=(world, _free_port(), fn_name, str(tmp_path)), nprocs=world, join=True)


def _make_cb(rec):
    import somcb
    cb = somcb.Codebook(patch_dim=rec["patch_dim"], image_dim=rec["image_dim"],
                        image_channel=rec["channels"], num_embeddings=rec["weight"].shape[0],
                        init_neighbour_range=rec["neighbourhood_range"])
    with torch.no_grad():
        cb.codebook.weight.copy_(rec["weight"])
    return cb


def _dp_body(rank, world, out_dir):
    import somcb
    import _oracle_ops
    from _helpers import load_case
    rec = load_case("c1_trained")
    cb = _make_cb(rec)
    tr = somcb.DataParallelSom(cb, lr=1e-4, neighbourhood_step=10 ** 9, ops=_oracle_ops)
    assert tr.world_size == world and tr.reduce_fn is not None
    x_local = somcb.split_batch(rec["x"], world, rank).contiguous()
    losses = [tr.step(x_local) for _ in range(3)]
    torch.save({"w": cb.codebook.weight.detach().clone(), "loss": torch.stack(losses)},
               os.path.join(out_dir, f"dp_{rank}.pt"))


def test_data_parallel_matches_single_process(tmp_path):
    import somcb
    import _oracle_ops
    from _helpers import assert_close_norm, load_case
    _run("_dp_body", tmp_path)
    r0 = torch.load(tmp_path / "dp_0.pt")
    r1 = torch.load(tmp_path / "dp_1.pt")
    assert torch.equal(r0["w"], r1["w"]), "replicas diverged"
    assert torch.equal(r0["loss"], r1["loss"])
    rec = load_case("c1_trained")
    cb = _make_cb(rec)
    tr = somcb.SomTrainer(cb, lr=1e-4, neighbourhood_step=10 ** 9, ops=_oracle_ops)
    losses = torch.stack([tr.step(rec["x"]) for _ in range(3)])
    assert_close_norm(r0["w"], cb.codebook.weight.detach(), 1e-6, "dp weights vs single")
    assert_close_norm(r0["loss"], losses, 1e-6, "dp loss vs single")


def _dp_ragged_body(rank, world, out_dir):
    """7 feature maps over 2 ranks (4 + 3), then a step where rank 1's share is EMPTY (1 fmap over 2 ranks)."""
    import somcb
    import _oracle_ops
    from _helpers import load_case
    rec = load_case("c1_trained")
    cb = _make_cb(rec)
    tr = somcb.DataParallelSom(cb, lr=1e-4, neighbourhood_step=10 ** 9, ops=_oracle_ops)
    losses = [tr.step(somcb.split_batch(rec["x"][:7], world, rank).contiguous()),
              tr.step(somcb.split_batch(rec["x"][7:8], world, rank).contiguous())]
    torch.save({"w": cb.codebook.weight.detach().clone(), "loss": torch.stack(losses)},
               os.path.join(out_dir, f"dpr_{rank}.pt"))


def test_data_parallel_ragged_and_empty_shares(tmp_path):
    import somcb
    import _oracle_ops
    from _helpers import assert_close_norm, load_case
    _run("_dp_ragged_body", tmp_path)
    r0 = torch.load(tmp_path / "dpr_0.pt")
    r1 = torch.load(tmp_path / "dpr_1.pt")
    assert torch.equal(r0["w"], r1["w"]) and torch.equal(r0["loss"], r1["loss"])
    rec = load_case("c1_trained")
    cb = _make_cb(rec)
    tr = somcb.SomTrainer(cb, lr=1e-4, neighbourhood_step=10 ** 9, ops=_oracle_ops)
    losses = torch.stack([tr.step(rec["x"][:7].contiguous()), tr.step(rec["x"][7:8].contiguous())])
    assert_close_norm(r0["w"], cb.codebook.weight.detach(), 1e-6, "ragged dp weights vs single")
    assert_close_norm(r0["loss"], losses, 1e-6, "ragged dp loss vs single")


def _shard_body(rank, world, out_dir):
    import somcb
    import _oracle_ops
    from somcb import ops
    from somcb.distributed import sharded_histogram
    from _helpers import load_case, load_golden
    res = {}
    for name, rec in (("c1_trained", load_case("c1_trained")), ("ties", load_golden("case_ties.pt"))):
        k = rec["weight"].shape[0]
        lo, hi = somcb.shard_bounds(k, world, rank)
        geom = ops.geometry(rec["x"].shape, rec["patch_dim"])
        idx = somcb.sharded_bmu(rec["x"], geom, rec["weight"][lo:hi].contiguous(), lo, ops=_oracle_ops)
        res[name] = idx
        res[name + "_hist"] = sharded_histogram(idx, lo, hi, ops=_oracle_ops)
    torch.save(res, os.path.join(out_dir, f"shard_{rank}.pt"))


def test_unit_sharded_search_matches_unsharded(tmp_path):
    from _helpers import assert_bmu_parity, flat_patches, load_case, load_golden
    _run("_shard_body", tmp_path)
    r0 = torch.load(tmp_path / "shard_0.pt")
    r1 = torch.load(tmp_path / "shard_1.pt")
    for name, rec in (("c1_trained", load_case("c1_trained")), ("ties", load_golden("case_ties.pt"))):
        assert torch.equal(r0[name], r1[name])
        flat = flat_patches(rec["x"], rec["patch_dim"])
        assert_bmu_parity(r0[name], rec["bmu"], flat, rec["weight"])
        full = torch.bincount(r0[name], minlength=rec["weight"].shape[0])
        assert torch.equal(torch.cat([r0[name + "_hist"], r1[name + "_hist"]]), full)
    # duplicated rows live in the second shard: the merge must still return the first copy
    assert torch.equal(r0["ties"], load_golden("case_ties.pt")["bmu"])


def _counts_body(rank, world, out_dir):
    from somcb.distributed import allreduce_counts
    c = torch.arange(5, dtype=torch.int64) * (rank + 1)
    allreduce_counts(c)
    torch.save(c, os.path.join(out_dir, f"counts_{rank}.pt"))


def test_patch_sharded_histogram_allreduce(tmp_path):
    _run("_counts_body", tmp_path)
    want = torch.arange(5, dtype=torch.int64) * 3
    assert torch.equal(torch.load(tmp_path / "counts_0.pt"), want)
    assert torch.equal(torch.load(tmp_path / "counts_1.pt"), want)
