"""GPU tests of the histogram / prune path, the C-ABI error behaviour, the host-buffer pipeline
and a literal replay of the reference scripts' call sequences on the drop-in module."""
import ctypes

import pytest
import torch

import somcb
from somcb import _lib, ops
from oracle.step_oracle import synthetic_fmaps, trained_like_codebook
from _helpers import assert_bmu_parity, flat_patches, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("k,n", [(16, 1000), (4096, 1 << 20), (12288, 50000), (262144, 1 << 20)])
def test_histogram_matches_bincount(k, n):
    g = torch.Generator().manual_seed(k)
    idx = torch.randint(0, k, (n,), generator=g)
    idx[::7] = idx[0]                                    # a heavy hitter
    counts = ops.histogram(idx.to(DEV), k)
    assert torch.equal(counts.cpu(), torch.bincount(idx, minlength=k))
    ops.histogram(idx.to(DEV), k, counts)                # accumulates
    assert torch.equal(counts.cpu(), 2 * torch.bincount(idx, minlength=k))
    bad = idx.clone()
    bad[:10] = -1
    bad[10:20] = k
    c2 = ops.histogram(bad.to(DEV), k)
    assert int(c2.sum()) == n - 20                       # out-of-range indices are ignored


def test_prune_matches_reference_golden():
    rec = load_golden("prune_case.pt")
    cb = somcb.Codebook(patch_dim=rec["patch_dim"], image_dim=rec["image_dim"],
                        image_channel=rec["channels"], num_embeddings=rec["weight"].shape[0],
                        init_neighbour_range=128)
    with torch.no_grad():
        cb.codebook.weight.copy_(rec["weight"])
    cb = cb.to(DEV).eval()
    batches = [synthetic_fmaps(rec["batch"], s).to(DEV) for s in rec["seeds"]]
    counts = somcb.bmu_histogram(cb, batches)
    assert torch.equal(counts.cpu(), rec["counts"])
    new_cb, keep, ck = somcb.prune_codebook(cb, counts, rec["threshold"], image_channel=rec["channels"],
                                            global_steps=17)
    assert torch.equal(keep.cpu(), rec["good"])
    assert torch.equal(new_cb.codebook.weight.detach().cpu(), rec["pruned_state_dict"]["codebook.weight"])
    assert ck["num_embeddings"] == rec["good"].numel() and ck["global_steps"] == 17
    assert list(ck["checkpoint"]) == ["codebook.weight"]
    # the literal reference loop (prune_codebook.py:138-142) on the drop-in gives the same counts
    total = {i: 0 for i in range(cb.num_embeddings)}
    for fm in batches:
        for j in cb.get_patches_bmu(fm).tolist():
            total[j] += 1
    assert [total[i] for i in range(cb.num_embeddings)] == rec["counts"].tolist()


def test_merge_candidates_rule():
    g = torch.Generator().manual_seed(1)
    r, n = 5, 10000
    rd = torch.randn(r, n, generator=g).round(decimals=1)         # many exact ties
    idx = torch.stack([torch.randint(0, 1000, (n,), generator=g) + 1000 * i for i in range(r)])
    got_i, got_rd = ops.merge_candidates(rd.to(DEV), idx.to(DEV))
    best = rd.min(dim=0).values
    masked = torch.where(rd == best, idx, torch.full_like(idx, 1 << 60))
    assert torch.equal(got_i.cpu(), masked.min(dim=0).values)
    assert torch.equal(got_rd.cpu(), best)


def test_cabi_error_codes_and_messages():
    lib = _lib.load()
    x = torch.zeros(2, 4, 32, 32, device=DEV)
    w = torch.zeros(64, 64, device=DEV)
    cn = torch.zeros(64, device=DEV)
    out = torch.zeros(128, dtype=torch.int64, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.som_bmu_nchw_f32(None, 2, 4, 32, 32, 4, 4, w.data_ptr(), cn.data_ptr(), 64, 0,
                              out.data_ptr(), None, None, 0, 0, st)
    assert rc == -1 and b"null pointer" in lib.som_last_error()
    rc = lib.som_bmu_nchw_f32(x.data_ptr(), 2, 4, 32, 32, 5, 4, w.data_ptr(), cn.data_ptr(), 64, 0,
                              out.data_ptr(), None, None, 0, 0, st)
    assert rc == -2 and b"not divisible" in lib.som_last_error()
    bmu = torch.zeros(128, dtype=torch.int64, device=DEV)
    rbar = torch.zeros(64, 64, device=DEV)
    ws = torch.zeros(256, dtype=torch.uint8, device=DEV)
    rc = lib.som_accumulate_nchw_f32(x.data_ptr(), 2, 4, 32, 32, 4, 4, bmu.data_ptr(), None, 64,
                                     rbar.data_ptr(), None, None, ws.data_ptr(), 256, st)
    assert rc == -3 and b"workspace" in lib.som_last_error()
    rc = lib.som_filter_f32(w.data_ptr(), w.data_ptr(), 64, 64, ctypes.c_double(4.0), ctypes.c_float(1.0), st)
    assert rc == -1
    with pytest.raises(_lib.SomError):
        _lib.check("som_filter_f32", rc)
    sm, major, _ = ops.device_info()
    assert sm > 0 and major >= 10, "these kernels are built for sm_100a only"


def test_cabi_raw_pointer_call_roundtrip():
    """Call the library exactly as a foreign host would: raw device pointers and a stream."""
    lib = _lib.load()
    rec = load_golden("case_ties.pt")
    x = rec["x"].to(DEV)
    w = rec["weight"].to(DEV)
    k, d = w.shape
    st = torch.cuda.current_stream().cuda_stream
    cn = torch.empty(k, device=DEV)
    assert lib.som_prepare_codebook_f32(w.data_ptr(), k, d, cn.data_ptr(), st) == 0
    assert torch.allclose(cn.cpu(), (rec["weight"] ** 2).sum(1), rtol=1e-6)
    n_p = 4 * 64
    out = torch.empty(n_p, dtype=torch.int64, device=DEV)
    nbytes = lib.som_bmu_workspace_bytes(n_p, d, k, 0)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=DEV)
    rc = lib.som_bmu_nchw_f32(x.data_ptr(), 4, 4, 32, 32, 4, 4, w.data_ptr(), cn.data_ptr(), k, 0,
                              out.data_ptr(), None, ws.data_ptr(), nbytes, 0, st)
    assert rc == 0, lib.som_last_error()
    assert_bmu_parity(out, rec["bmu"], flat_patches(rec["x"], rec["patch_dim"]), rec["weight"])


def test_host_tokenizer_matches_direct_call():
    pd, k = (2, 2), 1024
    w = trained_like_codebook(k, pd, 7)
    cb = somcb.Codebook(patch_dim=pd, image_dim=(32, 32), image_channel=4, num_embeddings=k,
                        init_neighbour_range=k // 2)
    with torch.no_grad():
        cb.codebook.weight.copy_(w)
    cb = cb.to(DEV)
    host = synthetic_fmaps(1000, 5).pin_memory()
    tok = somcb.HostTokenizer(cb, chunk_fmaps=192, depth=3)     # ragged last chunk, slot reuse
    want = cb.get_patches_bmu(host.to(DEV), reshape=True).cpu()
    torch.cuda.synchronize()
    got = tok.tokenize(host)                     # synchronous by default: valid on return, no device sync here
    assert got.shape == (1000, 256) and torch.equal(got, want)
    got2 = tok.tokenize(host)
    assert torch.equal(got2, want)
    out = torch.empty(1000, 256, dtype=torch.int64, pin_memory=True)
    got3 = tok.tokenize(host, out, sync=False)   # asynchronous form: wait on the tokenizer's event
    assert tok.done is not None
    tok.done.synchronize()
    assert got3 is out and torch.equal(out, want)


def test_reference_tokenisation_call_sequence():
    """train_quantized_transformer.py:411-421: two codebooks with different patch sizes over the
    same fmap, called in grad mode, reshape=True."""
    x = synthetic_fmaps(8, 9).to(DEV)
    outs = []
    for pd, k in (((4, 4), 512), ((2, 2), 2048)):
        cb = somcb.Codebook(patch_dim=pd, image_dim=(32, 32), image_channel=4, num_embeddings=k,
                            init_neighbour_range=k // 2)
        with torch.no_grad():
            cb.codebook.weight.copy_(trained_like_codebook(k, pd, 3))
        cb = cb.to(DEV)
        cb.eval()
        idx = cb.get_patches_bmu(x, reshape=True)
        outs.append(idx)
        img = cb.get_quantized_image(idx)                 # decode path (generate_images.py:225)
        assert img.shape == x.shape
        assert torch.equal(cb.get_patches_bmu(img, reshape=True), idx), "codec round trip"
    assert outs[0].shape == (8, 64) and outs[1].shape == (8, 256)


@pytest.mark.parametrize("base_model", [True, False])
def test_tokenize_pair_matches_reference_call_sequence(base_model):
    """train_quantized_transformer.py:411-455 on the drop-in modules: both BMU searches and the
    fused token assembly against the CPU restatement, bit-exact."""
    from oracle import tokenize_pair_oracle
    from oracle.step_oracle import make_oracle_codebook
    x = synthetic_fmaps(24, 77)
    specs = [((8, 8), 300), ((2, 2), 1000)]               # low-res / high-res codebooks
    ocs, gcs = [], []
    for pd, k in specs:
        w = trained_like_codebook(k, pd, 21 + k)
        ocs.append(make_oracle_codebook(w, pd, (32, 32), 4, k // 2))
        cb = somcb.Codebook(patch_dim=pd, image_dim=(32, 32), image_channel=4, num_embeddings=k,
                            init_neighbour_range=k // 2)
        with torch.no_grad():
            cb.codebook.weight.copy_(w)
        gcs.append(cb.to(DEV).eval())
    ref_in, ref_tgt, ref_lr = tokenize_pair_oracle(ocs[0], ocs[1], x, base_model)
    got_in, got_tgt, got_lr = somcb.tokenize_pair(gcs[0], gcs[1], x.to(DEV), base_model)
    assert got_in.dtype == torch.int64 and got_in.shape == ref_in.shape
    assert torch.equal(got_in.cpu(), ref_in) and torch.equal(got_tgt.cpu(), ref_tgt)
    if base_model:
        assert got_lr is None and ref_lr is None
        assert int(got_in[:, :16].max()) < 300 and int(got_in[:, 16:].min()) >= 300
    else:
        assert torch.equal(got_lr.cpu(), ref_lr) and int(got_in[:, 0].min()) == 1000
    assert int(got_tgt[:, -1].min()) == 1000 and int(got_tgt[:, -1].max()) == 1000


def test_shard_reader_feeds_host_tokenizer(tmp_path):
    """Packed shards -> pinned batches -> HostTokenizer == one resident-batch BMU call."""
    from somcb import fmap_shards as fs
    x = synthetic_fmaps(300, 31)
    fs.write_shard(str(tmp_path / "a.shard"), x[:180].numpy())
    fs.write_shard(str(tmp_path / "b.shard"), x[180:].numpy())
    reader = fs.ShardReader([str(tmp_path / "a.shard"), str(tmp_path / "b.shard")], batch_fmaps=64)
    pd, k = (2, 2), 777
    cb = somcb.Codebook(patch_dim=pd, image_dim=(32, 32), image_channel=4, num_embeddings=k,
                        init_neighbour_range=k // 2)
    with torch.no_grad():
        cb.codebook.weight.copy_(trained_like_codebook(k, pd, 5))
    cb = cb.to(DEV).eval()
    tok = somcb.HostTokenizer(cb, chunk_fmaps=48, depth=3)
    ref = cb.get_patches_bmu(x.to(DEV), reshape=True).cpu()
    counts = None
    for lo, hi, batch in reader.batches():
        idx = tok.tokenize(batch)
        assert torch.equal(idx, ref[lo:hi])
        counts = ops.histogram(idx.to(DEV).reshape(-1), k, counts)
    assert torch.equal(counts.cpu(), torch.bincount(ref.reshape(-1), minlength=k))
    # asynchronous consumer: the reader must not refill a pinned staging buffer that a copy is still reading
    outs, evs = [], []
    for lo, hi, batch in reader.batches():
        dst = torch.empty(batch.shape, device=DEV)
        dst.copy_(batch, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        reader.mark_in_flight(ev)
        outs.append((lo, hi, dst))
    torch.cuda.synchronize()
    for lo, hi, dst in outs:
        assert torch.equal(dst.cpu(), x[lo:hi])


def test_host_trainer_matches_device_trainer():
    """somcb.HostTrainer (pinned host batches, H2D on a copy stream, step graph on the staging buffers, loss through a
    pinned scalar) == the same steps on device-resident batches, bit for bit."""
    pd, k = (4, 4), 2048
    w0 = trained_like_codebook(k, pd, 7)
    trainers = []
    for _ in range(2):
        cb = somcb.Codebook(patch_dim=pd, image_dim=(32, 32), image_channel=4, num_embeddings=k,
                            init_neighbour_range=k // 2)
        with torch.no_grad():
            cb.codebook.weight.copy_(w0)
        trainers.append(somcb.SomTrainer(cb.to(DEV), lr=1e-4, neighbourhood_step=4, use_cuda_graph="alias"))
    ht = somcb.HostTrainer(trainers[0], depth=2)
    hosts = [synthetic_fmaps(96, 40 + i).pin_memory() for i in range(7)]
    pend = [ht.step(h) for h in hosts[:2]]
    losses = [p.item() for p in pend]
    for h in hosts[2:]:
        losses.append(ht.step(h).item())
    want = [float(trainers[1].step(h.to(DEV))) for h in hosts]
    assert losses == want
    assert torch.equal(trainers[0].cb.codebook.weight.data, trainers[1].cb.codebook.weight.data)
    assert trainers[0].cb.neighbourhood_range == trainers[1].cb.neighbourhood_range == k // 2 - 1


def test_tokenize_pair_matches_reference_golden():
    """GPU tokeniser against the outputs of the reference script's own lines (tests/golden/tokens_case.pt)."""
    rec = load_golden("tokens_case.pt")
    cbs = []
    for wkey, pkey in (("lr_weight", "lr_patch"), ("hr_weight", "hr_patch")):
        w = rec[wkey]
        cb = somcb.Codebook(patch_dim=rec[pkey], image_dim=rec["image_dim"], image_channel=rec["channels"],
                            num_embeddings=w.shape[0], init_neighbour_range=w.shape[0] // 2)
        with torch.no_grad():
            cb.codebook.weight.copy_(w)
        cbs.append(cb.to(DEV).eval())
    for base, tag in ((True, "base"), (False, "cond")):
        hi, ht, li = somcb.tokenize_pair(cbs[0], cbs[1], rec["x"].to(DEV), base)
        assert torch.equal(hi.cpu(), rec[f"hr_input_{tag}"]) and torch.equal(ht.cpu(), rec[f"hr_target_{tag}"])
        if rec[f"lr_input_{tag}"] is None:
            assert li is None
        else:
            assert torch.equal(li.cpu(), rec[f"lr_input_{tag}"])


@pytest.mark.parametrize("shape", [
    # (fmaps, C, H, W, pH, pW, K)            tensor-core mode exercised
    (33, 4, 32, 32, 2, 2, 1000),             # config S, ragged patch / unit tiles
    (21, 4, 32, 32, 4, 4, 700),              # resident-A
    (9, 4, 32, 32, 8, 8, 300),               # streamed (TMA-fed A, pre-pass workspace)
    (50, 4, 32, 32, 32, 32, 515),            # split-K (partial-distance workspace)
    (48, 3, 12, 18, 2, 3, 333),              # D = 18: scalar builder path
])
def test_bmu_writes_stay_inside_their_buffers(shape):
    """No sanitizer on this pool: call the C-ABI with every output and the workspace embedded in larger
    sentinel-filled buffers and check the guard bands afterwards (both variants)."""
    n, c, h, w, ph, pw, k = shape
    lib = _lib.load()
    g = torch.Generator().manual_seed(k)
    x = torch.tanh(torch.randn(n, c, h, w, generator=g)).to(DEV)
    d = c * ph * pw
    wgt = torch.tanh(torch.randn(k, d, generator=g)).to(DEV)
    cn = ops.prepare_codebook(wgt)
    n_p = n * (h // ph) * (w // pw)
    guard = 4096
    st = torch.cuda.current_stream().cuda_stream
    for variant in (ops.SOM_BMU_FFMA, ops.SOM_BMU_TC3X):
        ws_bytes = lib.som_bmu_workspace_bytes(n_p, d, k, variant)
        ws_big = torch.full((ws_bytes + 2 * guard + 256,), 0x5A, dtype=torch.uint8, device=DEV)
        base = ws_big.data_ptr() + guard
        ws_ptr = (base + 255) // 256 * 256
        off = ws_ptr - ws_big.data_ptr()
        idx_big = torch.full((n_p + 2 * 512,), -777, dtype=torch.int64, device=DEV)
        rd_big = torch.full((n_p + 2 * 512,), -777.0, dtype=torch.float32, device=DEV)
        rc = lib.som_bmu_nchw_f32(x.data_ptr(), n, c, h, w, ph, pw, wgt.data_ptr(), cn.data_ptr(), k, 0,
                                  idx_big.data_ptr() + 512 * 8, rd_big.data_ptr() + 512 * 4,
                                  ws_ptr if ws_bytes else None, ws_bytes, variant, st)
        assert rc == 0, lib.som_last_error()
        torch.cuda.synchronize()
        assert bool((idx_big[:512] == -777).all()) and bool((idx_big[512 + n_p:] == -777).all())
        assert bool((rd_big[:512] == -777.0).all()) and bool((rd_big[512 + n_p:] == -777.0).all())
        assert bool((ws_big[:off] == 0x5A).all()) and bool((ws_big[off + ws_bytes:] == 0x5A).all())
        got = idx_big[512:512 + n_p]
        assert int(got.min()) >= 0 and int(got.max()) < k
        ref = ops.bmu(x, ops.geometry(x.shape, (ph, pw)), wgt, cn, variant=ops.SOM_BMU_FFMA)
        assert int((got != ref).sum()) <= max(1, n_p // 2000)          # near-ties only


def test_flat_and_backward_entry_points_match_their_general_forms():
    """som_bmu_flat_f32 == BMU over patchified rows; som_backward_nchw_f32 == accumulate with Wt = NULL."""
    lib = _lib.load()
    x = synthetic_fmaps(12, 5).to(DEV)
    pd, k = (4, 4), 600
    wgt = trained_like_codebook(k, pd, 3).to(DEV)
    cn = ops.prepare_codebook(wgt)
    geom = ops.geometry(x.shape, pd)
    ref = ops.bmu(x, geom, wgt, cn, variant=ops.SOM_BMU_FFMA)
    flat = somcb.patchify(x, pd).reshape(-1, 64).contiguous()
    n_p = flat.shape[0]
    st = torch.cuda.current_stream().cuda_stream
    out = torch.empty(n_p, dtype=torch.int64, device=DEV)
    nb = lib.som_bmu_workspace_bytes(n_p, 64, k, ops.SOM_BMU_FFMA)
    ws = torch.empty(max(nb, 1), dtype=torch.uint8, device=DEV)
    rc = lib.som_bmu_flat_f32(flat.data_ptr(), n_p, 64, wgt.data_ptr(), cn.data_ptr(), k, 0, out.data_ptr(), None,
                              ws.data_ptr(), nb, ops.SOM_BMU_FFMA, st)
    assert rc == 0 and torch.equal(out, ref)
    g_out = torch.randn_like(x)
    want, _, _ = ops.accumulate(g_out, geom, ref, None, k)
    rbar = torch.empty(k, 64, device=DEV)
    nb2 = lib.som_accumulate_workspace_bytes(n_p, 64, k)
    ws2 = torch.empty(nb2, dtype=torch.uint8, device=DEV)
    rc = lib.som_backward_nchw_f32(g_out.data_ptr(), *geom, ref.data_ptr(), k, rbar.data_ptr(), ws2.data_ptr(), nb2, st)
    assert rc == 0 and torch.equal(rbar, want)


@pytest.mark.parametrize("k,n", [(12289, 800001), (32768, 1 << 21), (57344, 3670017), (57345, 3670081),
                                 (262144, (1 << 24) + 5),                       # range kernel: 1, 1, 1, 2, 5 ranges
                                 (12289, 70000), (32768, 1 << 20), (262144, (1 << 21) + 5),   # few indices per unit
                                 (148 * 57344 + 1, 100000)])                    # more ranges than SMs
def test_histogram_range_partitioned_paths(k, n):
    """Private shared-memory counters per unit range (1, 2 and 5 ranges; many and few indices per unit), direct global
    atomics beyond 148 ranges; skewed hits, a base address that is only 8-byte aligned, out-of-range
    values next to range edges."""
    g = torch.Generator().manual_seed(k % 1000)
    idx = (torch.rand(n, generator=g).pow(3) * k).long().clamp_(0, k - 1)
    idx[::5] = k - 1                                     # heavy hitter in the LAST range
    idx[1::11] = 57343 % k                               # ... and one on a range edge
    want = torch.bincount(idx, minlength=k)
    dev = idx.to(DEV)
    assert torch.equal(ops.histogram(dev, k).cpu(), want)
    shifted = torch.empty(n + 1, dtype=torch.int64, device=DEV)
    shifted[1:] = dev
    assert torch.equal(ops.histogram(shifted[1:], k).cpu(), want)       # scalar (unaligned) loop
    bad = idx.clone()
    bad[:3] = torch.tensor([-1, k, -(1 << 40)])
    assert int(ops.histogram(bad.to(DEV), k).sum()) == n - 3


@pytest.mark.parametrize("shape,patch", [((5, 4, 32, 32), (2, 2)), ((3, 4, 32, 32), (4, 4)), ((3, 4, 32, 32), (8, 8)),
                                         ((7, 4, 32, 32), (32, 32)), ((2, 3, 16, 24), (4, 2)), ((2, 3, 16, 24), (2, 4)),
                                         ((2, 2, 12, 24), (3, 6)), ((3, 1, 8, 6), (2, 3)), ((2, 2, 8, 8), (1, 1)),
                                         ((4101, 4, 32, 32), (4, 4))])
def test_quantize_every_layout_matches_view_ops(shape, patch):
    """Gather + fused unpatchify (output-ordered kernel for W % 4 == 0 with pW % 4 == 0 or pW == 2, patch-ordered
    kernel otherwise) against table[idx] pushed through the reference's view-op unpatchify (models/layers.py:37-71)."""
    n, c, h, w = shape
    d = c * patch[0] * patch[1]
    seq = (h // patch[0]) * (w // patch[1])
    k = 37
    g = torch.Generator().manual_seed(d)
    table = torch.randn(k, d, generator=g)
    idx = torch.randint(0, k, (n * seq,), generator=g)
    want = somcb.unpatchify(table[idx].view(n, seq, d), image_dim=(h, w), patch_dim=patch)
    geom = ops.geometry(shape, patch)
    got = ops.quantize(idx.to(DEV), table.to(DEV), geom)
    assert torch.equal(got.cpu(), want)
    # guard band: an output view that is only 4-byte aligned takes the patch-ordered kernel and stays inside
    buf = torch.full((n * c * h * w + 2,), 7.0, device=DEV)
    ops.quantize(idx.to(DEV), table.to(DEV), geom, out=buf[1:-1].view(n, c, h, w))
    assert torch.equal(buf[1:-1].view(n, c, h, w).cpu(), want)
    assert float(buf[0]) == 7.0 and float(buf[-1]) == 7.0


@pytest.mark.parametrize("k,d", [(1000, 256), (1000, 18), (513, 3), (70000, 16)])
def test_gather_rows_matches_indexing(k, d):
    g = torch.Generator().manual_seed(d)
    w = torch.randn(k, d, generator=g)
    keep = torch.nonzero(torch.rand(k, generator=g) < 0.4).flatten()
    assert torch.equal(ops.gather_rows(w.to(DEV), keep.to(DEV)).cpu(), w[keep])


@pytest.mark.parametrize("k,d", [(1, 16), (1000, 3), (4099, 8), (4099, 16), (4099, 17), (16385, 64), (77, 4096),
                                 (300000, 16)])
def test_codebook_norms_every_row_width(k, d):
    """||c||^2 with 8-, 16- and 32-lane groups per unit against fp64."""
    g = torch.Generator().manual_seed(k)
    w = torch.randn(k, d, generator=g)
    got = ops.prepare_codebook(w.to(DEV)).cpu().double()
    want = w.double().pow(2).sum(1)
    assert float(((got - want).abs() / want).max()) <= 2e-6


def test_adam_kernel_vector_body_and_tail_agree():
    """The 16-byte body and the scalar tail apply the same rule: a 4-byte-shifted view (scalar loop only) and the
    aligned tensor (vector body + tail) give bit-identical results."""
    g = torch.Generator().manual_seed(3)
    n = 100003
    w0, m0, gr = (torch.randn(n, generator=g).to(DEV) for _ in range(3))
    v0 = torch.rand(n, generator=g).to(DEV)
    res = []
    for shift in (0, 1):
        bufs = [torch.zeros(n + 1, device=DEV) for _ in range(4)]
        views = [b[shift:shift + n] for b in bufs]
        for dst, src in zip(views, (w0, m0, v0, gr)):
            dst.copy_(src)
        ops.adam_step(views[0], views[1], views[2], views[3], 1e-3, 5)
        res.append([t.clone() for t in views[:3]])
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_filter_ws_entry_point_contract():
    """som_filter_ws_f32: workspace size query is 0 for shapes that take the FFMA kernel; a NULL workspace falls back
    to som_filter_f32 (bit-identical); a short workspace is SOM_E_WORKSPACE; in-place and NULL arguments are rejected."""
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(2)
    w = torch.randn(4096, 64, generator=g).to(DEV)
    need = lib.som_filter_workspace_bytes(4096, 64, 2048.0)
    assert need > 0
    assert lib.som_filter_workspace_bytes(1024, 64, 512.0) == 0          # 8 tiles: FFMA kernel
    assert lib.som_filter_workspace_bytes(4096, 16, 2048.0) == 0         # short rows: FFMA kernel
    assert lib.som_filter_workspace_bytes(0, 64, 2048.0) == 0
    a, b = torch.empty_like(w), torch.empty_like(w)
    rc = lib.som_filter_ws_f32(w.data_ptr(), a.data_ptr(), 4096, 64, 2048.0, 1.0, None, 0, st)
    assert rc == 0
    assert lib.som_filter_f32(w.data_ptr(), b.data_ptr(), 4096, 64, 2048.0, 1.0, st) == 0
    assert torch.equal(a, b)
    ws = torch.empty(need, dtype=torch.uint8, device=DEV)
    rc = lib.som_filter_ws_f32(w.data_ptr(), a.data_ptr(), 4096, 64, 2048.0, 1.0, ws.data_ptr(), need - 1, st)
    assert rc == -3 and b"workspace" in lib.som_last_error()
    assert lib.som_filter_ws_f32(w.data_ptr(), a.data_ptr(), 4096, 64, 2048.0, 1.0, ws.data_ptr(), need, st) == 0
    assert float((a - b).norm() / b.norm()) <= 2e-6                     # tensor-core vs FFMA result
    assert lib.som_filter_ws_f32(w.data_ptr(), w.data_ptr(), 4096, 64, 2048.0, 1.0, ws.data_ptr(), need, st) == -1
    assert lib.som_filter_ws_f32(None, a.data_ptr(), 4096, 64, 2048.0, 1.0, ws.data_ptr(), need, st) == -1
    assert lib.som_filter_ws_f32(w.data_ptr(), a.data_ptr(), 4096, 64, -1.0, 1.0, ws.data_ptr(), need, st) == -1


def test_empty_batch_is_a_no_op_with_the_reference_shapes():
    """A (0, C, H, W) batch: the reference returns an empty int64 index tensor ((0,) and (0, Seq) reshaped,
    models/Codebook.py:77-99) and an empty image batch from get_quantized_image (:138-154).  torch hands out
    null data pointers for zero-element tensors; the C-ABI accepts them when the count is zero."""
    w = trained_like_codebook(256, (4, 4), 3)
    cb = somcb.Codebook(patch_dim=(4, 4), image_dim=(32, 32), image_channel=4, num_embeddings=256,
                        init_neighbour_range=128)
    with torch.no_grad():
        cb.codebook.weight.copy_(w)
    cb = cb.to(DEV)
    x = torch.zeros(0, 4, 32, 32, device=DEV)
    for variant in (ops.SOM_BMU_AUTO, ops.SOM_BMU_FFMA, ops.SOM_BMU_TC3X):
        cb.bmu_variant = variant
        idx = cb.get_patches_bmu(x)
        assert idx.dtype == torch.int64 and tuple(idx.shape) == (0,)
        assert tuple(cb.get_patches_bmu(x, reshape=True).shape) == (0, 64)
    img = cb.get_quantized_image(torch.zeros(0, 64, dtype=torch.int64, device=DEV))
    assert tuple(img.shape) == (0, 4, 32, 32)
    counts = ops.histogram(torch.zeros(0, dtype=torch.int64, device=DEV), 256)
    assert int(counts.sum()) == 0
    # the raw entry points: null data pointers are only accepted together with a zero count
    lib = _lib.load()
    cn = ops.prepare_codebook(cb.codebook.weight.detach())
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.som_bmu_nchw_f32(None, 0, 4, 32, 32, 4, 4, cb.codebook.weight.data_ptr(), cn.data_ptr(), 256, 0,
                              None, None, None, 0, ops.SOM_BMU_AUTO, st)
    assert rc == 0
    rc = lib.som_bmu_nchw_f32(None, 1, 4, 32, 32, 4, 4, cb.codebook.weight.data_ptr(), cn.data_ptr(), 256, 0,
                              None, None, None, 0, ops.SOM_BMU_AUTO, st)
    assert rc == -1 and b"null pointer" in lib.som_last_error()


def test_multi_gpu_paths_under_torchrun():
    """With at least two GPUs visible: tests/multi_gpu_check.py under torch.distributed.run over NCCL -- data-parallel
    trainer (eager and CUDA-graph, ragged shares) == one-GPU trainer with bit-identical replicas, unit-sharded search ==
    unsharded search, sharded histogram.  Skips on a one-GPU box (the gloo tests cover the host logic on CPU)."""
    import os
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29731",
                        os.path.join(root, "tests", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
