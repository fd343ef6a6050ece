"""GPU parity of the BMU search (K1) through the C-ABI, against the golden reference outputs,
the live CPU oracle at moderate sizes, and size-independent properties at BASELINE sizes."""
import pytest
import torch

import somcb
from somcb import ops
from oracle import OracleCodebook
from oracle.step_oracle import make_oracle_codebook, synthetic_fmaps, trained_like_codebook
from _helpers import CASES, assert_bmu_parity, flat_patches, load_case, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _f16_ok(d, k):
    """Shapes an FP16-split kernel covers (include/somcb.h, SOM_BMU_TC_F16)."""
    return d <= 16 or (d <= 256 and (k + 255) // 256 >= 3)


def _variants(n_patches, d, k):
    """FFMA, the tensor-core variant with its static split rule, and both split arithmetics forced."""
    v = [ops.SOM_BMU_FFMA]
    lib = somcb._lib.load()
    if lib.som_bmu_workspace_bytes(n_patches, d, k, ops.SOM_BMU_TC3X) > 0:
        v += [ops.SOM_BMU_TC3X, ops.SOM_BMU_TC_TF32]
        if _f16_ok(d, k):
            v.append(ops.SOM_BMU_TC_F16)
    return v


def _gpu_cb(weight, patch_dim, image_dim, channels, rng, variant=ops.SOM_BMU_AUTO):
    cb = somcb.Codebook(patch_dim=patch_dim, image_dim=image_dim, image_channel=channels,
                        num_embeddings=weight.shape[0], init_neighbour_range=rng)
    with torch.no_grad():
        cb.codebook.weight.copy_(weight)
    cb = cb.to(DEV)
    cb.bmu_variant = variant
    return cb


@pytest.mark.parametrize("name", CASES)
def test_bmu_matches_reference_golden(name):
    rec = load_case(name)
    x = rec["x"].to(DEV)
    flat = flat_patches(rec["x"], rec["patch_dim"])
    for variant in _variants(flat.shape[0], flat.shape[1], rec["weight"].shape[0]) + [ops.SOM_BMU_AUTO]:
        cb = _gpu_cb(rec["weight"], rec["patch_dim"], rec["image_dim"], rec["channels"],
                     rec["neighbourhood_range"], variant)
        idx = cb.get_patches_bmu(x)
        assert idx.dtype == torch.int64 and idx.shape == rec["bmu"].shape and not idx.requires_grad
        n_bad = assert_bmu_parity(idx, rec["bmu"], flat, rec["weight"])
        if "trained" in name:
            assert n_bad == 0, f"{n_bad} index mismatches on a trained-like codebook (variant {variant})"
        assert cb.get_patches_bmu(x, reshape=True).shape == rec["bmu_reshaped"].shape


def test_bmu_ties_resolve_to_lowest_index():
    rec = load_golden("case_ties.pt")
    flat = flat_patches(rec["x"], rec["patch_dim"])
    for variant in _variants(flat.shape[0], flat.shape[1], rec["weight"].shape[0]):
        cb = _gpu_cb(rec["weight"], rec["patch_dim"], rec["image_dim"], rec["channels"], 80, variant)
        idx = cb.get_patches_bmu(rec["x"].to(DEV)).cpu()
        assert int(idx.max()) < 64, "a duplicated row beat its first copy"
        assert_bmu_parity(idx, rec["bmu"], flat, rec["weight"])


@pytest.mark.parametrize("shape", [
    # (fmaps, patch, K, init)                     exercises
    (64, (2, 2), 4096, "trained"),              # BASELINE C2 shape, 16 384 patches
    (64, (2, 2), 4096, "fresh"),
    (32, (4, 4), 16384, "trained"),             # BASELINE C4 shape, 2 048 patches
    (2, (4, 4), 16384, "fresh"),                # few patches, many units -> unit-split + merge
    (16, (4, 4), 1000, "trained"),              # K not a multiple of the unit tile
    (3, (8, 8), 777, "trained"),                # 48 patches (partial patch tile), D = 256
    (64, (32, 32), 512, "trained"),             # BASELINE C3 shape: D = 4096, Seq = 1
    (256, (2, 2), 4093, "trained"),             # tcgen05 config S with a ragged last unit chunk
    (40, (4, 2), 600, "trained"),               # D = 32: config M, pW = 2 loads
    (16, (8, 8), 2048, "trained"),              # BASELINE C5 patch size: D = 256, config L (streamed A')
])
def test_bmu_matches_live_oracle(shape):
    b, pd, k, init = shape
    x = synthetic_fmaps(b, 4242)
    d = 4 * pd[0] * pd[1]
    if init == "trained":
        w = trained_like_codebook(k, pd, 7)
    else:
        torch.manual_seed(0)
        w = torch.empty(k, d).uniform_(-1 / k, 1 / k)
    oc = make_oracle_codebook(w, pd, (32, 32), 4, k // 2)
    with torch.no_grad():
        ref = oc.get_patches_bmu(x)
    flat = flat_patches(x, pd)
    for variant in _variants(flat.shape[0], d, k):
        cb = _gpu_cb(w, pd, (32, 32), 4, k // 2, variant)
        idx = cb.get_patches_bmu(x.to(DEV))
        n_bad = assert_bmu_parity(idx, ref, flat, w)
        if init == "trained":
            assert n_bad <= max(1, flat.shape[0] // 20000), f"{n_bad} mismatches (variant {variant})"


def test_bmu_called_in_grad_mode_on_noncontiguous_input():
    rec = load_case("c1_trained")
    cb = _gpu_cb(rec["weight"], rec["patch_dim"], rec["image_dim"], rec["channels"], 512)
    x = rec["x"].to(DEV)
    xt = x.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2)      # same values, strided
    assert not xt.is_contiguous()
    with torch.enable_grad():
        idx = cb.get_patches_bmu(xt, reshape=True)
    assert idx.dtype == torch.int64 and not idx.requires_grad
    assert_bmu_parity(idx, rec["bmu"], flat_patches(rec["x"], rec["patch_dim"]), rec["weight"])


def test_unit_sharded_emulation_on_one_gpu():
    """Shards are searched one after another on one device (never as concurrent waiting kernels)
    and merged with som_merge_candidates; the result equals the unsharded search."""
    rec = load_golden("case_ties.pt")
    x = rec["x"].to(DEV)
    w = rec["weight"].to(DEV)
    geom = ops.geometry(x.shape, rec["patch_dim"])
    full = ops.bmu(x, geom, w)
    for world in (2, 5):
        rds, idxs = [], []
        for r in range(world):
            lo, hi = somcb.shard_bounds(w.shape[0], world, r)
            i, rd = ops.bmu(x, geom, w[lo:hi].contiguous(), unit_offset=lo, want_rd=True)
            assert int(i.min()) >= lo and int(i.max()) < hi
            rds.append(rd)
            idxs.append(i)
        merged, _ = ops.merge_candidates(torch.stack(rds), torch.stack(idxs))
        assert torch.equal(merged, full)
    assert_bmu_parity(full, rec["bmu"], flat_patches(rec["x"], rec["patch_dim"]), rec["weight"])


def test_bmu_full_size_c2_properties():
    """BASELINE config 2 at full size (39 063 fmaps -> 10 000 128 patches, D=16, K=4096):
    batch-split invariance, histogram sum, and a 65 536-patch oracle sub-sample."""
    pd, k = (2, 2), 4096
    n_f = 39063
    g = torch.Generator(device=DEV).manual_seed(123)
    x = torch.tanh(torch.randn(n_f, 4, 32, 32, generator=g, device=DEV))
    w = trained_like_codebook(k, pd, 7)
    cb = _gpu_cb(w, pd, (32, 32), 4, k // 2)
    idx = cb.get_patches_bmu(x, reshape=True)
    assert idx.shape == (n_f, 256)
    half = n_f // 2
    a = cb.get_patches_bmu(x[:half].contiguous(), reshape=True)
    b = cb.get_patches_bmu(x[half:].contiguous(), reshape=True)
    assert torch.equal(torch.cat([a, b]), idx)
    counts = ops.histogram(idx.reshape(-1), k)
    assert int(counts.sum()) == n_f * 256 and int(idx.min()) >= 0 and int(idx.max()) < k
    assert torch.equal(counts.cpu(), torch.bincount(idx.reshape(-1).cpu(), minlength=k))
    # permutation invariance: shuffling the fmaps permutes the indices the same way
    perm = torch.randperm(4096, device=DEV)
    assert torch.equal(cb.get_patches_bmu(x[:4096][perm].contiguous(), reshape=True), idx[:4096][perm])
    # four strided 65 536-patch windows (the last one holds the ragged tail) against the oracle, in the rule's FP16
    # split and, on the same windows, in 3xTF32
    oc = make_oracle_codebook(w, pd, (32, 32), 4, k // 2)
    cb32 = _gpu_cb(w, pd, (32, 32), 4, k // 2, ops.SOM_BMU_TC_TF32)
    for lo in (0, 13000, 26000, n_f - 256):
        sub = x[lo:lo + 256].cpu()
        with torch.no_grad():
            ref = oc.get_patches_bmu(sub)
        n_bad = assert_bmu_parity(idx[lo:lo + 256].reshape(-1), ref, flat_patches(sub, pd), w)
        assert n_bad <= 4
        n_bad = assert_bmu_parity(cb32.get_patches_bmu(x[lo:lo + 256].contiguous()), ref, flat_patches(sub, pd), w)
        assert n_bad <= 4


@pytest.mark.parametrize("geom", [
    # (N, C, H, W, pH, pW, K)
    (64, 3, 16, 16, 2, 2, 700),                 # D = 12: config S, scalar tail
    (48, 3, 12, 18, 2, 3, 333),                 # D = 18, pW = 3: scalar loads everywhere
    (96, 1, 16, 16, 2, 2, 512),                 # D = 4: one k-block
    (80, 5, 8, 8, 4, 4, 260),                   # D = 80 > 73: config L with a small D
])
def test_bmu_generic_geometry_both_variants(geom):
    n, c, h, w, ph, pw, k = geom
    g = torch.Generator().manual_seed(n * k)
    x = torch.tanh(torch.randn(n, c, h, w, generator=g))
    d = c * ph * pw
    flat = flat_patches(x, (ph, pw))
    wgt = flat[torch.randperm(flat.shape[0], generator=g)[:k]].clone()
    wgt += 0.05 * torch.randn(k, d, generator=g)
    oc = OracleCodebook(patch_dim=(ph, pw), image_dim=(h, w), image_channel=c, num_embeddings=k,
                        init_neighbour_range=k // 2)
    with torch.no_grad():
        oc.codebook.weight.copy_(wgt)
        ref = oc.get_patches_bmu(x)
    for variant in (ops.SOM_BMU_FFMA, ops.SOM_BMU_TC3X):
        cb = _gpu_cb(wgt, (ph, pw), (h, w), c, k // 2, variant)
        idx = cb.get_patches_bmu(x.to(DEV))
        n_bad = assert_bmu_parity(idx, ref, flat, wgt)
        assert n_bad <= 1, f"{n_bad} mismatches, variant {variant}"


def test_tensor_core_ties_inside_and_across_chunks():
    """Exact duplicates inside one 8-unit refine chunk, in another chunk, another 256-unit tile and
    in the padding tail: the tensor-core variant must still return the first copy."""
    pd = (2, 2)
    x = synthetic_fmaps(64, 99)
    base = trained_like_codebook(300, pd, 13)
    w = torch.cat([base, base[:5], base[100:108], base, base[:40]], dim=0).contiguous()     # K = 653
    oc = make_oracle_codebook(w, pd, (32, 32), 4, 300)
    with torch.no_grad():
        ref = oc.get_patches_bmu(x)
    assert int(ref.max()) < 300
    for variant in (ops.SOM_BMU_FFMA, ops.SOM_BMU_TC3X):
        cb = _gpu_cb(w, pd, (32, 32), 4, 300, variant)
        idx = cb.get_patches_bmu(x.to(DEV)).cpu()
        assert int(idx.max()) < 300, f"variant {variant}: a later duplicate won"
        assert_bmu_parity(idx, ref, flat_patches(x, pd), w)


@pytest.mark.parametrize("variant", [ops.SOM_BMU_TC_TF32, ops.SOM_BMU_TC_F16])
@pytest.mark.parametrize("fmaps,p,k", [(640, 4, 2048), (2400, 8, 1024), (1200, 8, 3000)])
def test_cta_pair_paths_of_both_arithmetics(variant, fmaps, p, k):
    """The static dispatch uses CTA pairs (cta_group::2) from two waves of 128-patch tiles on (296 tiles on a
    148-SM part): 40 960 patches of D = 64 and 38 400 / 19 200 of D = 256 (the last one: one wave, unpaired 3xTF32,
    FP16 not selected by the rule but forced) -- each arithmetic against the live oracle."""
    pd = (p, p)
    x = synthetic_fmaps(fmaps, 91)
    w = trained_like_codebook(k, pd, 13)
    oc = make_oracle_codebook(w, pd, (32, 32), 4, k // 2)
    with torch.no_grad():
        ref = oc.get_patches_bmu(x)
    cb = _gpu_cb(w, pd, (32, 32), 4, k // 2, variant)
    idx = cb.get_patches_bmu(x.to(DEV)).cpu()
    n_bad = assert_bmu_parity(idx, ref, flat_patches(x, pd), w)
    assert n_bad == 0, f"{n_bad} index mismatches on a trained-like codebook (variant {variant})"


@pytest.mark.parametrize("pd,fmaps", [((4, 4), 64), ((8, 8), 16)])
def test_unit_split_keeps_the_first_of_duplicates_across_splits(pd, fmaps):
    """Few patch tiles against >= 8 unit tiles: the tensor-core kernel spreads the unit tiles over CTAs and
    merges candidates (resident-A mode at D = 64, streamed mode at D = 256).  The codebook is three copies of
    the same 700 units, so every winner has exact duplicates in other splits: the lowest index must survive."""
    x = synthetic_fmaps(fmaps, 5)
    base = trained_like_codebook(700, pd, 17)
    w = torch.cat([base, base, base], dim=0).contiguous()          # K = 2100 -> 9 unit tiles
    oc = make_oracle_codebook(w, pd, (32, 32), 4, 700)
    with torch.no_grad():
        ref = oc.get_patches_bmu(x)
    assert int(ref.max()) < 700
    cb = _gpu_cb(w, pd, (32, 32), 4, 700, ops.SOM_BMU_TC3X)
    idx = cb.get_patches_bmu(x.to(DEV)).cpu()
    assert int(idx.max()) < 700, "a duplicate from a later unit split won"
    assert_bmu_parity(idx, ref, flat_patches(x, pd), w)


@pytest.mark.parametrize("fmaps,p,k,bound", [(64, 32, 512, 5e-6), (32, 8, 2048, 3e-6), (16, 4, 4096, 1e-6),
                                              (64, 2, 4096, 1e-6)])
def test_reduced_distance_accuracy_of_both_variants(fmaps, p, k, bound):
    """The returned reduced distance rd = ||c||^2 - 2 x.c against fp64, relative to ||c||^2 + 2 sum|x c|.  The
    tensor-core variant carries the truncation bias of the fp32 TMEM accumulator (measured -2e-6 at D = 4096,
    -9e-7 at D = 256, -1e-7 at D = 64: DESIGN 5); the bound guards the accumulation scheme against regressions."""
    pd = (p, p)
    d = 4 * p * p
    x = synthetic_fmaps(fmaps, 17)
    w = trained_like_codebook(k, pd, 3)
    flat = flat_patches(x, pd).double()
    geom = ops.geometry(x.shape, pd)
    xd, wd = x.to(DEV), w.to(DEV)
    cn = ops.prepare_codebook(wd)
    for variant in _variants(flat.shape[0], d, k):
        idx, rd = ops.bmu(xd, geom, wd, cn, want_rd=True, variant=variant)
        wi = w.double()[idx.cpu()]
        true = (wi * wi).sum(1) - 2 * (flat * wi).sum(1)
        scale = (wi * wi).sum(1) + 2 * (flat * wi).abs().sum(1)
        err = float(((rd.double().cpu() - true).abs() / scale).max())
        assert err <= bound, f"variant {variant}: rd error {err:.2e} > {bound}"


@pytest.mark.parametrize("fmaps,p,k", [(64, 2, 4096), (320, 2, 4096), (16, 4, 4096), (64, 4, 2100), (32, 8, 2048),
                                       (16, 8, 20000), (64, 32, 512), (3, 8, 777)])
def test_non_finite_patches_take_unit_zero_and_leave_their_neighbours_alone(fmaps, p, k):
    """Non-finite input (a diverged encoder).  The reference's distance row of a patch that holds a NaN is all
    NaN and torch.argmin returns its first position, unit 0 (models/Codebook.py:86-94; the training loop then
    stops on its NaN-loss guard, train_codebook.py:237-238).  Every kernel mode must return 0 for such a patch
    as well, an in-range index for a patch that holds +-inf (the reference's own pick there depends on which
    inf - inf of its sgemm turns into NaN first), and exactly the clean-run index for every other patch, also
    for the rows that share a 128-row MMA tile with a poisoned one.  Shapes: config S in its TF32 and (81 920 patches)
    FP16-split mode, resident-A, unit split, streamed, streamed with a large codebook, split-K, ragged."""
    pd = (p, p)
    d = 4 * p * p
    x = synthetic_fmaps(fmaps, 23)
    w = trained_like_codebook(k, pd, 5)
    seq = (32 // p) ** 2
    n = fmaps * seq
    g = torch.Generator().manual_seed(n + k)
    poisoned = torch.randperm(n, generator=g)[:max(3, min(48, n // 8))]
    kinds = [float("nan"), float("inf"), float("-inf")]
    xp = x.clone()
    for t, pi in enumerate(poisoned.tolist()):
        img, s = divmod(pi, seq)
        ph, pw = divmod(s, 32 // p)
        f = int(torch.randint(0, d, (1,), generator=g))
        c, r = divmod(f, p * p)
        i, j = divmod(r, p)
        xp[img, c, ph * p + i, pw * p + j] = kinds[t % 3]
    nan_rows = poisoned[0::3]
    clean = torch.ones(n, dtype=torch.bool)
    clean[poisoned] = False
    oc = make_oracle_codebook(w, pd, (32, 32), 4, k // 2)
    with torch.no_grad():
        ref = oc.get_patches_bmu(xp[:min(fmaps, max(8, 2048 // seq))])
    m = ref.numel()
    assert bool((ref[nan_rows[nan_rows < m]] == 0).all()), "the oracle itself does not send NaN patches to unit 0"
    for variant in _variants(n, d, k):
        cb = _gpu_cb(w, pd, (32, 32), 4, k // 2, variant)
        base = cb.get_patches_bmu(x.to(DEV)).cpu()
        idx = cb.get_patches_bmu(xp.to(DEV)).cpu()
        assert int(idx.min()) >= 0 and int(idx.max()) < k, f"variant {variant}: index outside [0, K)"
        assert bool((idx[nan_rows] == 0).all()), f"variant {variant}: NaN patch not at unit 0: {idx[nan_rows]}"
        assert torch.equal(idx[clean], base[clean]), f"variant {variant}: a poisoned row leaked into its neighbours"
        same = (idx[:m] == ref)[clean[:m]]
        assert float(same.float().mean()) > 0.999


@pytest.mark.parametrize("variant", [ops.SOM_BMU_TC_F16, ops.SOM_BMU_TC_TF32])
def test_split_arithmetics_over_magnitudes(variant):
    """Both operand splits of the tensor-core kernels -- FP16 (kind::f16, per-patch and per-codebook power-of-two
    scaling) and TF32 -- on config S (D <= 16), D = 64 and D = 256: trained-like and fresh codebooks, ragged D, and
    data / codebook magnitudes from 1e-6 to 1e4 (outside FP16's range without the scaling)."""
    runs = [((64, 2, 4096), 1.0, False), ((64, 2, 4096), 1.0, True), ((48, 2, 777), 1.0, False),
            ((64, 2, 4096), 1e-6, False), ((64, 2, 4096), 1e-3, False), ((64, 2, 4096), 1e4, False),
            ((64, 1, 300), 1.0, False), ((40, 2, 20000), 1.0, False),
            ((320, 4, 2048), 1.0, False), ((320, 4, 2048), 1.0, True), ((320, 4, 2048), 1e-6, False),
            ((320, 4, 2048), 1e-3, False), ((320, 4, 2048), 1e4, False), ((64, 4, 1000), 1.0, False),
            ((160, 8, 1024), 1.0, False), ((160, 8, 1024), 1e-6, False), ((160, 8, 1024), 1e4, False),
            ((160, 8, 1024), 1.0, True)]
    for (fmaps, p, k), scale, fresh in runs:
        pd = (p, p)
        d = 4 * p * p
        x = synthetic_fmaps(fmaps, 4242)
        if fresh:
            w = torch.empty(k, d).uniform_(-1 / k, 1 / k, generator=torch.Generator().manual_seed(0))
        else:
            w = trained_like_codebook(k, pd, 7)
        x, w = x * scale, w * scale
        oc = make_oracle_codebook(w, pd, (32, 32), 4, k // 2)
        with torch.no_grad():
            ref = oc.get_patches_bmu(x)
        cb = _gpu_cb(w, pd, (32, 32), 4, k // 2, variant)
        idx = cb.get_patches_bmu(x.to(DEV)).cpu()
        n_bad = assert_bmu_parity(idx, ref, flat_patches(x, pd), w)
        if scale == 1.0 and not fresh:
            assert n_bad == 0, f"variant {variant} {fmaps, p, k}: {n_bad} near-tie diffs on a trained-like codebook"


def test_mixed_row_magnitudes_fp16_split():
    """Patch rows of wildly different magnitude inside ONE 128-row tile (each row carries its own power-of-two
    scale) and an all-zero row, FP16 split at D = 64 and D = 16, against the fp64 argmin."""
    from oracle import bmu_fp64
    for p, k in ((4, 2048), (2, 4096)):
        pd = (p, p)
        x = synthetic_fmaps(64, 7)
        mags = 10.0 ** torch.randint(-6, 5, (64, 1, 32 // p, 1, 32 // p, 1),
                                     generator=torch.Generator().manual_seed(3)).float()
        x = (x.reshape(64, 4, 32 // p, p, 32 // p, p) * mags).reshape(64, 4, 32, 32).contiguous()
        x[5] = 0.0
        w = trained_like_codebook(k, pd, 7)
        flat = flat_patches(x, pd)
        ref = bmu_fp64(flat, w)
        cb = _gpu_cb(w, pd, (32, 32), 4, k // 2, ops.SOM_BMU_TC_F16)
        idx = cb.get_patches_bmu(x.to(DEV)).cpu()
        assert_bmu_parity(idx, ref, flat, w)


def test_c4_shape_parity_65536_patches_both_arithmetics():
    """BASELINE config 4's shape (D = 64, K = 16 384) on 65 536 patches -- 512 patch tiles, the CTA-pair path -- against
    the chunked live oracle, for the rule's choice (FP16 split) and for 3xTF32."""
    pd, k = (4, 4), 16384
    x = synthetic_fmaps(1024, 314)
    w = trained_like_codebook(k, pd, 7)
    oc = make_oracle_codebook(w, pd, (32, 32), 4, k // 2)
    with torch.no_grad():
        ref = torch.cat([oc.get_patches_bmu(x[i:i + 128]) for i in range(0, 1024, 128)])
    flat = flat_patches(x, pd)
    assert somcb._lib.load().som_bmu_split_mode(65536, 64, k) == 1
    for variant in (ops.SOM_BMU_AUTO, ops.SOM_BMU_TC_TF32):
        cb = _gpu_cb(w, pd, (32, 32), 4, k // 2, variant)
        idx = cb.get_patches_bmu(x.to(DEV)).cpu()
        n_bad = assert_bmu_parity(idx, ref, flat, w)
        assert n_bad <= 2, f"variant {variant}: {n_bad} near-tie differences"


def test_c5_full_codebook_262144_units():
    """BASELINE config 5 with ALL 262 144 units (D = 256, 268 MB of codebook) on one GPU: 1024 patches against the live
    oracle, 38 400 patches (FP16 split, CTA pairs) against the eight-shard search merged with som_merge_candidates."""
    pd, k = (8, 8), 262144
    gw = torch.Generator().manual_seed(77)
    w = torch.tanh(torch.randn(k, 256, generator=gw))
    x = synthetic_fmaps(2400, 2718)
    oc = make_oracle_codebook(w, pd, (32, 32), 4, k // 2)
    with torch.no_grad():
        ref = oc.get_patches_bmu(x[:64])
    wd, xd = w.to(DEV), x.to(DEV)
    geom_s = ops.geometry(x[:64].shape, pd)
    small = ops.bmu(xd[:64].contiguous(), geom_s, wd).cpu()
    assert_bmu_parity(small, ref, flat_patches(x[:64], pd), w)
    geom = ops.geometry(x.shape, pd)
    full = ops.bmu(xd, geom, wd)
    assert_bmu_parity(full[:1024].cpu(), ref, flat_patches(x[:64], pd), w)
    rds, idxs = [], []
    for r in range(8):
        lo, hi = somcb.shard_bounds(k, 8, r)
        i, rd = ops.bmu(xd, geom, wd[lo:hi].contiguous(), unit_offset=lo, want_rd=True)
        rds.append(rd)
        idxs.append(i)
    merged, _ = ops.merge_candidates(torch.stack(rds), torch.stack(idxs))
    diff = torch.nonzero(merged != full).flatten().cpu()
    if diff.numel():          # shards carry their own codebook scale: near-ties may resolve differently
        flat = flat_patches(x, pd)[diff]
        assert_bmu_parity(merged.cpu()[diff], full.cpu()[diff], flat, w)
    assert diff.numel() <= 4
