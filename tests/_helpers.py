"""Shared test helpers: golden loading and the parity rules of SURVEY.md 8c."""
import os

import torch

from oracle import distance_fp64, patchify

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["c1_fresh", "c1_trained", "c2_trained", "c2_fresh", "c3_shape_small", "c5_shape_small",
         "odd_geom", "range_floor"]
BMU_EPS = 1e-6      # north-star: relative fp64 distance gap allowed between our pick and the oracle's
W_TOL = 1e-5        # north-star: norm-relative weight tolerance


def load_case(name):
    return torch.load(os.path.join(GOLDEN, f"case_{name}.pt"), weights_only=True)


def load_golden(fname):
    return torch.load(os.path.join(GOLDEN, fname), weights_only=True)


def flat_patches(x, patch_dim):
    p = patchify(x, patch_dim)
    return p.reshape(-1, p.shape[-1])


def assert_bmu_parity(idx_new, idx_ref, flat, weight, eps=BMU_EPS):
    """Indices equal, or our pick within eps (relative, fp64 direct distance) of the oracle's."""
    idx_new = idx_new.cpu().reshape(-1)
    idx_ref = idx_ref.cpu().reshape(-1)
    assert idx_new.shape == idx_ref.shape
    bad = torch.nonzero(idx_new != idx_ref).flatten()
    if bad.numel() == 0:
        return 0
    d_new = distance_fp64(flat[bad], weight, idx_new[bad])
    d_ref = distance_fp64(flat[bad], weight, idx_ref[bad])
    gap = (d_new - d_ref).abs()
    ok = gap <= eps * d_ref
    assert bool(ok.all()), (f"{int((~ok).sum())} of {idx_new.numel()} BMU picks are worse than the "
                            f"oracle's by more than {eps} relative (max {float((gap / d_ref).max()):.3e})")
    return int(bad.numel())


def rel_fro(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def rel_max(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def assert_close_norm(a, b, tol=W_TOL, what="tensor"):
    f, m = rel_fro(a, b), rel_max(a, b)
    assert f <= tol and m <= tol, f"{what}: rel_fro={f:.3e} rel_max={m:.3e} > {tol}"
