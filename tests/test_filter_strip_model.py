"""CPU model of the index arithmetic of the tensor-core neighbourhood filter (csrc/som_filter_tc.cu): the banded product
as per-tile GEMMs over 32-row k-blocks, the A operand of 5 consecutive k-blocks as windows of ONE Toeplitz strip (unit
rows of the tile reversed), accumulation chains as contiguous quarters of the k-blocks, and the two-MMA form on the
concatenated [B_hi ; B_lo] tile.  Pure numpy: it pins the formulas the kernel's comments state against the dense
T @ W of models/Codebook.py:112-130, so an edit of the kernel's indexing has a reference to be checked against."""
import numpy as np
import pytest

TMU, KBLK, CK, NACC = 128, 32, 5, 4


def _weights(h, two_var):
    t = np.arange(-h - 200, h + 201)
    w = np.exp(-((t.astype(np.float64) ** 2) / two_var))
    w[np.abs(t) > h] = 0.0
    return {int(k): float(v) for k, v in zip(t, w)}


def _filter_by_strips(inp, h, two_var):
    k_units, d = inp.shape
    w = _weights(h, two_var)
    length = (TMU + 2 * h + KBLK - 1) // KBLK * KBLK           # padded inner extent of a tile
    nkb = length // KBLK
    q = (nkb + NACC - 1) // NACC                                # k-blocks per accumulation chain
    na = (nkb + q - 1) // q
    n_ut = (k_units + TMU - 1) // TMU
    kp = (n_ut - 1) * TMU + length
    bt = np.zeros((d, kp))                                      # split_in_t_kernel: column h + j holds in[j][:]
    bt[:, h:h + k_units] = inp.T
    out = np.zeros((k_units, d))
    for ut in range(n_ut):
        chains = np.zeros((na, TMU, d))
        for c0 in range(0, nkb, CK):                            # one strip per CK k-blocks
            nk = min(CK, nkb - c0)
            rows = TMU + KBLK * (nk - 1)
            strip = np.array([[w[KBLK * c0 + r + c - h - (TMU - 1)] for c in range(KBLK)] for r in range(rows)])
            for j in range(nk):
                kb = c0 + j
                a_blk = strip[KBLK * j:KBLK * j + TMU]          # rows [32 j, 32 j + 128): MMA row m = unit 127 - m
                b_blk = bt[:, ut * TMU + kb * KBLK:ut * TMU + (kb + 1) * KBLK]      # (d, 32)
                chains[kb // q] += a_blk @ b_blk.T
        acc = chains[0]
        for a in range(1, na):
            acc = acc + chains[a]
        for m in range(TMU):
            u = ut * TMU + (TMU - 1) - m
            if u < k_units:
                out[u] = acc[m]
    return out, (nkb, q, na)


@pytest.mark.parametrize("k_units,d,h,two_var", [(300, 6, 40, 300.0), (128, 3, 5, 4.0), (513, 4, 150, 5000.0), (260, 2, 0, 1.0)])
def test_strip_formulation_equals_dense_toeplitz(k_units, d, h, two_var):
    rng = np.random.default_rng(k_units + h)
    inp = rng.standard_normal((k_units, d))
    ids = np.arange(k_units)
    dist = ids[None, :] - ids[:, None]
    dense = np.where(np.abs(dist) <= h, np.exp(-((dist.astype(np.float64) ** 2) / two_var)), 0.0)
    want = dense @ inp
    got, (nkb, q, na) = _filter_by_strips(inp, h, two_var)
    assert nkb >= 4 and 1 <= na <= NACC and (na - 1) * q < nkb <= na * q       # every chain non-empty
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)


def test_two_products_on_the_concatenated_tile_equal_the_split_product():
    """(A_hi + A_lo) @ [B_hi ; B_lo]^T, column halves added = the full product of the split operands (the kernel's
    three-product form drops the lo.lo term, the two-MMA form keeps it)."""
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal((128, 32)), rng.standard_normal((64, 32))
    a_hi, b_hi = a.astype(np.float16).astype(np.float64), b.astype(np.float16).astype(np.float64)
    a_lo, b_lo = a - a_hi, b - b_hi
    cat = np.concatenate([b_hi, b_lo], axis=0)                  # 128 rows: one N = 128 operand
    acc = a_hi @ cat.T + a_lo @ cat.T
    full = acc[:, :64] + acc[:, 64:]
    np.testing.assert_allclose(full, a @ b.T, rtol=1e-12, atol=1e-12)
    three = a_hi @ b_hi.T + a_lo @ b_hi.T + a_hi @ b_lo.T
    assert np.abs(full - three).max() == pytest.approx(np.abs(a_lo @ b_lo.T).max(), rel=1e-6)
