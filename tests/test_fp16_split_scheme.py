"""CPU emulation of the config-S FP16-split arithmetic (csrc/som_bmu_tc_s.cu: cb_scale_kernel, split_w_s16_kernel and the
builders' per-patch scale), product by product, against the fp64 argmin: pins the scale rules and the claim that the
split is as index-stable as the TF32 one for any magnitude of data and codebook.  Products of two FP16 values are exact
in fp32; sums are taken in fp64 here, so the test isolates the operand split from the accumulator."""
import math

import pytest
import torch

from oracle.step_oracle import synthetic_fmaps, trained_like_codebook
from _helpers import flat_patches


def _exp2_floor(v):
    """floor(log2(v)) of a positive float, as the kernels read it from the exponent bits"""
    return math.frexp(float(v))[1] - 1


def _h(v):
    return v.half().float()


def fp16_split_rd(x, w):
    """s_p s_c rd per (patch, unit) exactly as the kernel's four MMAs form it (x, w fp32; D <= 16)"""
    cn = (w * w).sum(1)
    mc, mn = float(w.abs().max()), float(cn.abs().max())
    e = g = 0
    if mc > 0 and math.isfinite(mc) and math.isfinite(mn):
        e = max(-60, min(60, 6 - _exp2_floor(mc)))                  # max |s_c c| in [64, 128)
        msn = mn * 2.0 ** e
        if msn > 0:
            g = max(-60, min(60, 14 - _exp2_floor(msn)))            # max t_c s_c ||c||^2 in [2^14, 2^15)
    sc, tcs = 2.0 ** e, 2.0 ** g
    b = -2.0 * sc * w
    bh = _h(b)
    bl = _h(b - bh)
    n = cn * sc * tcs
    n1 = _h(n)
    n2 = _h(n - n1)
    n3 = _h(n - n1 - n2)
    m = x.abs().max(1).values
    ep = torch.tensor([6 - _exp2_floor(v) if v > 0 and math.isfinite(v) else 0 for v in m.tolist()])
    ka = ep - g
    far = ka < -24                                 # |x| beyond ~2^24 |c|: the norm column is dropped, the row stays in range
    ka = ka.clamp(-24, 15)                         # exact FP16 powers of two (subnormal below 2^-14)
    es = torch.where(far, ep, ka + g).clamp(-120, 120)
    sp = torch.pow(2.0, es.double()).float()
    ap = torch.where(far, torch.zeros(()), torch.pow(2.0, ka.double()).float())
    assert bool((_h(ap) == ap).all())
    a = x * sp[:, None]
    ah = _h(a)
    al = _h(a - ah)
    assert bool(torch.isfinite(ah).all()) and bool(torch.isfinite(bh).all()) and bool(torch.isfinite(n1).all())
    acc = (ap.double()[:, None] * (n1.double() + n2.double() + n3.double())[None, :]
           + ah.double() @ bh.double().T + al.double() @ bh.double().T + ah.double() @ bl.double().T)
    return acc, sp.double() * sc


@pytest.mark.parametrize("scale", [1.0, 1e-6, 1e-3, 1e4])
@pytest.mark.parametrize("fresh", [False, True])
def test_fp16_split_argmin_matches_fp64(scale, fresh):
    pd, k = (2, 2), 2048
    x = flat_patches(synthetic_fmaps(16, 5), pd) * scale
    if fresh:
        torch.manual_seed(0)
        w = torch.empty(k, 16).uniform_(-1 / k, 1 / k) * scale
    else:
        w = trained_like_codebook(k, pd, 7) * scale
    acc, factor = fp16_split_rd(x, w)
    x64, w64 = x.double(), w.double()
    true = (w64 ** 2).sum(1)[None, :] - 2 * x64 @ w64.T
    # the accumulator is a positive per-row multiple of rd
    rd = acc / factor[:, None]
    err = float(((rd - true).abs().max(1).values / true.abs().max(1).values).max())
    assert err <= 2e-6, f"reduced distance off by {err:.2e} relative"
    idx, ti = acc.argmin(1), true.argmin(1)
    bad = idx != ti
    if bad.any():
        xx = (x64 ** 2).sum(1)
        d_t = (true.gather(1, ti[:, None]).squeeze(1) + xx).clamp_min(0).sqrt()
        d_o = (true.gather(1, idx[:, None]).squeeze(1) + xx).clamp_min(0).sqrt()
        assert bool((((d_o - d_t) <= 1e-6 * d_t) | ~bad).all()), "a pick is worse than the fp64 optimum by more than 1e-6"
    if not fresh:
        assert int(bad.sum()) == 0, f"{int(bad.sum())} index mismatches on a trained-like codebook"


def test_fp16_split_degenerate_rows_and_codebooks():
    pd, k = (2, 2), 256
    x = flat_patches(synthetic_fmaps(2, 9), pd)
    x[0] = 0.0                                     # all-zero patch: scale exponent 0, rd = ||c||^2
    x[1] = 3.0e38                                  # near FLT_MAX: scaled DOWN into FP16's range
    x[2] = 1.0e-40                                 # fp32 subnormal
    w = trained_like_codebook(k, pd, 3)
    acc, factor = fp16_split_rd(x, w)
    assert bool(torch.isfinite(acc).all())
    true = (w.double() ** 2).sum(1)[None, :] - 2 * x.double() @ w.double().T
    assert int(acc[0].argmin()) == int(true[0].argmin())
    assert int(acc[2].argmin()) == int(true[2].argmin())
    # an all-zero codebook keeps both codebook scales at 1 and every distance at 0: unit 0 wins
    acc0, _ = fp16_split_rd(x[3:40], torch.zeros(k, 16))
    assert bool((acc0 == 0).all())
