"""CPU feasibility study (no GPU): BMU through a two-way FP16 split (x = h + l, h = fp16(x), l = fp16(x - h);
x.c ~ h.h + h.l + l.h, fp32-class accumulation) against the 3xTF32 split the tcgen05 kernels use today and the
fp64 truth.  kind::f16 MMAs run at twice the kind::tf32 rate, so three FP16 MMAs cost 1.5 TF32 MMAs instead of 3.
Counts, per case: index mismatches vs fp64 argmin, and how many of those exceed the parity rule (relative fp64
distance gap > 1e-6).  Products of two fp16 / tf32 values are exact in fp32/fp64; sums are taken in fp64 here, so
the study isolates the SPLIT error from the accumulator's.
usage: python tools/fp16_split_study.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
from oracle.step_oracle import synthetic_fmaps, trained_like_codebook  # noqa: E402
from _helpers import flat_patches  # noqa: E402


def tf32(v):
    """round-to-nearest-away on the 13 dropped mantissa bits (cvt.rna.tf32.f32)"""
    i = v.contiguous().view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF)
    return r.view(torch.float32)


def split_tf32(v):
    h = tf32(v)
    return h, tf32(v - h)


def split_f16(v, scale=1.0):
    s = v * scale
    h = s.half().float()
    l = (s - h).half().float()
    return h / scale, l / scale


def rd_split(x, w, split):
    xh, xl = split(x)
    wh, wl = split(w)
    xh, xl, wh, wl = xh.double(), xl.double(), wh.double(), wl.double()
    dot = xh @ wh.T + xh @ wl.T + xl @ wh.T
    cn = (w.double() ** 2).sum(1)
    return cn[None, :] - 2 * dot


def study(name, x, w):
    x64, w64 = x.double(), w.double()
    true = (w64 ** 2).sum(1)[None, :] - 2 * x64 @ w64.T
    ti = true.argmin(1)
    xx = (x64 ** 2).sum(1)
    out = [name, f"n={x.shape[0]} D={x.shape[1]} K={w.shape[0]}"]
    for label, split in (("3xTF32", split_tf32), ("3xFP16", split_f16),
                         ("3xFP16 x2^8", lambda v: split_f16(v, 256.0))):
        rd = rd_split(x, w, split)
        idx = rd.argmin(1)
        bad = idx != ti
        d_t = (true.gather(1, ti[:, None]).squeeze(1) + xx).clamp_min(0).sqrt()
        d_o = (true.gather(1, idx[:, None]).squeeze(1) + xx).clamp_min(0).sqrt()
        worse = ((d_o - d_t) > 1e-6 * d_t) & bad
        err = ((rd - true).abs().max() / true.abs().max()).item()
        out.append(f"{label}: {int(bad.sum())} mismatches, {int(worse.sum())} beyond 1e-6, max rd err {err:.1e}")
    print(" | ".join(out), flush=True)


def main():
    torch.manual_seed(0)
    for fm, p, k in ((256, 2, 4096), (64, 4, 16384), (32, 8, 8192), (512, 32, 512)):
        pd = (p, p)
        x = flat_patches(synthetic_fmaps(fm, 11), pd)[:16384]
        study(f"trained-like P={p}", x, trained_like_codebook(k, pd, 7))
        d = 4 * p * p
        study(f"fresh init   P={p}", x, torch.empty(k, d).uniform_(-1 / k, 1 / k))
        study(f"small data   P={p}", x * 1e-3, trained_like_codebook(k, pd, 7) * 1e-3)


if __name__ == "__main__":
    main()
