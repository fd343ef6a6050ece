"""CPU test double for ``somcb.ops`` (TEST INFRASTRUCTURE, lives in tests/ only).

Implements the ops interface with plain torch on CPU in the factorised form of SURVEY.md A.3, so
the host-side logic (SomTrainer orchestration, data-parallel packing + all-reduce, unit-sharded
merge) can be exercised under gloo without a GPU.  The product never imports this.
"""
import math

import torch

from oracle import neighbourhood_two_var, patchify, unpatchify
from somcb.ops import geometry, flat_geometry, n_patches_of, dim_of  # noqa: F401  (pure python)

REQUIRES_CUDA = False
SOM_BMU_AUTO = 0


def _flat(x, geom):
    n, c, h, w, p_h, p_w = geom
    p = patchify(x.reshape(n, c, h, w), (p_h, p_w))
    return p.reshape(-1, p.shape[-1])


def _toeplitz(k, rng):
    ids = torch.arange(k)
    return torch.exp(-(((ids.unsqueeze(0) - ids.unsqueeze(1)) ** 2) / neighbourhood_two_var(rng)))


def prepare_codebook(weight, out=None):
    return (weight * weight).sum(dim=1)


def neighbourhood_filter(inp, neighbourhood_range, scale=1.0, out=None):
    res = (_toeplitz(inp.shape[0], neighbourhood_range).double() @ inp.double() * scale).float()
    if out is not None:
        out.copy_(res)
        return out
    return res


def bmu(x, geom, weight, c_norm2=None, unit_offset=0, want_rd=False, variant=0, out=None):
    flat = _flat(x, geom).double()
    w = weight.double()
    rd = (w * w).sum(dim=1).unsqueeze(0) - 2.0 * flat @ w.t()
    val, idx = rd.min(dim=1)
    idx = idx + unit_offset
    if out is not None:
        out.copy_(idx)
        idx = out
    return (idx, val.float()) if want_rd else idx


def merge_candidates(rd, idx):
    r, n = rd.shape
    best_rd, best_idx = rd[0].clone(), idx[0].clone()
    for i in range(1, r):
        take = (rd[i] < best_rd) | ((rd[i] == best_rd) & (idx[i] < best_idx))
        best_rd = torch.where(take, rd[i], best_rd)
        best_idx = torch.where(take, idx[i], best_idx)
    return best_idx, best_rd


def histogram(idx, num_units, counts=None):
    if counts is None:
        counts = torch.zeros(num_units, dtype=torch.int64)
    ok = (idx >= 0) & (idx < num_units)
    counts += torch.bincount(idx[ok], minlength=num_units)
    return counts


def accumulate(x, geom, bmu_idx, table, num_units, want_counts=False, want_sse=False, out=None):
    flat = _flat(x, geom).double()
    d = flat.shape[1]
    if table is not None:
        r = table.double()[bmu_idx] - flat
    else:
        r = flat
    rbar = torch.zeros(num_units, d, dtype=torch.float64).index_add_(0, bmu_idx, r).float()
    if out is not None:
        out.copy_(rbar)
        rbar = out
    counts = torch.bincount(bmu_idx, minlength=num_units) if want_counts else None
    sse = (r ** 2).sum().reshape(1) if (want_sse and table is not None) else (
        torch.zeros(1, dtype=torch.float64) if want_sse else None)
    return rbar, counts, sse


def accumulate_packed(x, geom, bmu_idx, table, num_units, packed=None, ws=None):
    d = dim_of(geom)
    if packed is None:
        packed = torch.empty(num_units * d + 4, dtype=torch.float32)
    rbar, _, sse = accumulate(x, geom, bmu_idx, table, num_units, want_sse=True)
    packed[:num_units * d] = rbar.reshape(-1)
    hi = sse.to(torch.float32)
    n = n_patches_of(geom)
    packed[num_units * d:] = torch.tensor([float(hi), float((sse - hi.double()).to(torch.float32)),
                                           float(n >> 12), float(n & 4095)])
    return packed


def adam_step_dp(weight, m, v, grad, dim, lr, steps_done, tail, loss_out=None, betas=(0.5, 0.999), eps=1e-8):
    numel = (float(tail[2]) * 4096.0 + float(tail[3])) * dim
    step = int(steps_done[0]) + 1
    adam_step(weight, m, v, grad * torch.tensor(2.0 / numel, dtype=torch.float32), lr, step, betas, eps)
    steps_done[0] += 1
    loss = ((tail[0].double() + tail[1].double()) / numel).reshape(1)
    if loss_out is not None:
        loss_out.copy_(loss)
        return loss_out
    return loss


def quantize(idx, table, geom, out=None):
    n, c, h, w, p_h, p_w = geom
    q = table[idx].reshape(n, -1, table.shape[1])
    return unpatchify(q, (h, w), (p_h, p_w)).contiguous()


def adam_step(weight, m, v, grad, lr, step, betas=(0.5, 0.999), eps=1e-8):
    b1, b2 = betas
    m.lerp_(grad, 1 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    weight.addcdiv_(m, denom, value=-(lr / bc1))
    return weight


def gather_rows(weight, keep):
    return weight[keep].clone()
