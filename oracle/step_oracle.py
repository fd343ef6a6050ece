"""Training-step, histogram and synthetic-data oracles (TEST INFRASTRUCTURE).

* ``reference_step``   : the literal step body of /root/reference/train_codebook.py:225-249
                         (zero_grad -> forward(use_gaussian=True) -> mse_loss -> backward ->
                         Adam(betas=(0.5,0.999)).step), driven with an in-memory batch.
* ``closed_form_step`` : SURVEY.md §A.2 (dense S, closed-form gradient and Adam), any dtype.
* ``factorised_grad``  : SURVEY.md §A.3 (W~ = T W, per-unit residual sums, G = 2/numel T Rbar)
                         in fp64 -- the form the CUDA kernels implement.
* ``histogram_prune``  : /root/reference/prune_codebook.py:129-162.
* ``synthetic_fmaps`` / ``trained_like_codebook`` : the seeded inputs of SURVEY.md §8d.
"""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from .codebook_oracle import OracleCodebook, patchify, neighbourhood_two_var


# ----------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------
def synthetic_fmaps(batch, seed, channels=4, height=32, width=32):
    """tanh(randn) latents: the encoder's last activation is tanh (README.md:92-93)."""
    g = torch.Generator().manual_seed(int(seed))
    return torch.tanh(torch.randn(batch, channels, height, width, generator=g))


def trained_like_codebook(num_embeddings, patch_dim, seed=7, channels=4, height=32, width=32):
    """K distinct data patches drawn from an independent pool of synthetic fmaps."""
    p_h, p_w = patch_dim
    seq = (height // p_h) * (width // p_w)
    n_fmaps = max(2, (2 * num_embeddings + seq - 1) // seq)
    pool = patchify(synthetic_fmaps(n_fmaps, 1000 + seed, channels, height, width), patch_dim)
    pool = pool.reshape(-1, pool.shape[-1])
    g = torch.Generator().manual_seed(int(seed))
    pick = torch.randperm(pool.shape[0], generator=g)[:num_embeddings]
    return pool[pick].clone().contiguous()


# ----------------------------------------------------------------------------------------
# the reference step (train_codebook.py:183-186, 225-249)
# ----------------------------------------------------------------------------------------
def make_reference_optimizer(codebook, lr):
    return torch.optim.Adam(codebook.parameters(), lr=lr, betas=(0.5, 0.999))


def reference_step(codebook, optim, feature_map):
    """One literal reference step; returns the loss tensor (train_codebook.py:227-242)."""
    codebook.train()
    optim.zero_grad()
    quant = codebook(feature_map, use_gaussian=True)
    loss = F.mse_loss(quant, feature_map)
    if torch.isnan(loss):
        raise Exception("NaN encountered during training")
    loss.backward()
    optim.step()
    return loss.detach()


# ----------------------------------------------------------------------------------------
# closed forms
# ----------------------------------------------------------------------------------------
@dataclass
class AdamState:
    m: torch.Tensor
    v: torch.Tensor
    t: int = 0

    @staticmethod
    def zeros_like(w):
        return AdamState(torch.zeros_like(w), torch.zeros_like(w), 0)


def adam_update(w, g, st, lr, b1=0.5, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (no weight decay / amsgrad); SURVEY §A.2."""
    st.t += 1
    st.m.mul_(b1).add_(g, alpha=1 - b1)
    st.v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** st.t
    bc2 = 1 - b2 ** st.t
    step_size = lr / bc1
    denom = (st.v.sqrt() / math.sqrt(bc2)).add_(eps)
    w.addcdiv_(st.m, denom, value=-step_size)
    return w


def closed_form_step(w, st, feature_map, patch_dim, neighbourhood_range, lr, bmu=None):
    """Dense closed form of one step in w's dtype.  Returns (loss, bmu, grad)."""
    p = patchify(feature_map, patch_dim)
    p = p.reshape(-1, p.shape[-1]).to(w.dtype)
    if bmu is None:
        bmu = torch.argmin(torch.cdist(p, w), dim=-1)
    k = w.shape[0]
    two_var = neighbourhood_two_var(neighbourhood_range)
    ids = torch.arange(k).unsqueeze(0)
    s = torch.exp(-(((ids - bmu.unsqueeze(-1)) ** 2).to(w.dtype) / two_var))
    q = s @ w
    numel = feature_map.numel()
    loss = ((q - p) ** 2).sum() / numel
    grad = s.t() @ ((2.0 / numel) * (q - p))
    adam_update(w, grad, st, lr)
    return loss, bmu, grad


def factorised_grad(w, flat_patches, bmu, neighbourhood_range, numel, dtype=torch.float64):
    """returns (W~, Rbar, counts, sse, G) computed in ``dtype``."""
    k, d = w.shape
    two_var = neighbourhood_two_var(neighbourhood_range)
    ids = torch.arange(k)
    diff2 = (ids.unsqueeze(0) - ids.unsqueeze(1)) ** 2
    # fp32 weights exactly as the reference generates them, then widened
    t32 = torch.exp(-(diff2 / two_var))
    t = t32.to(dtype)
    wd = w.to(dtype)
    wt = t @ wd
    r = wt[bmu] - flat_patches.to(dtype)
    rbar = torch.zeros(k, d, dtype=dtype).index_add_(0, bmu, r)
    counts = torch.bincount(bmu, minlength=k)
    sse = (r ** 2).sum()
    g = (2.0 / numel) * (t @ rbar)
    return wt, rbar, counts, sse, g


# ----------------------------------------------------------------------------------------
# histogram + prune (prune_codebook.py:129-162)
# ----------------------------------------------------------------------------------------
def histogram_prune(codebook, batches, prune_threshold, literal_loop=True):
    """Returns (counts list, kept unit ids, new weight rows)."""
    k = codebook.num_embeddings
    total = {i: 0 for i in range(k)}
    for fmap in batches:
        codebook.eval()
        idx = codebook.get_patches_bmu(fmap)
        if literal_loop:
            for j in idx.tolist():
                total[j] += 1
        else:
            c = torch.bincount(idx, minlength=k).tolist()
            for j in range(k):
                total[j] += c[j]
    good = [i for i, c in total.items() if c >= prune_threshold]
    with torch.no_grad():
        new_rows = codebook.codebook.weight[good].clone()
    return [total[i] for i in range(k)], good, new_rows


def make_oracle_codebook(weight, patch_dim, image_dim, channels, neighbourhood_range):
    cb = OracleCodebook(patch_dim=patch_dim, image_dim=image_dim, image_channel=channels,
                        num_embeddings=weight.shape[0],
                        init_neighbour_range=neighbourhood_range)
    with torch.no_grad():
        cb.codebook.weight.copy_(weight)
    return cb
