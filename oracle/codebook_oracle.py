"""Restatement of the reference's SOM codebook (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows, op for op and with the same torch calls (so it is bit-identical to the
reference on the same torch build):

* ``patchify`` / ``unpatchify``      -> /root/reference/models/layers.py:8-34, 37-71
* ``OracleCodebook.__init__``        -> /root/reference/models/Codebook.py:18-46
* ``custom_load_state_dict``         -> models/Codebook.py:48-66
* ``decrease_neighbourhood``         -> models/Codebook.py:68-74
* ``get_patches_bmu``                -> models/Codebook.py:77-99
* ``get_quantized_patches``          -> models/Codebook.py:102-135
* ``get_quantized_image``            -> models/Codebook.py:138-154
* ``forward``                        -> models/Codebook.py:156-164

plus fp64 "ground truth" helpers used by the near-tie parity rule.
"""
import math

import torch
import torch.nn as nn


def patchify(image, patch_dim=(4, 4)):
    """(N,C,H,W) -> (N, Seq, D); D index order (c, i, j).  models/layers.py:8-34."""
    p_h, p_w = patch_dim
    n, c, h, w = image.shape
    g_h, g_w = h // p_h, w // p_w
    t = image.reshape(n, c, g_h, p_h, g_w, p_w)
    t = t.permute(0, 2, 4, 1, 3, 5)
    return t.reshape(n, g_h * g_w, c * p_h * p_w)


def unpatchify(patches, image_dim=(32, 32), patch_dim=(4, 4)):
    """(N, Seq, D) -> (N,C,H,W).  models/layers.py:37-71."""
    i_h, i_w = image_dim
    p_h, p_w = patch_dim
    n, _, d = patches.shape
    g_h, g_w = i_h // p_h, i_w // p_w
    c = d // (p_h * p_w)
    t = patches.reshape(n, g_h, g_w, c, p_h, p_w)
    t = t.permute(0, 3, 1, 4, 2, 5)
    return t.reshape(n, c, p_h * g_h, p_w * g_w)


def neighbourhood_two_var(neighbourhood_range):
    """``2 * variance`` exactly as models/Codebook.py:118,123 computes it (Python double)."""
    variance = -(neighbourhood_range / (2 * math.log(0.1)))
    return 2 * variance


def band_half_width(neighbourhood_range):
    """Largest |j - bmu| whose fp32 Gaussian weight is non-zero."""
    two_var = neighbourhood_two_var(neighbourhood_range)
    t = torch.arange(0, int(math.sqrt(110.0 * two_var)) + 8)
    w = torch.exp(-((t ** 2) / two_var))
    nz = torch.nonzero(w > 0).flatten()
    return int(nz[-1].item())


class OracleCodebook(nn.Module):
    """Same constructor, attributes, methods and state_dict as the reference class."""

    def __init__(self, patch_dim=(2, 2), image_dim=(32, 32), image_channel=4,
                 num_embeddings=512, init_neighbour_range=256):
        super().__init__()
        # models/Codebook.py:27-28 (the `and` means this never fires; kept as is).
        if init_neighbour_range > num_embeddings and init_neighbour_range < 1:
            raise Exception("Invalid value for init_neighbour_range.")
        self.neighbourhood_range = init_neighbour_range
        self.patch_dim = patch_dim
        self.image_dim = image_dim
        p_h, p_w = self.patch_dim
        self.embedding_dim = image_channel * p_h * p_w
        self.num_embeddings = num_embeddings
        self.codebook = nn.Embedding(self.num_embeddings, self.embedding_dim)
        self.codebook.weight.data.uniform_(-1 / self.num_embeddings, 1 / self.num_embeddings)

    def custom_load_state_dict(self, state_dict, ignore_msgs=False):
        own_state = self.state_dict()
        for name, param in state_dict.items():
            if name not in own_state:
                if not ignore_msgs:
                    print(f"No Layer found: {name}, skipping")
                continue
            if own_state[name].shape != param.data.shape:
                if not ignore_msgs:
                    print(f"Skipped: {name}")
                continue
            if isinstance(param, torch.nn.parameter.Parameter):
                param = param.data
            own_state[name].copy_(param)

    def decrease_neighbourhood(self, steps=1):
        if steps < 1:
            raise Exception("Invalid value for steps, should be > 1.")
        self.neighbourhood_range = 1.0 if self.neighbourhood_range <= 1 \
            else self.neighbourhood_range - 1

    def get_patches_bmu(self, x, reshape=False):
        x_patches = patchify(x, self.patch_dim)
        n, seq, d = x_patches.shape
        flat = x_patches.reshape(n * seq, d)
        distances = torch.cdist(flat, self.codebook.weight)
        bmu = torch.argmin(distances, dim=-1, keepdim=False)
        if reshape:
            bmu = bmu.reshape(n, seq)
        return bmu

    def get_quantized_patches(self, x, use_gaussian=True):
        bmu = self.get_patches_bmu(x)
        n = x.shape[0]
        if use_gaussian:
            bmu = bmu.unsqueeze(dim=-1)
            unit_ids = torch.arange(start=0, end=self.codebook.num_embeddings
                                    ).unsqueeze(dim=0).to(x.device)
            variance = -(self.neighbourhood_range / (2 * math.log(0.1)))
            scale = torch.exp(-((unit_ids - bmu) ** 2 / (2 * variance)))
            q = torch.matmul(scale, self.codebook.weight)
        else:
            q = self.codebook(bmu)
        return q.view(n, -1, self.embedding_dim)

    def get_quantized_image(self, indices, unpatchify_input=True):
        n, seq = indices.shape
        q = self.codebook(indices.flatten()).view(n, seq, self.embedding_dim)
        if unpatchify_input:
            return unpatchify(q, self.image_dim, self.patch_dim)
        return q

    def forward(self, x, use_gaussian=True):
        q = self.get_quantized_patches(x, use_gaussian=use_gaussian)
        return unpatchify(q, self.image_dim, self.patch_dim)


# ----------------------------------------------------------------------------------------
# fp64 ground truth for the near-tie rule
# ----------------------------------------------------------------------------------------
def distance_fp64(flat_patches, weight, idx):
    """Direct Euclidean distance in fp64 between row i and unit idx[i]."""
    p = flat_patches.double()
    c = weight.double()[idx]
    return (p - c).pow(2).sum(dim=-1).sqrt()


def bmu_fp64(flat_patches, weight, chunk=4096):
    """argmin of the direct (non-expanded) fp64 distance; first index on ties."""
    out = torch.empty(flat_patches.shape[0], dtype=torch.int64)
    w = weight.double()
    for s in range(0, flat_patches.shape[0], chunk):
        p = flat_patches[s:s + chunk].double()
        d = torch.cdist(p, w, compute_mode="donot_use_mm_for_euclid_dist")
        out[s:s + chunk] = torch.argmin(d, dim=-1)
    return out
