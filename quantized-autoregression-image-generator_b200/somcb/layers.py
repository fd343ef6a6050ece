"""patchify / unpatchify with the reference's signatures (models/layers.py:8-71).

Inside the kernels the permutation is address arithmetic (csrc/som_common.cuh: patch_base +
feat_off); these torch view-op versions exist for API parity (reference scripts import them
next to Codebook) and for tests.  They are pure data movement, device-agnostic.
"""


def patchify(image, patch_dim=(4, 4)):
    """(N, C, H, W) -> (N, Seq, D) with D ordered (c, i, j)."""
    p_h, p_w = patch_dim
    n, c, h, w = image.shape
    g_h, g_w = h // p_h, w // p_w
    return (image.reshape(n, c, g_h, p_h, g_w, p_w)
            .permute(0, 2, 4, 1, 3, 5)
            .reshape(n, g_h * g_w, c * p_h * p_w))


def unpatchify(patches, image_dim=(32, 32), patch_dim=(4, 4)):
    """(N, Seq, D) -> (N, C, H, W)."""
    i_h, i_w = image_dim
    p_h, p_w = patch_dim
    n, _, d = patches.shape
    g_h, g_w = i_h // p_h, i_w // p_w
    c = d // (p_h * p_w)
    return (patches.reshape(n, g_h, g_w, c, p_h, p_w)
            .permute(0, 3, 1, 4, 2, 5)
            .reshape(n, c, p_h * g_h, p_w * g_w))
