"""Dual-codebook tokenisation for the Transformer trainer.

``tokenize_pair`` reproduces /root/reference/train_quantized_transformer.py:411-455: the same
feature maps are BMU-searched against the low- and the high-resolution codebook, then the token
tensors are assembled (HR indices shifted by ``lr_num_embeddings`` for the base model,
``hr_num_embeddings`` as <start>/<end> token).  Here: two BMU launches on the resident batch and
ONE assembly launch (``som_assemble_tokens_i64``) instead of add / cat / repeat / cat.
"""
import torch

from . import ops as _default_ops


@torch.no_grad()
def tokenize_pair(lr_codebook, hr_codebook, feature_map, train_base_model, ops=None):
    """Returns ``(hr_input, hr_target, lr_input)`` exactly as the reference builds them:

    * base model:  hr_input = cat(lr_indices, hr_indices + lr_K), lr_input = None
    * otherwise :  hr_input = cat(<start>, hr_indices),           lr_input = lr_indices
    * always    :  hr_target = cat(hr_indices, <end>),  <start> = <end> = hr_num_embeddings
    """
    ops = ops or _default_ops
    lr_idx = lr_codebook.get_patches_bmu(feature_map, reshape=True)      # (N, lr_Seq)
    hr_idx = hr_codebook.get_patches_bmu(feature_map, reshape=True)      # (N, hr_Seq)
    hr_input, hr_target = ops.assemble_tokens(lr_idx if train_base_model else None, hr_idx,
                                              lr_codebook.num_embeddings, hr_codebook.num_embeddings,
                                              train_base_model)
    return hr_input, hr_target, (None if train_base_model else lr_idx)
