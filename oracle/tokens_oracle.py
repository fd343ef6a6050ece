"""CPU restatement of the dual-codebook tokenisation of the Transformer trainer.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/train_quantized_transformer.py:411-455 with the reference's own torch
calls (cat / repeat / add), on OracleCodebook modules.  Pinned: oracle/make_golden.py executes the
reference script's own lines 407-455 (read from the reference tree at generation time, never copied
here) on reference Codebook modules and commits tests/golden/tokens_case.pt;
tests/test_oracle_golden.py checks this restatement against it bit-exactly.
"""
import torch


@torch.no_grad()
def tokenize_pair_oracle(lr_codebook, hr_codebook, feature_map, train_base_model):
    """(hr_input, hr_target, lr_input) as train_quantized_transformer.py builds them."""
    n = feature_map.shape[0]
    lr_num_embeddings = lr_codebook.num_embeddings
    hr_num_embeddings = hr_codebook.num_embeddings
    lr_indices = lr_codebook.get_patches_bmu(feature_map, reshape=True)          # :413-415
    hr_indices = hr_codebook.get_patches_bmu(feature_map, reshape=True)          # :419-421
    if train_base_model:
        hr_indices_shifted = hr_indices + lr_num_embeddings                      # :425
        hr_input = torch.cat((lr_indices, hr_indices_shifted), dim=1)            # :428-430
        lr_input = None                                                          # :433
    else:
        start_tensor = torch.tensor([[hr_num_embeddings]]).repeat(n, 1)          # :436-438
        hr_input = torch.cat((start_tensor, hr_indices), dim=1)                  # :439-441
        lr_input = lr_indices                                                    # :444
    end_tensor = torch.tensor([[hr_num_embeddings]]).repeat(n, 1)                # :449-451
    hr_target = torch.cat((hr_indices, end_tensor), dim=1)                       # :452-454
    return hr_input, hr_target, lr_input
