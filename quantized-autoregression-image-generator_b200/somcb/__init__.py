"""somcb -- Self-Organizing-Map codebook on B200 (sm_100a).

Drop-in for /root/reference/models/Codebook.py (class ``Codebook``: same constructor, attributes,
methods, state_dict and checkpoint layout) whose BMU search, neighbourhood-weighted update and
hit histogram run in hand-written CUDA behind the C-ABI of include/somcb.h.  No CPU fallback.
"""
SOM_ABI_VERSION = 2

from . import _lib, ops  # noqa: E402,F401
from .codebook import Codebook  # noqa: E402,F401
from .layers import patchify, unpatchify  # noqa: E402,F401
from .trainer import SomTrainer, prune_codebook, bmu_histogram  # noqa: E402,F401
from .distributed import (  # noqa: E402,F401
    DataParallelSom, sharded_bmu, shard_bounds, split_batch)
from .host_pipeline import HostTokenizer, HostTrainer, bind_host_to_gpu_node  # noqa: E402,F401
from .tokenizer import tokenize_pair  # noqa: E402,F401
from . import fmap_shards  # noqa: E402,F401
from .fmap_shards import ShardReader, convert_reference_dataset  # noqa: E402,F401

__all__ = ["Codebook", "patchify", "unpatchify", "SomTrainer", "prune_codebook", "bmu_histogram",
           "DataParallelSom", "sharded_bmu", "shard_bounds", "split_batch", "HostTokenizer", "HostTrainer", "bind_host_to_gpu_node", "tokenize_pair", "ShardReader", "convert_reference_dataset", "ops"]
