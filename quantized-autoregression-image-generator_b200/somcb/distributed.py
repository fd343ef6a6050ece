"""Multi-GPU SOM codebook: one process per GPU, torch.distributed (NCCL over NVLink) as plumbing.

Two shardings:

* patch-sharded data parallel (``DataParallelSom``): every rank holds the full codebook and Adam
  state, takes about 1/R of the step's feature maps (shares may be ragged), runs BMU + per-unit
  accumulation locally, then ONE all-reduce(sum) of the packed [Rbar (K*D fp32) | SSE (2 fp32) |
  patch count (2 fp32)] buffer; filter + Adam are replicated, so replicas stay bit-identical (the
  all-reduce returns identical bits everywhere).  The whole step, collective included, is captured
  in a CUDA graph (``use_cuda_graph``).
  BMU-only / histogram workloads need no exchange beyond a final all-reduce of K int64 counts.
* unit-sharded search (``sharded_bmu``): rank r owns units [lo_r, hi_r); patches are replicated;
  each rank returns (reduced distance, global index) candidates from the SAME kernel arithmetic as
  the unsharded path, one all-gather of 12 B/patch/rank, then som_merge_candidates (smaller
  distance, tie -> smaller global index == the single-device first-minimum rule).

``ops`` is injectable so the host logic (sharding arithmetic, collective wiring) is testable
under gloo on CPU with a test double; the default is the CUDA library, which raises on CPU.
"""
import torch
import torch.distributed as dist

from . import ops as _default_ops
from .trainer import SomTrainer


def shard_bounds(total, world_size, rank):
    """Contiguous near-equal shard [lo, hi) of ``total`` items for ``rank``."""
    base, rem = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def split_batch(feature_map, world_size, rank):
    """Contiguous near-equal slice of the batch dimension for ``rank``.  Shares may be ragged (and empty when there
    are fewer feature maps than ranks): the step all-reduces the patch count next to the accumulators."""
    lo, hi = shard_bounds(feature_map.shape[0], world_size, rank)
    return feature_map[lo:hi]


class DataParallelSom(SomTrainer):
    """SomTrainer whose accumulators are summed across the process group once per step."""

    def __init__(self, codebook, lr, neighbourhood_step, group=None, **kw):
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        super().__init__(codebook, lr, neighbourhood_step, world_size=world,
                         reduce_fn=self._allreduce if world > 1 else None, **kw)

    def _allreduce(self, packed):
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group)

    def broadcast_weights(self, src=0):
        """Make every replica start from rank ``src``'s codebook."""
        dist.broadcast(self.cb.codebook.weight.data, src=src, group=self.group)
        self.cb._norm_cache = None


@torch.no_grad()
def sharded_bmu(x, geom, weight_shard, unit_offset, group=None, ops=None, c_norm2=None):
    """Unit-sharded BMU.  ``weight_shard``: this rank's rows [unit_offset, unit_offset + K_r).
    ``x`` is the same on every rank.  Returns the global indices (n_patches,) on every rank."""
    ops = ops or _default_ops
    idx, rd = ops.bmu(x, geom, weight_shard, c_norm2, unit_offset=unit_offset, want_rd=True)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return idx
    n = idx.numel()
    all_rd = torch.empty(world * n, dtype=rd.dtype, device=rd.device)
    all_idx = torch.empty(world * n, dtype=idx.dtype, device=idx.device)
    dist.all_gather_into_tensor(all_rd, rd, group=group)      # rank-major: row r = rank r
    dist.all_gather_into_tensor(all_idx, idx, group=group)
    merged, _ = ops.merge_candidates(all_rd.view(world, n), all_idx.view(world, n))
    return merged


@torch.no_grad()
def sharded_histogram(idx, lo, hi, counts=None, ops=None):
    """Counts of this rank's own unit range [lo, hi) from replicated merged indices: no reduction
    is needed afterwards -- the global histogram is the concatenation over ranks."""
    ops = ops or _default_ops
    return ops.histogram(idx - lo, hi - lo, counts)


@torch.no_grad()
def allreduce_counts(counts, group=None):
    """Patch-sharded histogram: sum the K int64 counts across ranks."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts
