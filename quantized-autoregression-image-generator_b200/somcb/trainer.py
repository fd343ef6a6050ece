"""Fused SOM training step, hit histogram and pruning -- the fast path next to the drop-in.

``SomTrainer.step`` performs exactly the arithmetic of one iteration of
/root/reference/train_codebook.py:225-249 + :300-304 (forward with use_gaussian=True, mse_loss,
backward, Adam(betas=(0.5, 0.999)).step, lr halving, neighbourhood decrease), but through the
factorised form of SURVEY.md A.3, never materialising the quantised batch or the N x K Gaussian:

    W~   = T @ W                                   som_filter_f32
    bmu  = argmin_j ||x - W_j||                    som_bmu_nchw_f32
    Rbar = segsum_bmu(W~[bmu] - x), SSE            som_accumulate_nchw_f32
           (data parallel: one all-reduce of Rbar and SSE here)
    G    = (2 / numel) * T @ Rbar                  som_filter_f32
    Adam(W, G)                                     som_adam_f32

The drop-in ``Codebook`` + torch autograd + torch.optim.Adam path produces the same update; the
tests check both against the reference's outputs.
"""
import torch

from . import ops as _default_ops


class SomTrainer:
    def __init__(self, codebook, lr, neighbourhood_step, lr_step=100000, global_steps=0,
                 betas=(0.5, 0.999), eps=1e-8, ops=None, reduce_fn=None, world_size=1,
                 use_cuda_graph=False):
        """``codebook``: a somcb.Codebook on a CUDA device.  ``reduce_fn(list_of_tensors)`` sums
        tensors in place across data-parallel ranks (None: single device).

        ``use_cuda_graph``: small batches (BASELINE config 1: 512 patches per step) are bound by the
        ~10 launches of a step, not by the kernels.  With this flag the step is captured once per
        (batch shape, neighbourhood range, lr) into a CUDA graph and replayed: the batch is copied
        into a static buffer, Adam's step count lives in device memory (``som_adam_devstep_f32``).
        ``use_cuda_graph="alias"`` captures on the caller's own input buffer instead of copying into a
        static one: for loops that refill ONE device staging buffer every step (a changed address forces a
        re-capture).  Single-device only; the first step always runs eagerly (it performs the one-time
        kernel attribute set-up that must not happen during capture)."""
        self.cb = codebook
        self.lr = float(lr)
        self.neighbourhood_step = int(neighbourhood_step)
        self.lr_step = int(lr_step)
        self.global_steps = int(global_steps)
        self.betas = betas
        self.eps = eps
        self.ops = ops or _default_ops
        self.reduce_fn = reduce_fn
        self.world_size = int(world_size)
        w = codebook.codebook.weight
        # the reference does not checkpoint Adam state: a resume restarts the moments at zero
        self.m = torch.zeros_like(w, requires_grad=False)
        self.v = torch.zeros_like(w, requires_grad=False)
        self.t = 0
        self.last_bmu = None
        self.use_cuda_graph = bool(use_cuda_graph) and reduce_fn is None
        self._graph_alias = use_cuda_graph == "alias"
        self._graph = None            # (key, CUDAGraph, x_static, loss_static, t_dev)

    @torch.no_grad()
    def step(self, feature_map, bmu=None):
        """One training step on a (local) batch; returns the loss as a 0-dim float64 device
        tensor (global mean squared error, as F.mse_loss over the global batch)."""
        if self.use_cuda_graph and bmu is None and self.t >= 1 and feature_map.is_cuda:
            return self._graph_step(feature_map)
        return self._eager_step(feature_map, bmu)

    def _graph_step(self, feature_map):
        cb = self.cb
        alias = self._graph_alias and feature_map.is_contiguous() and feature_map.dtype == torch.float32
        key = (tuple(feature_map.shape), float(cb.neighbourhood_range), float(self.lr),
               feature_map.data_ptr() if alias else 0)
        if self._graph is None or self._graph[0] != key:
            x_static = feature_map if alias else \
                torch.empty(feature_map.shape, dtype=torch.float32, device=feature_map.device)
            t_dev = torch.full((1,), self.t, dtype=torch.int64, device=feature_map.device)
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(feature_map.device)
            with torch.cuda.graph(graph):
                loss_static = self._eager_step(x_static, None, t_dev=t_dev, bookkeeping=False)
            self._graph = (key, graph, x_static, loss_static, t_dev)
        _, graph, x_static, loss_static, t_dev = self._graph
        if x_static is not feature_map:
            x_static.copy_(feature_map)
        graph.replay()
        self.t += 1
        cb._norm_cache = None
        self.last_bmu = None
        self._bookkeeping()
        return loss_static.clone()

    def _bookkeeping(self):
        # schedule bookkeeping, in the reference's order (train_codebook.py:247-249, 300-304)
        if self.global_steps % self.lr_step == 0 and self.global_steps > 0:
            self.lr = self.lr * 0.5
        self.global_steps += 1
        if self.global_steps % self.neighbourhood_step == 0:
            self.cb.decrease_neighbourhood(steps=1)

    def _eager_step(self, feature_map, bmu=None, t_dev=None, bookkeeping=True):
        ops = self.ops
        cb = self.cb
        w = cb.codebook.weight.data
        x, geom = cb._input(feature_map, require_cuda=getattr(ops, 'REQUIRES_CUDA', True))
        k = cb.num_embeddings
        rng = cb.neighbourhood_range

        wt = ops.neighbourhood_filter(w, rng)
        if bmu is None:
            bmu = ops.bmu(x, geom, w, ops.prepare_codebook(w), variant=cb.bmu_variant)
        numel = x.numel() * self.world_size            # every rank holds an equal share
        if self.reduce_fn is None:
            rbar, _, sse = ops.accumulate(x, geom, bmu, wt, k, want_sse=True)
        else:
            # ONE collective per step: Rbar and the squared error travel in one packed fp32
            # buffer; the fp64 SSE is carried as a (hi, lo) float pair.
            kd = k * cb.embedding_dim
            packed = torch.empty(kd + 2, dtype=torch.float32, device=x.device)
            rbar = packed[:kd].view(k, cb.embedding_dim)
            _, _, sse = ops.accumulate(x, geom, bmu, wt, k, want_sse=True, out=rbar)
            hi = sse.to(torch.float32)
            packed[kd:kd + 1] = hi
            packed[kd + 1:kd + 2] = (sse - hi.double()).to(torch.float32)
            self.reduce_fn(packed)
            sse = packed[kd:kd + 1].double() + packed[kd + 1:kd + 2].double()
        grad = ops.neighbourhood_filter(rbar, rng, scale=2.0 / numel)
        if t_dev is not None:                          # graph capture: step count in device memory
            ops.adam_step_dev(w, self.m, self.v, grad, self.lr, t_dev, self.betas, self.eps)
        else:
            self.t += 1
            ops.adam_step(w, self.m, self.v, grad, self.lr, self.t, self.betas, self.eps)
        cb._norm_cache = None                          # W changed under torch's feet
        self.last_bmu = bmu
        loss = (sse / numel).reshape(())
        if bookkeeping:
            self._bookkeeping()
        return loss

    def checkpoint_dict(self, image_channel):
        """The reference's checkpoint layout (train_codebook.py:271-278)."""
        cb = self.cb
        return {"patch_dim": cb.patch_dim, "image_dim": cb.image_dim, "image_C": image_channel,
                "num_embeddings": cb.num_embeddings, "neighbourhood_range": cb.neighbourhood_range,
                "global_steps": self.global_steps, "checkpoint": cb.state_dict()}


@torch.no_grad()
def bmu_histogram(codebook, batches, counts=None, ops=None):
    """Hit counts over an iterable of feature-map batches (prune_codebook.py:129-142), kept on
    the device: one BMU launch + one histogram launch per batch, no host round trip."""
    ops = ops or _default_ops
    for fmap in batches:
        idx = codebook.get_patches_bmu(fmap)
        counts = ops.histogram(idx, codebook.num_embeddings, counts)
    if counts is None:
        raise ValueError("bmu_histogram: no batches")
    return counts


@torch.no_grad()
def prune_codebook(codebook, counts, prune_threshold, image_channel=None, global_steps=0, ops=None):
    """Keep units with count >= threshold in ascending index order and build the smaller
    codebook + the reference's checkpoint dict (prune_codebook.py:144-178)."""
    from .codebook import Codebook
    ops = ops or _default_ops
    keep = torch.nonzero(counts >= prune_threshold).flatten().to(torch.int64)
    rows = ops.gather_rows(codebook.codebook.weight.data, keep)
    if image_channel is None:
        image_channel = codebook.embedding_dim // (codebook.patch_dim[0] * codebook.patch_dim[1])
    new_cb = Codebook(patch_dim=codebook.patch_dim, image_dim=codebook.image_dim,
                      image_channel=image_channel, num_embeddings=int(keep.numel()),
                      init_neighbour_range=codebook.neighbourhood_range).to(rows.device)
    new_cb.codebook.weight.data.copy_(rows)
    ck = {"patch_dim": codebook.patch_dim, "image_dim": codebook.image_dim, "image_C": image_channel,
          "num_embeddings": int(keep.numel()), "neighbourhood_range": codebook.neighbourhood_range,
          "global_steps": global_steps, "checkpoint": new_cb.state_dict()}
    return new_cb, keep, ck
