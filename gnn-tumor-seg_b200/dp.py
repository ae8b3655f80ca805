"""Whole-graph data parallelism over the GPUs of one box.

The reference is single-device (model/gnn_model.py:23); its batch is a
block-diagonal union of whole graphs (data_processing/data_loader.py:168), so
the natural shard is the graph: rank r takes graphs r, r+R, ... of the global
batch.  One collective per step: a sum all-reduce (NCCL over NVLink/NVSwitch) of
a flat fp32 arena holding every parameter gradient plus two trailing scalars,
[sum_i w[y_i]*nll_i, sum_i w[y_i]].  Each rank back-propagates the
UN-normalised weighted loss sum; after the all-reduce the gradients are divided
by the global weight sum, which makes R ranks x B graphs numerically the same
step as one device on the union batch of R*B graphs (the reference's
CrossEntropyLoss(weight) is a weighted MEAN over the whole batch).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world_size: int):
    """Round-robin shard: rank r owns items r, r+R, r+2R, ..."""
    return list(range(rank, n_items, world_size))


class GradArena:
    """One contiguous fp32 buffer; every parameter's .grad is a view into it.
    Layout: [grads of p0 | grads of p1 | ... | loss_sum | weight_sum]."""

    def __init__(self, params, n_extra: int = 2):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradArena needs at least one trainable parameter")
        dev = self.params[0].device
        self.n_grad = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.n_grad + n_extra, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.extra = self.flat[self.n_grad:]

    @property
    def grads(self):
        return self.flat[:self.n_grad]

    def zero_(self):
        self.flat.zero_()


class _CeSums(torch.autograd.Function):
    """[sum w*nll, sum w] with the un-normalised gradient (fused K8 kernel)."""

    @staticmethod
    def forward(ctx, logits, labels, class_w):
        from . import ops
        sums, dl = ops.ce_weighted(logits, labels, class_w, want_grad=True)
        ctx.save_for_backward(dl)
        return sums

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g[0], None, None


def ce_sums_device(logits, labels, class_w):
    return _CeSums.apply(logits, labels, class_w)


class DataParallelTrainer:
    """Data-parallel fwd + loss + bwd + gradient all-reduce for any module with
    the reference call convention ``net(graph, feats)``.

    ``loss_sums_fn(logits, labels, w) -> tensor[2] = [sum w*nll, sum w]``
    (differentiable in its first entry); default: the fused device kernel.
    """

    def __init__(self, net, class_weights, process_group=None, loss_sums_fn=None):
        self.net = net
        self.class_weights = class_weights
        self.pg = process_group
        self.loss_sums_fn = loss_sums_fn or ce_sums_device
        self.arena = GradArena(net.parameters())
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1

    def forward_backward(self, graph, feats, labels):
        """Returns the GLOBAL weighted-mean loss (0-d tensor, no host sync);
        parameter .grad hold the global-batch gradients afterwards."""
        self.arena.zero_()
        logits = self.net(graph, feats)
        sums = self.loss_sums_fn(logits, labels, self.class_weights)
        sums[0].backward()                      # un-normalised: d(sum w*nll)/dtheta accumulates into the arena
        with torch.no_grad():
            self.arena.extra.copy_(sums.detach())
            if self.world_size > 1:
                dist.all_reduce(self.arena.flat, op=dist.ReduceOp.SUM, group=self.pg)
            denom = self.arena.extra[1]
            self.arena.grads.div_(denom)
            return self.arena.extra[0] / denom
