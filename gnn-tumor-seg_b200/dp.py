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

    Gradient buffer: when the whole backward ran through gts_sage_backward, every
    parameter gradient already is a view of ONE flat device buffer
    (ops.last_flat_grads) with two spare floats at its end; the loss sums are
    written there and the buffer is all-reduced in place — no per-parameter copy
    or add kernels.  Any other network (GAT, the CPU oracle in the gloo tests)
    goes through a packed copy of the gradients.
    """

    def __init__(self, net, class_weights, process_group=None, loss_sums_fn=None):
        self.net = net
        self.class_weights = class_weights
        self.pg = process_group
        self.loss_sums_fn = loss_sums_fn or ce_sums_device
        self.params = [p for p in net.parameters() if p.requires_grad]
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.flat = None          # the buffer that was all-reduced in the last step
        self.n_grad = 0

    @property
    def grads(self):
        return self.flat[:self.n_grad]

    @property
    def extra(self):
        return self.flat[self.n_grad:self.n_grad + 2]

    def _flat_from_stack(self):
        """The flat buffer of the last stack backward, if it backs every parameter gradient."""
        if not self.class_weights.is_cuda:
            return None
        from . import ops
        flat = ops.last_flat_grads["flat"]
        if flat is None:
            return None
        base = flat.untyped_storage().data_ptr()
        for p in self.params:
            if p.grad is None or p.grad.untyped_storage().data_ptr() != base:
                return None
        return flat, ops.last_flat_grads["n_grad"]

    def forward_backward(self, graph, feats, labels):
        """Returns the GLOBAL weighted-mean loss (0-d tensor, no host sync);
        parameter .grad hold the global-batch gradients afterwards."""
        for p in self.params:
            p.grad = None                     # autograd then adopts the produced tensors: no accumulate kernels
        if self.class_weights.is_cuda:
            from . import ops
            ops.last_flat_grads["flat"] = None
        logits = self.net(graph, feats)
        sums = self.loss_sums_fn(logits, labels, self.class_weights)
        sums[0].backward()                      # un-normalised: d(sum w*nll)/dtheta
        with torch.no_grad():
            got = self._flat_from_stack()
            if got is not None:
                flat, n_grad = got
            else:                               # generic path: pack, reduce, hand views back
                n_grad = sum(p.numel() for p in self.params)
                flat = torch.empty(n_grad + 2, dtype=torch.float32, device=self.params[0].device)
                off = 0
                for p in self.params:
                    n = p.numel()
                    flat[off:off + n].copy_(p.grad.reshape(-1))
                    p.grad = flat[off:off + n].view_as(p)
                    off += n
            flat[n_grad:n_grad + 2].copy_(sums.detach())
            if self.world_size > 1:
                dist.all_reduce(flat[:n_grad + 2], op=dist.ReduceOp.SUM, group=self.pg)
            denom = flat[n_grad + 1]
            flat[:n_grad].div_(denom)
            self.flat, self.n_grad = flat, n_grad
            return flat[n_grad] / denom
