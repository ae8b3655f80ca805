"""Oracle: DGL ``SAGEConv(in, out, 'pool')`` and the reference GraphSage stack.

TEST INFRASTRUCTURE (see oracle/__init__.py).  **Parity unpinned**: DGL is not
installable here; this follows SURVEY.md Appendix A.1 (DGL 0.6-2.x public
semantics) and the reference's stacking code model/networks.py:20-36.

Pure PyTorch on the CPU, works in fp32 and fp64.  The neighbour max uses an
explicit first-maximum-wins arg-max over the canonical CSR (row = in-edges of
v ordered by edge id) with a hand-written backward (scatter through the saved
arg-max), so tie-breaking is defined and checkable bit-exactly.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def segment_max_first_ref(P, indptr, indices, chunk=8192):
    """neigh[v,k] = max over in-edges (u->v) of P[u,k]; arg[v,k] = that u.
    Scan in CSR order, replace only on strictly greater -> FIRST maximum wins.
    Zero in-degree: neigh = 0, arg = -1 (DGL fills 0; its arg=0 gradient leak
    is deliberately not replicated)."""
    N, D = P.shape
    indptr_t = torch.as_tensor(np.asarray(indptr), dtype=torch.int64)
    indices_t = torch.as_tensor(np.asarray(indices), dtype=torch.int64)
    deg = indptr_t[1:] - indptr_t[:-1]
    neigh = torch.zeros_like(P)
    arg = torch.full((N, D), -1, dtype=torch.int64)
    neg_inf = torch.tensor(float("-inf"), dtype=P.dtype)
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        d = deg[lo:hi]
        md = int(d.max()) if hi > lo else 0
        if md == 0:
            continue
        ar = torch.arange(md).view(1, md)
        valid = ar < d.view(-1, 1)                                  # [n, md]
        pos = (indptr_t[lo:hi].view(-1, 1) + ar).clamp_(max=max(indices_t.numel() - 1, 0))
        nb = torch.where(valid, indices_t[pos], torch.zeros_like(pos))   # [n, md]
        vals = P[nb]                                                # [n, md, D]
        vals = torch.where(valid.unsqueeze(-1), vals, neg_inf)
        mx, am = vals.max(dim=1)        # torch returns the first maximal index
        has = d > 0
        a = torch.gather(nb, 1, am)                                 # [n, D]
        neigh[lo:hi] = torch.where(has.view(-1, 1), mx, torch.zeros_like(mx))
        arg[lo:hi] = torch.where(has.view(-1, 1), a, torch.full_like(a, -1))
    return neigh, arg


class _SegMaxFirst(torch.autograd.Function):
    @staticmethod
    def forward(ctx, P, indptr, indices):
        neigh, arg = segment_max_first_ref(P.detach(), indptr, indices)
        ctx.save_for_backward(arg)
        ctx.n = P.shape[0]
        ctx.mark_non_differentiable(arg)
        return neigh, arg

    @staticmethod
    def backward(ctx, d_neigh, _):
        (arg,) = ctx.saved_tensors
        dP = torch.zeros((ctx.n, d_neigh.shape[1]), dtype=d_neigh.dtype)
        m = arg >= 0
        cols = torch.arange(d_neigh.shape[1]).view(1, -1).expand_as(arg)
        dP.index_put_((arg[m], cols[m]), d_neigh[m], accumulate=True)
        return dP, None, None


def segment_max_first(P, indptr, indices):
    return _SegMaxFirst.apply(P, indptr, indices)


class SAGEConvPoolRef(nn.Module):
    """SAGEConv(in, out, 'pool', feat_drop=0, bias=True, activation) with the
    DGL<=0.7 parameter layout (fc_self.bias and fc_neigh.bias both present;
    effective bias is their sum) — the layout the 1 263 276-parameter count of
    the shipped 7x256 model implies."""

    def __init__(self, in_feats, out_feats, activation=None):
        super().__init__()
        self.fc_pool = nn.Linear(in_feats, in_feats)
        self.fc_self = nn.Linear(in_feats, out_feats)
        self.fc_neigh = nn.Linear(in_feats, out_feats)
        self.activation = activation
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_pool.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)
        self.last_argmax = None

    def forward(self, csr, h):
        indptr, indices = csr
        P = F.relu(self.fc_pool(h))
        neigh, arg = segment_max_first(P, indptr, indices)
        self.last_argmax = arg
        rst = self.fc_self(h) + self.fc_neigh(neigh)
        if self.activation is not None:
            rst = self.activation(rst)
        return rst


class GraphSageRef(nn.Module):
    """model/networks.py:20-36 with aggregator_type='pool', dropout=0."""

    def __init__(self, in_feats, layer_sizes, n_classes):
        super().__init__()
        self.layers = nn.ModuleList()
        self.layers.append(SAGEConvPoolRef(in_feats, layer_sizes[0], F.relu))
        for i in range(1, len(layer_sizes)):
            self.layers.append(SAGEConvPoolRef(layer_sizes[i - 1], layer_sizes[i], F.relu))
        self.layers.append(SAGEConvPoolRef(layer_sizes[-1], n_classes, None))

    def forward(self, csr, features):
        h = features
        for layer in self.layers:
            h = layer(csr, h)
        return h


def sage_pool_dense_ref(h, adj, Wp, bp, Ws, Wn, b, relu_out):
    """Independent formulation for tiny graphs: dense adjacency mask
    ``adj[v,u] = True`` iff edge u->v.  No arg-max, plain torch autograd."""
    P = torch.relu(h @ Wp.T + bp)
    masked = torch.where(adj.unsqueeze(-1), P.unsqueeze(0), torch.tensor(float("-inf"), dtype=h.dtype))
    neigh = masked.max(dim=1).values
    neigh = torch.where(adj.any(dim=1, keepdim=True), neigh, torch.zeros_like(neigh))
    out = h @ Ws.T + neigh @ Wn.T + b
    return torch.relu(out) if relu_out else out
