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


# ---------------------------------------------------------------------------
# Tie-aware comparison (tests/test_gpu_full_config.py).  The stack is piecewise linear: every arg-max of the
# neighbour pooling and every ReLU is a DECISION; two fp32 evaluations with different summation orders agree on all
# but the handful of decisions whose inputs sit within rounding of a tie, and each flipped decision re-routes a whole
# gradient contribution.  So the oracle can (a) report its own decisions and (b) evaluate forward + backward with the
# decisions of another run imposed: same routing, its own arithmetic — then gradients are comparable element-wise.
# ---------------------------------------------------------------------------
def graphsage_decisions(ref: "GraphSageRef", csr, features):
    """Free forward.  Returns (logits, decisions); decisions[l] = {'arg': int64 [N,Din] first-maximum arg-max,
    'neigh_pos': bool [N,Din] (pooled feature at the arg-max > 0, i.e. fc_pool's ReLU at the selected entry),
    'out_pos': bool [N,Dout] (output ReLU, None for the last layer)}."""
    indptr, indices = csr
    h = features
    decisions = []
    with torch.no_grad():
        for layer in ref.layers:
            P = F.relu(layer.fc_pool(h))
            neigh, arg = segment_max_first_ref(P, indptr, indices)
            rst = layer.fc_self(h) + layer.fc_neigh(neigh)
            d = {"arg": arg, "neigh_pos": neigh > 0, "out_pos": None}
            if layer.activation is not None:
                d["out_pos"] = rst > 0
                rst = layer.activation(rst)
            decisions.append(d)
            h = rst
    return h, decisions


def graphsage_forward_forced(ref: "GraphSageRef", features, decisions):
    """Differentiable forward with the arg-max routing and the ReLU masks IMPOSED (decisions of another evaluation):
    neigh[v,k] = pre_pool[arg[v,k],k] * neigh_pos[v,k];  out = (fc_self(h) + fc_neigh(neigh)) * out_pos."""
    h = features
    for layer, d in zip(ref.layers, decisions):
        pre = layer.fc_pool(h)
        arg = d["arg"]
        gathered = torch.gather(pre, 0, arg.clamp(min=0))
        neigh = gathered * (d["neigh_pos"] & (arg >= 0)).to(pre.dtype)
        rst = layer.fc_self(h) + layer.fc_neigh(neigh)
        if d["out_pos"] is not None:
            rst = rst * d["out_pos"].to(rst.dtype)
        h = rst
    return h


def count_decision_flips(a, b):
    """Number of differing decisions between two decision lists: (arg-max flips, pool-ReLU flips, output-ReLU flips, total
    decisions of each kind)."""
    n_arg = n_pool = n_out = 0
    t_arg = t_out = 0
    for x, y in zip(a, b):
        n_arg += int((x["arg"] != y["arg"]).sum())
        n_pool += int((x["neigh_pos"] != y["neigh_pos"]).sum())
        t_arg += x["arg"].numel()
        if x["out_pos"] is not None:
            n_out += int((x["out_pos"] != y["out_pos"]).sum())
            t_out += x["out_pos"].numel()
    return {"argmax_flips": n_arg, "pool_relu_flips": n_pool, "out_relu_flips": n_out,
            "argmax_decisions": t_arg, "out_relu_decisions": t_out}
