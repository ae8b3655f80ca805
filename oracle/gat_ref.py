"""Oracle: DGL ``GATConv`` and the reference GAT stack.

TEST INFRASTRUCTURE (see oracle/__init__.py).  **Parity unpinned** (DGL not
installable); follows SURVEY.md Appendix A.2 and model/networks.py:39-66.
Pure PyTorch (CPU, fp32/fp64); backward is torch autograd over these ops.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class GATConvRef(nn.Module):
    """GATConv(in, F, H, feat_drop=0, attn_drop=0, negative_slope, residual,
    activation, bias=True) — DGL>=0.7 layout (``bias`` [H*F]; ``res_fc`` is a
    bias-free Linear when in != H*F, Identity when equal)."""

    def __init__(self, in_feats, out_feats, num_heads, negative_slope=0.2,
                 residual=False, activation=None):
        super().__init__()
        self.H, self.F = num_heads, out_feats
        self.fc = nn.Linear(in_feats, out_feats * num_heads, bias=False)
        self.attn_l = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.attn_r = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.bias = nn.Parameter(torch.zeros(num_heads * out_feats))
        if residual:
            if in_feats != out_feats * num_heads:
                self.res_fc = nn.Linear(in_feats, num_heads * out_feats, bias=False)
            else:
                self.res_fc = nn.Identity()
        else:
            self.res_fc = None
        self.negative_slope = negative_slope
        self.activation = activation
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_normal_(self.fc.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        nn.init.xavier_normal_(self.attn_r, gain=gain)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)

    def forward(self, edges, x):
        """edges = (src, dst) int64 tensors, N = x.shape[0]."""
        src, dst = edges
        N = x.shape[0]
        if N > 0 and torch.bincount(dst, minlength=N).min() == 0:
            raise RuntimeError("There are 0-in-degree nodes in the graph")   # DGLError in DGL
        Z = self.fc(x).view(N, self.H, self.F)
        el = (Z * self.attn_l).sum(-1)              # [N,H]
        er = (Z * self.attn_r).sum(-1)
        e = F.leaky_relu(el[src] + er[dst], self.negative_slope)        # [E,H]
        m = torch.full((N, self.H), float("-inf"), dtype=x.dtype)
        m = m.scatter_reduce(0, dst.view(-1, 1).expand(-1, self.H), e, reduce="amax", include_self=True)
        a = torch.exp(e - m[dst])
        l = torch.zeros((N, self.H), dtype=x.dtype).index_add(0, dst, a)
        alpha = a / l[dst]
        out = torch.zeros((N, self.H, self.F), dtype=x.dtype)
        for h in range(self.H):        # per head to bound the E x F temporary
            out[:, h, :] = torch.zeros((N, self.F), dtype=x.dtype).index_add(
                0, dst, alpha[:, h:h + 1] * Z[src, h, :])
        if self.res_fc is not None:
            out = out + self.res_fc(x).view(N, self.H, self.F)
        out = out + self.bias.view(1, self.H, self.F)
        if self.activation is not None:
            out = self.activation(out)
        return out


class GATRef(nn.Module):
    """model/networks.py:39-66 (feat_drop = attn_drop = 0, activation ELU)."""

    def __init__(self, in_feats, layer_sizes, n_classes, heads, residuals,
                 activation=F.elu, negative_slope=0.2):
        super().__init__()
        self.layers = nn.ModuleList()
        self.layers.append(GATConvRef(in_feats, layer_sizes[0], heads[0], negative_slope, False, activation))
        for i in range(1, len(layer_sizes)):
            self.layers.append(GATConvRef(layer_sizes[i - 1] * heads[i - 1], layer_sizes[i], heads[i],
                                          negative_slope, residuals[i], activation))
        self.layers.append(GATConvRef(layer_sizes[-1] * heads[-1], n_classes, 1, negative_slope, False, None))

    def forward(self, edges, inputs):
        h = inputs
        for l in range(len(self.layers) - 1):
            h = self.layers[l](edges, h).flatten(1)
        return self.layers[-1](edges, h).mean(1)
