"""GNN architectures with the reference's signatures (model/networks.py:20-81),
built on libgts.so kernels instead of DGL.

``SAGEConv`` / ``GATConv`` mirror the constructor order, parameter names and
state-dict keys of ``dgl.nn.pytorch`` so reference
checkpoints load unchanged; ``GraphSage``, ``GAT`` and ``init_graph_net`` are the
reference's own classes/functions, same arguments, same errors.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import GtsError


class SAGEConv(nn.Module):
    """DGL ``SAGEConv(in_feats, out_feats, aggregator_type, feat_drop=0., bias=True,
    norm=None, activation=None)`` for aggregator 'pool' (the hot path), 'mean'
    and 'gcn'.

    Parameter layout follows DGL<=0.7 (``fc_self.bias`` and ``fc_neigh.bias``,
    effective bias = their sum) — the layout the shipped 7x256 weights imply.
    State dicts in the DGL 0.8-0.9 layout (separate ``bias``) or the DGL>=1.0
    layout (``fc_self.bias`` only) are converted on load.
    """

    def __init__(self, in_feats, out_feats, aggregator_type, feat_drop=0., bias=True, norm=None, activation=None):
        super().__init__()
        if aggregator_type not in ("pool", "mean", "gcn"):
            raise KeyError(f"Invalid aggregator_type. Must be one of ('mean','gcn','pool'). "
                           f"But got {aggregator_type!r} instead.")
        self._in_feats, self._out_feats = in_feats, out_feats
        self._aggre_type = aggregator_type
        self.norm = norm
        self.feat_drop = nn.Dropout(feat_drop)
        self.activation = activation
        if aggregator_type == "pool":
            self.fc_pool = nn.Linear(in_feats, in_feats)
        if aggregator_type != "gcn":
            self.fc_self = nn.Linear(in_feats, out_feats, bias=bias)
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=bias)
        self._has_bias = bias
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        if self._aggre_type == "pool":
            nn.init.xavier_uniform_(self.fc_pool.weight, gain=gain)
        if self._aggre_type != "gcn":
            nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def _effective_bias(self):
        if not self._has_bias:
            return torch.zeros(self._out_feats, dtype=torch.float32, device=self.fc_neigh.weight.device)
        if self._aggre_type == "gcn":
            return self.fc_neigh.bias
        return self.fc_self.bias + self.fc_neigh.bias

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # Accept the three DGL bias layouts (Appendix A.3): fold whatever exists into ours.
        ks, kn, kb = prefix + "fc_self.bias", prefix + "fc_neigh.bias", prefix + "bias"
        if self._has_bias and (kb in state_dict or (ks in state_dict) != (kn in state_dict)):
            parts = [state_dict[k] for k in (ks, kn, kb) if k in state_dict]
            if parts:
                eff = torch.stack([p.float() for p in parts]).sum(0)
                state_dict.pop(kb, None)
                if self._aggre_type == "gcn":
                    state_dict.pop(ks, None)
                    state_dict[kn] = eff
                else:
                    state_dict[ks] = eff
                    state_dict[kn] = torch.zeros_like(eff)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def forward(self, graph, feat, edge_weight=None, _input_is_relu=False, _grad_premasked=False):
        if edge_weight is not None:
            raise NotImplementedError("edge_weight is not used by the reference (model/networks.py) and is not supported")
        h = self.feat_drop(feat)
        fused_relu = self.activation is F.relu or self.activation is torch.relu
        b = self._effective_bias()
        if self._aggre_type == "pool":
            rst = ops.sage_pool_layer(h, self.fc_pool.weight, self.fc_pool.bias, self.fc_self.weight,
                                      self.fc_neigh.weight, b, graph, fused_relu,
                                      input_is_relu=_input_is_relu, grad_premasked=_grad_premasked and fused_relu,
                                      deterministic=ops.deterministic_backward())
        else:
            ws = self.fc_self.weight if self._aggre_type != "gcn" else None
            rst = ops.SageSumLayerFn.apply(h, ws, self.fc_neigh.weight, b, graph, self._aggre_type, fused_relu)
        if self.activation is not None and not fused_relu:
            rst = self.activation(rst)
        if self.norm is not None:
            rst = self.norm(rst)
        return rst


class GATConv(nn.Module):
    """DGL ``GATConv(in_feats, out_feats, num_heads, feat_drop=0., attn_drop=0.,
    negative_slope=0.2, residual=False, activation=None,
    allow_zero_in_degree=False, bias=True)`` (DGL>=0.7 parameter layout)."""

    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0., attn_drop=0., negative_slope=0.2,
                 residual=False, activation=None, allow_zero_in_degree=False, bias=True):
        super().__init__()
        self._num_heads, self._in_feats, self._out_feats = num_heads, in_feats, out_feats
        self._allow_zero_in_degree = allow_zero_in_degree
        self.fc = nn.Linear(in_feats, out_feats * num_heads, bias=False)
        self.attn_l = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.attn_r = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.feat_drop = nn.Dropout(feat_drop)
        if attn_drop:
            raise NotImplementedError("attn_drop > 0 is not supported (the reference fixes it at 0, model/networks.py:41)")
        self.negative_slope = negative_slope
        if bias:
            self.bias = nn.Parameter(torch.zeros(num_heads * out_feats))
        else:
            self.register_buffer("bias", None)
        if residual:
            if in_feats != out_feats * num_heads:
                self.res_fc = nn.Linear(in_feats, num_heads * out_feats, bias=False)
            else:
                self.res_fc = nn.Identity()
        else:
            self.register_buffer("res_fc", None)
        self.activation = activation
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_normal_(self.fc.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        nn.init.xavier_normal_(self.attn_r, gain=gain)
        if self.bias is not None:
            nn.init.constant_(self.bias, 0)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)

    def forward(self, graph, feat):
        if not self._allow_zero_in_degree and graph.has_zero_in_degree():
            raise GtsError("There are 0-in-degree nodes in the graph, output for those nodes will be invalid. "
                           "Add self-loops or set allow_zero_in_degree=True.")   # DGLError in DGL
        h = self.feat_drop(feat)
        fused_elu = self.activation is F.elu
        wres = self.res_fc.weight if isinstance(self.res_fc, nn.Linear) else None
        rst = ops.GatLayerFn.apply(h, self.fc.weight, self.attn_l, self.attn_r, self.bias, wres, graph,
                                   self._num_heads, self._out_feats, self.negative_slope,
                                   isinstance(self.res_fc, nn.Identity), fused_elu)
        if self.activation is not None and not fused_elu:
            rst = self.activation(rst)
        return rst


class GraphSage(nn.Module):
    """reference model/networks.py:20-36."""

    def __init__(self, in_feats, layer_sizes, n_classes, aggregator_type, dropout):
        super().__init__()
        self.layers = nn.ModuleList()
        # input layer
        self.layers.append(SAGEConv(in_feats, layer_sizes[0], aggregator_type, feat_drop=dropout, activation=F.relu))
        # hidden layers
        for i in range(1, len(layer_sizes)):
            self.layers.append(SAGEConv(layer_sizes[i - 1], layer_sizes[i], aggregator_type, feat_drop=dropout,
                                        activation=F.relu))
        # output layer
        self.layers.append(SAGEConv(layer_sizes[-1], n_classes, aggregator_type, feat_drop=0, activation=None))

    def _stack_fast_path_ok(self):
        # the whole-stack backward takes dlogits as the gradient of the last layer's OUTPUT: an activation there (never
        # in the reference, model/networks.py:30) goes through the per-layer path, which applies its mask
        return all(l._aggre_type == "pool" and l.norm is None
                   and (l.feat_drop.p == 0 or not self.training)
                   and (l.activation is None or l.activation is F.relu or l.activation is torch.relu)
                   for l in self.layers) and self.layers[-1].activation is None

    def forward(self, graph, features):
        if self._stack_fast_path_ok() and ops.use_stack_path():
            # whole stack in one library call per direction (gts_sage_forward / gts_sage_backward)
            flat, relus = [], []
            for l in self.layers:
                flat += [l.fc_pool.weight, l.fc_pool.bias, l.fc_self.weight, l.fc_self.bias, l.fc_neigh.weight,
                         l.fc_neigh.bias]
                relus.append(l.activation is not None)
            return ops.SageStackFn.apply(graph, features, tuple(relus), ops.deterministic_backward(), *flat)
        h = features
        n = len(self.layers)
        for i, layer in enumerate(self.layers):
            # the stack is strictly sequential, so each layer may hand its input
            # gradient back already masked by the producer's ReLU (fused epilogue)
            no_drop = layer.feat_drop.p == 0 or not self.training
            input_is_relu = i > 0 and no_drop and layer._aggre_type == "pool"
            nxt = self.layers[i + 1] if i + 1 < n else None
            premasked = (nxt is not None and nxt._aggre_type == "pool"
                         and (nxt.feat_drop.p == 0 or not self.training))
            h = layer(graph, h, _input_is_relu=input_is_relu, _grad_premasked=premasked)
        return h


class GAT(nn.Module):
    """reference model/networks.py:39-66."""

    def __init__(self, in_feats, layer_sizes, n_classes, heads, residuals,
                 activation=F.elu, feat_drop=0, attn_drop=0, negative_slope=0.2):
        super().__init__()
        self.layers = nn.ModuleList()
        self.activation = activation
        # input projection (no residual)
        self.layers.append(GATConv(in_feats, layer_sizes[0], heads[0],
                                   feat_drop, attn_drop, negative_slope, False, self.activation))
        # hidden layers
        for i in range(1, len(layer_sizes)):
            # due to multi-head, the in_dim = num_hidden * num_heads
            self.layers.append(GATConv(layer_sizes[i - 1] * heads[i - 1], layer_sizes[i], heads[i],
                                       feat_drop, attn_drop, negative_slope, residuals[i], self.activation))
        # output projection
        self.layers.append(GATConv(layer_sizes[-1] * heads[-1], n_classes, 1,
                                   feat_drop, attn_drop, negative_slope, False, None))

    def forward(self, g, inputs):
        h = inputs
        for l in range(len(self.layers) - 1):
            h = self.layers[l](g, h).flatten(1)
        # output projection
        logits = self.layers[-1](g, h).mean(1)
        return logits


def init_graph_net(model_type, hp):
    """reference model/networks.py:68-81 — same model types, same namedtuple
    fields, same exception."""
    dropout = hp.feature_dropout if 'feature_dropout' in hp._fields else 0
    if model_type == 'GSpool':
        net = GraphSage(in_feats=hp.in_feats, layer_sizes=hp.layer_sizes, n_classes=hp.out_classes,
                        aggregator_type='pool', dropout=dropout)
    elif model_type == 'GSgcn':
        net = GraphSage(in_feats=hp.in_feats, layer_sizes=hp.layer_sizes, n_classes=hp.out_classes,
                        aggregator_type='gcn', dropout=dropout)
    elif model_type == 'GSmean':
        net = GraphSage(in_feats=hp.in_feats, layer_sizes=hp.layer_sizes, n_classes=hp.out_classes,
                        aggregator_type='mean', dropout=dropout)
    elif model_type == 'GAT':
        net = GAT(in_feats=hp.in_feats, layer_sizes=hp.layer_sizes, n_classes=hp.out_classes,
                  heads=hp.gat_heads, residuals=hp.gat_residuals)
    else:
        raise Exception(f"Unknown model type: {model_type}")
    return net
