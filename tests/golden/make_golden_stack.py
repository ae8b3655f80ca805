"""Generate tests/golden/reference_stack.npz by RUNNING THE REFERENCE'S OWN model/networks.py.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_stack.py
The fixture is committed; nothing at test/bench time reads /root/reference.

``model/networks.py`` imports ``dgl.nn.pytorch.{GATConv, GraphConv}`` and ``dgl.nn.pytorch.conv.SAGEConv``; DGL cannot
be installed here.  The three classes are stubbed by thin adapters with DGL's constructor signatures over the oracle's
PER-LAYER modules (oracle/sage_ref.py: SAGEConvPoolRef, oracle/gat_ref.py: GATConvRef), and then the reference's own
``GraphSage`` / ``GAT`` / ``init_graph_net`` (model/networks.py:20-81) build and run the stacks.  What this pins to the
reference's code (not to our reading of it): the layer wiring — input / hidden / output dimensions, which layers get
the activation and the feature dropout, ``heads[i]`` / ``residuals[i]`` indexing, ``flatten(1)`` between GAT layers and
``mean(1)`` over the output heads — and ``init_graph_net``'s hyper-parameter handling.  What it does NOT pin is the
arithmetic inside one SAGEConv / GATConv (still the oracle's restatement of DGL: "parity unpinned").
"""
import collections
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(OUT))


def install_dgl_stub(record):
    """dgl.nn.pytorch.{GATConv, GraphConv}, dgl.nn.pytorch.conv.SAGEConv over the oracle's per-layer modules.
    Every construction is appended to ``record`` with the arguments the reference passed."""
    sys.path.insert(0, ROOT)
    from oracle.gat_ref import GATConvRef
    from oracle.sage_ref import SAGEConvPoolRef

    class SAGEConv(SAGEConvPoolRef):                      # DGL: SAGEConv(in, out, aggregator_type, feat_drop, bias, norm, activation)
        def __init__(self, in_feats, out_feats, aggregator_type, feat_drop=0., bias=True, norm=None, activation=None):
            super().__init__(in_feats, out_feats, activation)
            self.feat_drop = nn.Dropout(feat_drop)
            record.append(("SAGEConv", in_feats, out_feats, aggregator_type, float(feat_drop),
                           None if activation is None else activation.__name__))

        def forward(self, graph, h):
            return super().forward(graph, self.feat_drop(h))

    class GATConv(GATConvRef):                            # DGL: GATConv(in, out, heads, feat_drop, attn_drop, slope, residual, activation)
        def __init__(self, in_feats, out_feats, num_heads, feat_drop=0., attn_drop=0., negative_slope=0.2,
                     residual=False, activation=None):
            super().__init__(in_feats, out_feats, num_heads, negative_slope, residual, activation)
            record.append(("GATConv", in_feats, out_feats, num_heads, float(feat_drop), float(attn_drop),
                           float(negative_slope), bool(residual), None if activation is None else activation.__name__))

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    conv = mod("dgl.nn.pytorch.conv", SAGEConv=SAGEConv)
    pt = mod("dgl.nn.pytorch", GATConv=GATConv, GraphConv=object, SAGEConv=SAGEConv, conv=conv)
    nn_ = mod("dgl.nn", pytorch=pt)
    mod("dgl", nn=nn_)


def small_graph(rng, n):
    """Symmetric random graph with a self-loop on every node (what find_adjacent_nodes produces), lexicographic (src, dst)."""
    a = rng.random((n, n)) < 6.0 / n                     # ~13 neighbours per node incl. the self-loop, like a RAG
    a = a | a.T | np.eye(n, dtype=bool)
    src, dst = np.nonzero(a)
    return src.astype(np.int64), dst.astype(np.int64)


def csr_by_dst(src, dst, n):
    order = np.argsort(dst, kind="stable")
    indptr = np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=n))]).astype(np.int64)
    return indptr, src[order].astype(np.int64)


def main():
    record = []
    install_dgl_stub(record)
    sys.path.insert(0, REF)
    from model import networks as ref_networks                  # the reference's own file, unmodified

    rng = np.random.default_rng(20211019)
    n = 300
    src, dst = small_graph(rng, n)
    indptr, indices = csr_by_dst(src, dst, n)
    x = rng.standard_normal((n, 20)).astype(np.float32)
    out = {"src": src, "dst": dst, "n_nodes": np.int64(n), "x": x}

    # ---- GraphSage (model/networks.py:20-36) ----
    torch.manual_seed(1)
    record.clear()
    sage = ref_networks.GraphSage(20, [64, 32, 128], 4, "pool", 0.0)
    out["sage_ctor"] = np.array([repr(r) for r in record])
    with torch.no_grad():
        out["sage_logits"] = sage((indptr, indices), torch.as_tensor(x)).numpy()
    for k, v in sage.state_dict().items():
        out["sage_sd/" + k] = v.numpy()

    # ---- GAT (model/networks.py:39-66); residuals[0] is given as True and ignored by the reference (input projection) ----
    torch.manual_seed(2)
    record.clear()
    # layer 1: 32 -> 4 x 8 = 32 (identity residual), layer 2: 32 -> 2 x 32 = 64 (linear residual)
    gat = ref_networks.GAT(20, [16, 8, 32], 4, [2, 4, 2], [True, True, True])
    with torch.no_grad():
        for l in gat.layers:
            l.bias.normal_(std=0.1)                      # DGL initialises the bias to 0: make it count
    out["gat_ctor"] = np.array([repr(r) for r in record])
    # heads / residuals lists longer than layer_sizes, like the reference's default hyper-parameters
    # (utils/hyperparam_helpers.py:39-42): the output layer is sized with heads[-1] while the last hidden layer emits
    # heads[len(layer_sizes)-1] heads — construction succeeds, forward raises a shape error.  Recorded, not "fixed".
    record.clear()
    bad = ref_networks.GAT(20, [16, 8, 32], 4, [2, 4, 2, 5], [True, True, True, True])
    out["gat_ctor_long_lists"] = np.array([repr(r) for r in record])
    try:
        with torch.no_grad():
            bad((torch.as_tensor(src), torch.as_tensor(dst)), torch.as_tensor(x))
        out["gat_long_lists_forward"] = np.array(["ok"])
    except RuntimeError as e:
        out["gat_long_lists_forward"] = np.array(["RuntimeError: " + str(e).splitlines()[0]])
    with torch.no_grad():
        out["gat_logits"] = gat((torch.as_tensor(src), torch.as_tensor(dst)), torch.as_tensor(x)).numpy()
    for k, v in gat.state_dict().items():
        out["gat_sd/" + k] = v.numpy()

    # ---- init_graph_net (model/networks.py:68-81): what it builds for the two hyper-parameter tuple layouts ----
    Eval = collections.namedtuple("EvalParamSet", ["in_feats", "out_classes", "layer_sizes", "gat_heads", "gat_residuals"])
    Full = collections.namedtuple("FullParamSet", ["n_epochs", "in_feats", "out_classes", "lr", "lr_decay", "weight_decay",
                                                   "class_weights", "layer_sizes", "feature_dropout", "gat_heads", "gat_residuals"])
    cases = [("GSpool", Eval(20, 4, [16, 12], [2, 2], [False, True])),
             ("GSmean", Eval(20, 4, [16, 12], [2, 2], [False, True])),
             ("GSgcn", Eval(20, 4, [16], None, None)),
             ("GAT", Eval(20, 4, [16, 12], [2, 3], [False, True])),
             ("GSpool", Full(10, 20, 4, 1e-4, 0.98, 1e-4, [0.1, 1, 2, 2], [16, 16, 8], 0.25, None, None)),
             ("GAT", Full(10, 20, 4, 1e-4, 0.98, 1e-4, [0.1, 1, 2, 2], [8, 8], 0.25, [4, 4], [False, False]))]
    rows = []
    for model_type, hp in cases:
        record.clear()
        net = ref_networks.init_graph_net(model_type, hp)
        rows.append(repr((model_type, type(hp).__name__, tuple(hp), type(net).__name__, list(record))))
    try:
        ref_networks.init_graph_net("GSlstm2", cases[0][1])
    except Exception as e:
        rows.append(repr(("error", type(e).__name__, str(e))))
    out["init_graph_net"] = np.array(rows)
    np.savez_compressed(os.path.join(OUT, "reference_stack.npz"), **out)
    print("wrote reference_stack.npz:", len(out), "arrays;", "sage logits", out["sage_logits"].shape, "gat logits", out["gat_logits"].shape)
    for r in out["sage_ctor"]:
        print("  ", r)
    for r in out["gat_ctor"]:
        print("  ", r)


if __name__ == "__main__":
    main()
