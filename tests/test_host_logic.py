"""Host-side logic that runs without a GPU: graph batching order, module
construction / state-dict compatibility, factory errors, loud failure of the
compute path on CPU."""
import os
from collections import namedtuple

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gnn_tumor_seg_b200 import graph as G, networks, synth
from gnn_tumor_seg_b200._lib import GtsError
from oracle import graph_ref, sage_ref, gat_ref

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_kat.npz"))
FullParamSet = namedtuple("FullParamSet", 'n_epochs in_feats out_classes lr lr_decay w_decay class_weights layer_sizes feature_dropout gat_heads gat_residuals')
EvalParamSet = namedtuple("EvalParamSet", 'in_feats out_classes layer_sizes gat_heads gat_residuals')


def test_from_networkx_edge_order_matches_reference_fixture():
    import networkx as nx
    n3 = int(GOLD["n3"])
    nxg = nx.Graph()
    nxg.add_nodes_from(range(n3))
    nxg.add_edges_from(zip(GOLD["r3"].tolist(), GOLD["c3"].tolist()))
    g = G.from_networkx(nxg)
    s, d = g.edges()
    assert np.array_equal(np.stack([s.numpy(), d.numpy()], 1), GOLD["nx_edges"])
    assert g.number_of_nodes() == n3 and g.number_of_edges() == len(GOLD["r3"])
    assert np.array_equal(g.in_degrees().numpy(), np.bincount(GOLD["c3"], minlength=n3))


def test_batch_matches_dgl_batch_semantics():
    gs = [synth.make_small_graph(s, n_nodes=20 + s) for s in range(3)]
    b = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in gs])
    s, d, n, noff, eoff = graph_ref.batch_graphs_ref([(g.src, g.dst, g.n_nodes) for g in gs])
    bs, bd = b.edges()
    assert np.array_equal(bs.numpy(), s) and np.array_equal(bd.numpy(), d)
    assert b.number_of_nodes() == n and b.batch_size == 3
    assert b.batch_num_nodes().tolist() == [20, 21, 22]
    # nested batch keeps per-graph bookkeeping
    bb = G.batch([b, G.from_edge_list(gs[0].src, gs[0].dst, gs[0].n_nodes)])
    assert bb.batch_size == 4 and bb.number_of_nodes() == n + 20


def test_minibatch_graphs_signature():
    gs = [synth.make_small_graph(s, n_nodes=10) for s in range(2)]
    samples = [(g.mri_id, G.from_edge_list(g.src, g.dst, g.n_nodes), g.features.astype(np.float64), g.labels) for g in gs]
    ids, bg, feats, labels = G.minibatch_graphs(samples)
    assert ids == [g.mri_id for g in gs]
    assert feats.dtype == torch.float32 and feats.shape == (20, 20)
    assert labels.dtype == torch.int64 and labels.shape == (20,)
    assert bg.number_of_nodes() == 20


def test_from_edge_list_validates():
    with pytest.raises(ValueError):
        G.from_edge_list([0, 5], [1, 1], 3)
    with pytest.raises(ValueError):
        G.from_edge_list([0], [1, 1], 3)


def test_graphsage_7x256_parameter_count_and_keys():
    net = networks.GraphSage(20, [256] * 7, 4, "pool", 0)
    assert len(net.layers) == 8
    assert sum(p.numel() for p in net.parameters()) == 1263276
    keys = set(net.state_dict())
    for i in range(8):
        for k in ("fc_pool.weight", "fc_pool.bias", "fc_self.weight", "fc_self.bias", "fc_neigh.weight", "fc_neigh.bias"):
            assert f"layers.{i}.{k}" in keys
    assert net.layers[0].fc_pool.weight.shape == (20, 20) and net.layers[7].fc_self.weight.shape == (4, 256)
    # same keys/shapes as the oracle restatement -> state dicts interchange
    ref = sage_ref.GraphSageRef(20, [256] * 7, 4)
    assert {k: v.shape for k, v in ref.state_dict().items()} == {k: v.shape for k, v in net.state_dict().items()}


def test_state_dict_bias_layouts_convert():
    net = networks.GraphSage(20, [8], 4, "pool", 0)
    sd = net.state_dict()
    # DGL 0.8-0.9: separate `bias`, bias-free linears
    sd89 = {k: v.clone() for k, v in sd.items() if not k.endswith(("fc_self.bias", "fc_neigh.bias"))}
    for i in range(2):
        sd89[f"layers.{i}.bias"] = torch.full((net.layers[i]._out_feats,), 0.5 + i)
    net.load_state_dict(sd89)
    assert torch.allclose(net.layers[1]._effective_bias(), torch.full((4,), 1.5))
    # DGL >= 1.0: fc_self.bias only
    sd10 = {k: v.clone() for k, v in sd.items() if not k.endswith("fc_neigh.bias")}
    sd10["layers.0.fc_self.bias"] = torch.full((8,), 3.0)
    net.load_state_dict(sd10)
    assert torch.allclose(net.layers[0]._effective_bias(), torch.full((8,), 3.0))
    # DGL <= 0.7 (native layout) round trip
    net.load_state_dict(sd)
    assert torch.allclose(net.layers[0]._effective_bias(), sd["layers.0.fc_self.bias"] + sd["layers.0.fc_neigh.bias"])


def test_gat_shapes_and_keys():
    net = networks.GAT(20, [256] * 4, 4, [4, 4, 4, 4], [False, False, True, False])
    assert sum(p.numel() for p in net.parameters()) == 3182604
    assert isinstance(net.layers[2].res_fc, torch.nn.Identity) and net.layers[1].res_fc is None
    ref = gat_ref.GATRef(20, [256] * 4, 4, [4, 4, 4, 4], [False, False, True, False])
    assert {k: v.shape for k, v in ref.state_dict().items()} == {k: v.shape for k, v in net.state_dict().items()}
    assert net.layers[0].attn_l.shape == (1, 4, 256) and net.layers[4].fc.weight.shape == (4, 1024)


def test_init_graph_net_factory():
    hp = EvalParamSet(20, 4, [16] * 2, [2, 2], [False, True])
    assert isinstance(networks.init_graph_net("GSpool", hp), networks.GraphSage)
    assert networks.init_graph_net("GSmean", hp).layers[0]._aggre_type == "mean"
    assert networks.init_graph_net("GSgcn", hp).layers[0]._aggre_type == "gcn"
    assert isinstance(networks.init_graph_net("GAT", hp), networks.GAT)
    with pytest.raises(Exception, match="Unknown model type"):
        networks.init_graph_net("GSlstm", hp)
    full = FullParamSet(10, 20, 4, 1e-4, 0.98, 1e-4, [0.1, 1, 2, 2], [16], 0.25, None, None)
    assert networks.init_graph_net("GSpool", full).layers[0].feat_drop.p == 0.25
    assert networks.init_graph_net("GSpool", full).layers[-1].feat_drop.p == 0         # networks.py:30


def test_reference_default_gat_hyperparameters_shape_bug_is_preserved():
    # hyperparam_helpers.py:39-42: 4 layer sizes with a 6-long heads list -> the output layer is
    # sized with heads[-1] while the last hidden layer emits heads[3]
    net = networks.GAT(20, [256] * 4, 4, [4, 4, 3, 3, 4, 4], [False, False, True, False, False, True])
    assert net.layers[3].fc.weight.shape[0] == 3 * 256 and net.layers[4].fc.weight.shape[1] == 4 * 256


def test_compute_path_fails_loudly_on_cpu():
    net = networks.GraphSage(20, [8], 4, "pool", 0)
    g = synth.make_small_graph(0, n_nodes=10)
    hg = G.from_edge_list(g.src, g.dst, g.n_nodes)
    with pytest.raises(GtsError):
        net(hg, torch.as_tensor(g.features))        # CPU tensors: no fallback
    with pytest.raises(GtsError):
        hg.csr                                      # host graph has no device CSR


def test_product_package_never_imports_oracle():
    import re
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gnn-tumor-seg_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn


def test_overlap_helpers_fail_loudly_without_cuda():
    """DevicePrefetcher / VolumeDownloader are CUDA-stream helpers: no silent CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from gnn_tumor_seg_b200 import project
    from gnn_tumor_seg_b200._lib import GtsError
    from gnn_tumor_seg_b200.data_loader import DevicePrefetcher
    with pytest.raises(GtsError):
        DevicePrefetcher([], "cuda")
    with pytest.raises(GtsError):
        project.VolumeDownloader()


def test_has_duplicate_edges():
    import numpy as np
    from gnn_tumor_seg_b200 import graph as G
    a = G.from_edge_list(np.array([0, 1, 1, 2]), np.array([1, 0, 2, 1]), 3)
    b = G.from_edge_list(np.array([0, 1, 0, 2]), np.array([1, 0, 1, 1]), 3)
    assert not a.has_duplicate_edges() and b.has_duplicate_edges()
    assert not G.batch([a, a]).has_duplicate_edges() and G.batch([a, b]).has_duplicate_edges()
