"""Oracle self-consistency (DGL semantics are unpinned — see oracle/__init__.py):
independent dense formulation, hand-computed tie case, fp64 gradcheck."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import gat_ref, graph_ref, loss_ref, sage_ref
from gnn_tumor_seg_b200 import synth


def _csr(g):
    return graph_ref.csr_by_dst_ref(g.src, g.dst, g.n_nodes)[:2]


def test_csr_hand_case():
    # path 0-1-2 with self loops: SURVEY §8c-3 edge order
    src = np.array([0, 0, 1, 1, 1, 2, 2]); dst = np.array([0, 1, 0, 1, 2, 1, 2])
    indptr, indices, order = graph_ref.csr_by_dst_ref(src, dst, 3)
    assert indptr.tolist() == [0, 2, 5, 7]
    assert indices.tolist() == [0, 1, 0, 1, 2, 1, 2]
    assert order.tolist() == [0, 2, 1, 3, 5, 4, 6]


def test_batch_offsets():
    a = (np.array([0, 1]), np.array([1, 0]), 2)
    b = (np.array([0, 2]), np.array([2, 1]), 3)
    s, d, n, noff, eoff = graph_ref.batch_graphs_ref([a, b])
    assert n == 5 and s.tolist() == [0, 1, 2, 4] and d.tolist() == [1, 0, 4, 3]
    assert noff.tolist() == [0, 2, 5] and eoff.tolist() == [0, 2, 4]


def test_segmax_first_wins_and_zero_degree():
    P = torch.tensor([[1., 0., 5.], [1., 0., 2.], [0., 0., 9.], [7., 7., 7.]])
    # node 0 <- {1, 0}; node 1 <- {0,1,2}; node 2 <- {}; node 3 <- {2}
    indptr = np.array([0, 2, 5, 5, 6]); indices = np.array([1, 0, 0, 1, 2, 2])
    neigh, arg = sage_ref.segment_max_first_ref(P, indptr, indices)
    assert neigh.tolist() == [[1., 0., 5.], [1., 0., 9.], [0., 0., 0.], [0., 0., 9.]]
    assert arg.tolist() == [[1, 1, 0], [0, 0, 2], [-1, -1, -1], [2, 2, 2]]     # ties -> first in CSR order


def test_sage_matches_dense_formulation():
    g = synth.make_small_graph(3, n_nodes=60, isolated=3)
    csr = _csr(g)
    torch.manual_seed(1)
    layer = sage_ref.SAGEConvPoolRef(20, 12, F.relu).double()
    x = torch.tensor(g.features, dtype=torch.float64)
    adj = torch.zeros(60, 60, dtype=torch.bool)
    adj[torch.as_tensor(g.dst.astype(np.int64)), torch.as_tensor(g.src.astype(np.int64))] = True
    dense = sage_ref.sage_pool_dense_ref(x, adj, layer.fc_pool.weight, layer.fc_pool.bias, layer.fc_self.weight,
                                         layer.fc_neigh.weight, layer.fc_self.bias + layer.fc_neigh.bias, True)
    assert torch.allclose(layer(csr, x), dense, atol=1e-12)


def test_sage_gradcheck_fp64():
    g = synth.make_small_graph(4, n_nodes=12, avg_deg=3, in_feats=5)
    csr = _csr(g)
    torch.manual_seed(2)
    net = sage_ref.GraphSageRef(5, [6], 3).double()
    x = torch.randn(12, 5, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda t: net(csr, t), (x,), eps=1e-6, atol=1e-5)
    # parameter gradients vs dense formulation autograd
    adj = torch.zeros(12, 12, dtype=torch.bool)
    adj[torch.as_tensor(g.dst.astype(np.int64)), torch.as_tensor(g.src.astype(np.int64))] = True
    l0 = net.layers[0]
    out = l0(csr, x.detach()).sum()
    g1 = torch.autograd.grad(out, list(l0.parameters()))
    dense = sage_ref.sage_pool_dense_ref(x.detach(), adj, l0.fc_pool.weight, l0.fc_pool.bias, l0.fc_self.weight,
                                         l0.fc_neigh.weight, l0.fc_self.bias + l0.fc_neigh.bias, True).sum()
    g2 = torch.autograd.grad(dense, list(l0.parameters()))
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, atol=1e-10)


def test_gat_dense_softmax_and_gradcheck():
    g = synth.make_small_graph(5, n_nodes=10, avg_deg=3, in_feats=4)
    src = torch.as_tensor(g.src.astype(np.int64)); dst = torch.as_tensor(g.dst.astype(np.int64))
    torch.manual_seed(3)
    layer = gat_ref.GATConvRef(4, 3, 2, 0.2, residual=True, activation=F.elu).double()
    with torch.no_grad():
        layer.bias.normal_()
    x = torch.randn(10, 4, dtype=torch.float64, requires_grad=True)
    out = layer((src, dst), x)
    # dense formulation
    Z = layer.fc(x).view(10, 2, 3)
    el = (Z * layer.attn_l).sum(-1); er = (Z * layer.attn_r).sum(-1)
    S = F.leaky_relu(el.unsqueeze(0) + er.unsqueeze(1), 0.2)        # [v,u,h]
    adj = torch.zeros(10, 10, dtype=torch.bool); adj[dst, src] = True
    S = S.masked_fill(~adj.unsqueeze(-1), float("-inf"))
    A = torch.softmax(S, dim=1)
    O = torch.einsum("vuh,uhf->vhf", A, Z) + layer.res_fc(x).view(10, 2, 3) + layer.bias.view(1, 2, 3)
    assert torch.allclose(out, F.elu(O), atol=1e-12)
    assert torch.autograd.gradcheck(lambda t: layer((src, dst), t), (x,), eps=1e-6, atol=1e-5)


def test_gat_zero_in_degree_raises():
    layer = gat_ref.GATConvRef(4, 3, 1)
    import pytest
    with pytest.raises(RuntimeError, match="0-in-degree"):
        layer((torch.tensor([0]), torch.tensor([1])), torch.randn(3, 4))


def test_weighted_ce_matches_torch():
    torch.manual_seed(0)
    z = torch.randn(50, 4); y = torch.randint(0, 4, (50,)); w = torch.tensor([0.1, 1., 2., 2.])
    assert torch.allclose(loss_ref.weighted_ce_ref(z, y, w), F.cross_entropy(z, y, weight=w), atol=1e-6)
