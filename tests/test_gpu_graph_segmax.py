"""GPU parity: K6 CSR/CSC build and K2/K2b neighbour max vs the oracle —
bit-exact (integer / index work)."""
import numpy as np
import pytest
import torch

from gnn_tumor_seg_b200 import graph as G, ops, synth
from oracle import graph_ref, sage_ref

pytestmark = pytest.mark.gpu


def _dev_csr(src, dst, n, dev):
    s = torch.as_tensor(np.asarray(src, np.int32)).to(dev)
    d = torch.as_tensor(np.asarray(dst, np.int32)).to(dev)
    return ops.csr_build(d, s, n, want_eid=True)


@pytest.mark.parametrize("case", ["small", "isolated", "hub", "empty_edges", "single", "dup_edges"])
def test_csr_build_bit_exact(cuda_dev, case):
    rng = np.random.default_rng(7)
    if case == "small":
        g = synth.make_small_graph(1, n_nodes=500, avg_deg=10); src, dst, n = g.src, g.dst, g.n_nodes
    elif case == "isolated":
        g = synth.make_small_graph(2, n_nodes=300, avg_deg=4, isolated=40); src, dst, n = g.src, g.dst, g.n_nodes
    elif case == "hub":       # rows far longer than a warp (rank sort's long path) and > one scan tile of nodes
        n = 9000
        src = np.concatenate([rng.integers(0, n, 3000), rng.integers(0, n, 20000)])
        dst = np.concatenate([np.full(3000, 17), rng.integers(0, n, 20000)])
        p = rng.permutation(src.size); src, dst = src[p], dst[p]
    elif case == "empty_edges":
        src, dst, n = np.zeros(0, np.int32), np.zeros(0, np.int32), 10
    elif case == "single":
        src, dst, n = np.array([0]), np.array([0]), 1
    else:                     # multigraph: duplicate (src,dst) pairs keep edge-id order
        n = 50
        src = rng.integers(0, n, 2000); dst = rng.integers(0, n, 2000)
    indptr, indices, eid = _dev_csr(src, dst, n, cuda_dev)
    r_ptr, r_idx, r_eid = graph_ref.csr_by_dst_ref(src, dst, n)
    assert np.array_equal(indptr.cpu().numpy(), r_ptr)
    assert np.array_equal(indices.cpu().numpy(), r_idx)
    if len(src):
        assert np.array_equal(eid.cpu().numpy(), r_eid)


def test_batched_graph_to_device_matches_dgl_batch_oracle(cuda_dev):
    gs = [synth.make_small_graph(s, n_nodes=200 + 31 * s, avg_deg=7, isolated=s) for s in range(4)]
    bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in gs]).to(cuda_dev)
    s, d, n, _, _ = graph_ref.batch_graphs_ref([(g.src, g.dst, g.n_nodes) for g in gs])
    r_ptr, r_idx, r_eid = graph_ref.csr_by_dst_ref(s, d, n)
    assert np.array_equal(bg.csr[0].cpu().numpy(), r_ptr) and np.array_equal(bg.csr[1].cpu().numpy(), r_idx)
    c_ptr, c_idx, c_eid = graph_ref.csc_by_src_ref(s, d, n)
    cptr, cidx, c2r = bg.csc
    assert np.array_equal(cptr.cpu().numpy(), c_ptr) and np.array_equal(cidx.cpu().numpy(), c_idx)
    pos = np.empty(len(s), np.int64); pos[r_eid] = np.arange(len(s))
    assert np.array_equal(c2r.cpu().numpy(), pos[c_eid])
    assert np.array_equal(bg.in_degrees().cpu().numpy(), np.bincount(d, minlength=n))
    assert bg.has_zero_in_degree()


def test_full_size_csr_properties(cuda_dev):
    """BASELINE config-2 size (B=6 x 15k nodes): structural properties + oracle."""
    gs = [synth.make_graph(s) for s in range(6)]
    bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in gs]).to(cuda_dev)
    indptr, indices = (t.cpu().numpy() for t in bg.csr)
    N, E = bg.number_of_nodes(), bg.number_of_edges()
    assert N == 90000 and indptr[0] == 0 and indptr[-1] == E and np.all(np.diff(indptr) >= 1)
    s, d, n, _, _ = graph_ref.batch_graphs_ref([(g.src, g.dst, g.n_nodes) for g in gs])
    r_ptr, r_idx, _ = graph_ref.csr_by_dst_ref(s, d, n)
    assert np.array_equal(indptr, r_ptr) and np.array_equal(indices, r_idx)
    rows = np.repeat(np.arange(N), np.diff(indptr))
    assert np.all((indices // 15000) == (rows // 15000))          # block diagonal: no edge crosses graphs
    eid = bg.csr_eid.cpu().numpy()
    assert np.array_equal(np.sort(eid), np.arange(E))             # a permutation of the edge ids


@pytest.mark.parametrize("D", [4, 20, 64, 256, 260, 7, 1024])
def test_segmax_fwd_values_and_argmax_bit_exact(cuda_dev, D):
    g = synth.make_small_graph(11, n_nodes=700, avg_deg=12, isolated=5)
    indptr, indices, _ = graph_ref.csr_by_dst_ref(g.src, g.dst, g.n_nodes)
    gen = torch.Generator().manual_seed(D)
    P = torch.relu(torch.randn(g.n_nodes, D, generator=gen))     # many exact-zero ties, as after fc_pool's ReLU
    P[:, 0] = 1.0                                                 # a column that ties everywhere
    r_n, r_a = sage_ref.segment_max_first_ref(P, indptr, indices)
    dn, da = ops.segmax_fwd(P.to(cuda_dev), torch.as_tensor(indptr).to(cuda_dev), torch.as_tensor(indices).to(cuda_dev))
    assert torch.equal(dn.cpu(), r_n)
    assert torch.equal(da.cpu().long(), r_a)
    dn2, none = ops.segmax_fwd(P.to(cuda_dev), torch.as_tensor(indptr).to(cuda_dev), torch.as_tensor(indices).to(cuda_dev),
                               want_argmax=False)
    assert none is None and torch.equal(dn2.cpu(), r_n)


def test_segmax_long_rows_and_strided_input(cuda_dev):
    rng = np.random.default_rng(3)
    n = 400
    src = np.concatenate([rng.integers(0, n, 150), rng.integers(0, n, 3000)])
    dst = np.concatenate([np.full(150, 5), rng.integers(0, n, 3000)])      # row 5 has > 150 entries
    indptr, indices, _ = graph_ref.csr_by_dst_ref(src, dst, n)
    big = torch.randn(n, 512)
    P = big[:, 128:384]                                                     # ld = 512, D = 256, still 16B aligned
    r_n, r_a = sage_ref.segment_max_first_ref(P.contiguous(), indptr, indices)
    dn, da = ops.segmax_fwd(big.to(cuda_dev)[:, 128:384], torch.as_tensor(indptr).to(cuda_dev),
                            torch.as_tensor(indices).to(cuda_dev))
    assert torch.equal(dn.cpu(), r_n) and torch.equal(da.cpu().long(), r_a)


@pytest.mark.parametrize("D", [20, 256, 7])
def test_segmax_bwd_scatter(cuda_dev, D):
    g = synth.make_small_graph(12, n_nodes=600, avg_deg=10, isolated=4)
    indptr, indices, _ = graph_ref.csr_by_dst_ref(g.src, g.dst, g.n_nodes)
    P = torch.relu(torch.randn(g.n_nodes, D)).double().requires_grad_(True)
    neigh, arg = sage_ref.segment_max_first(P, indptr, indices)
    gout = torch.randn(g.n_nodes, D, dtype=torch.float64)
    neigh.backward(gout)
    dP = ops.segmax_bwd(gout.float().to(cuda_dev), arg.int().to(cuda_dev), g.n_nodes)
    assert torch.allclose(dP.cpu().double(), P.grad, atol=1e-5, rtol=1e-5)
    # deterministic transposed-gather form: same numbers, and bit-identical run to run
    bg = G.from_edge_list(g.src, g.dst, g.n_nodes).to(cuda_dev)
    d1 = ops.segmax_bwd(gout.float().to(cuda_dev), arg.int().to(cuda_dev), g.n_nodes, csc=bg.csc[:2])
    d2 = ops.segmax_bwd(gout.float().to(cuda_dev), arg.int().to(cuda_dev), g.n_nodes, csc=bg.csc[:2])
    assert torch.equal(d1, d2)
    assert torch.allclose(d1.cpu().double(), P.grad, atol=1e-5, rtol=1e-5)


def test_segmax_full_size_idempotence(cuda_dev):
    """Config-2 size property: with a self-loop on every node, max-aggregating
    twice over the same features is monotone and arg-max rows point inside the
    row's neighbour list."""
    g = synth.make_graph(0)
    bg = G.from_edge_list(g.src, g.dst, g.n_nodes).to(cuda_dev)
    P = torch.relu(torch.randn(g.n_nodes, 256, device=cuda_dev))
    n1, a1 = ops.segmax_fwd(P, *bg.csr)
    assert bool((n1 >= P).all())                      # self-loop: max over a set that contains v
    assert torch.equal(P.gather(0, a1.long()), n1)    # the arg-max reproduces the max exactly
    indptr, indices = (t.cpu().numpy() for t in bg.csr)
    r_n, r_a = sage_ref.segment_max_first_ref(P.cpu(), indptr, indices)
    assert torch.equal(n1.cpu(), r_n) and torch.equal(a1.cpu().long(), r_a)


@pytest.mark.parametrize("D", [128, 256])
@pytest.mark.parametrize("mode", ["sum", "mean", "gcn"])
def test_segsum_wide_matches_scalar_and_fp64(cuda_dev, D, mode):
    """The vectorised sum/mean/gcn aggregation (forward over the in-edge CSR, backward over the out-edge CSC) against
    an fp64 index_add restatement, on a graph with isolated nodes and a 40-neighbour hub."""
    import numpy as np
    from gnn_tumor_seg_b200 import graph as G, ops
    rng = np.random.default_rng(11)
    N = 3000
    src = rng.integers(0, N - 5, size=40000); dst = rng.integers(0, N - 5, size=40000)        # last 5 nodes isolated
    src = np.concatenate([src, rng.integers(0, N - 5, size=40)]); dst = np.concatenate([dst, np.full(40, 7)])
    g = G.from_edge_list(src, dst, N).to(cuda_dev)
    X = torch.randn(N, D, generator=torch.Generator().manual_seed(D))
    s, d = torch.as_tensor(src), torch.as_tensor(dst)
    deg = torch.bincount(d, minlength=N).double().view(-1, 1)
    summed = torch.zeros(N, D, dtype=torch.float64).index_add(0, d, X.double()[s])
    ref = {"sum": summed, "mean": torch.where(deg > 0, summed / deg.clamp(min=1), torch.zeros_like(summed)),
           "gcn": (summed + X.double()) / (deg + 1)}[mode]
    indptr, indices = g.csr
    out = ops.segsum_fwd(X.to(cuda_dev), indptr, indices, mode)
    assert (out.cpu().double() - ref).abs().max() < 1e-5 * max(1.0, ref.abs().max())
    # backward = transpose of the forward operator: <out_bar, A x> == <A^T out_bar, x>
    cptr, cidx, _ = g.csc
    Gr = torch.randn(N, D, generator=torch.Generator().manual_seed(1))
    dX = ops.segsum_bwd(Gr.to(cuda_dev), cptr, cidx, indptr, mode)
    sc = {"sum": torch.ones_like(deg), "mean": torch.where(deg > 0, 1 / deg.clamp(min=1), torch.zeros_like(deg)), "gcn": 1 / (deg + 1)}[mode]
    refb = torch.zeros(N, D, dtype=torch.float64).index_add(0, s, (Gr.double() * sc)[d])
    if mode == "gcn":
        refb = refb + Gr.double() * sc
    assert (dX.cpu().double() - refb).abs().max() < 1e-5 * max(1.0, refb.abs().max())
