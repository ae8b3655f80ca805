"""Oracle: graph construction order, batching and canonical CSR (numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import numpy as np


def edges_from_networkx(nx_graph):
    """Directed edge list in the order ``dgl.from_networkx`` assigns edge ids
    (data_processing/data_loader.py:72): ``to_directed()`` of the undirected
    graph, edges in adjacency iteration order; self-loops once."""
    import networkx as nx
    g = nx_graph.to_directed() if not nx_graph.is_directed() else nx_graph
    nodes = sorted(g.nodes())
    remap = {n: i for i, n in enumerate(nodes)}
    e = [(remap[u], remap[v]) for u, v in g.edges()]
    if len(e) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), len(nodes)
    e = np.asarray(e, dtype=np.int64)
    return e[:, 0], e[:, 1], len(nodes)


def batch_graphs_ref(graphs):
    """``dgl.batch`` (data_loader.py:168): node ids of graph j offset by the
    node count of graphs < j; edge lists concatenated in sample order.
    ``graphs`` = iterable of (src, dst, n_nodes).  Returns global
    (src, dst, N, node_offsets, edge_offsets)."""
    srcs, dsts, noff, eoff = [], [], [0], [0]
    for s, d, n in graphs:
        srcs.append(np.asarray(s, np.int64) + noff[-1])
        dsts.append(np.asarray(d, np.int64) + noff[-1])
        noff.append(noff[-1] + int(n))
        eoff.append(eoff[-1] + len(s))
    src = np.concatenate(srcs) if srcs else np.zeros(0, np.int64)
    dst = np.concatenate(dsts) if dsts else np.zeros(0, np.int64)
    return src, dst, noff[-1], np.asarray(noff, np.int64), np.asarray(eoff, np.int64)


def csr_by_dst_ref(src, dst, n_nodes):
    """Canonical CSR by destination: row v lists the sources of v's in-edges
    ordered by edge id (stable sort — SURVEY.md Appendix A.4)."""
    src = np.asarray(src, np.int64)
    dst = np.asarray(dst, np.int64)
    order = np.argsort(dst, kind="stable")
    indptr = np.zeros(n_nodes + 1, np.int32)
    np.cumsum(np.bincount(dst, minlength=n_nodes), out=indptr[1:])
    return indptr, src[order].astype(np.int32), order.astype(np.int32)


def csc_by_src_ref(src, dst, n_nodes):
    """Out-edge view: row u lists the destinations of u's out-edges by edge id."""
    return csr_by_dst_ref(dst, src, n_nodes)
