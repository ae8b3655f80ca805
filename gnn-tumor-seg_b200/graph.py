"""Batched supervoxel graphs as device-resident int32 CSR/CSC.

Replaces the DGLGraph objects the reference moves around
(data_processing/data_loader.py:67-83 ``get_graph``; :165-169
``minibatch_graphs``; model/gnn_model.py:38 ``.to(device)``).  The reference's
host code only ever calls ``.to(device)`` on the batched graph (plus
``number_of_edges()``, ``in_degrees()`` and ``ndata[...] =`` inside the dead
'norm' block of get_graph), so that is the surface kept.

Host side: per-graph LOCAL edge lists (int32) concatenated in sample order plus
node/edge offsets — exactly what dgl.batch would union.  ``.to('cuda')`` copies
the edge lists once (pinned when possible) and builds, on the device,
* the in-edge CSR (row = destination, entries ordered by edge id), and lazily
* the out-edge CSC + CSC->CSR edge map (GAT backward / deterministic max backward).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class BatchedGraph:
    def __init__(self, src_local, dst_local, node_counts, edge_counts, pin=False):
        self._src = torch.as_tensor(np.ascontiguousarray(src_local, dtype=np.int32))
        self._dst = torch.as_tensor(np.ascontiguousarray(dst_local, dtype=np.int32))
        self._node_counts = [int(n) for n in node_counts]
        self._edge_counts = [int(e) for e in edge_counts]
        self._node_off = torch.as_tensor(np.concatenate([[0], np.cumsum(self._node_counts)]).astype(np.int32))
        self._edge_off = torch.as_tensor(np.concatenate([[0], np.cumsum(self._edge_counts)]).astype(np.int64))
        if pin and torch.cuda.is_available():
            self._src, self._dst = self._src.pin_memory(), self._dst.pin_memory()
            self._node_off, self._edge_off = self._node_off.pin_memory(), self._edge_off.pin_memory()
        self.device = torch.device("cpu")
        self.ndata = {}
        self._csr = None
        self._csc = None
        self._gsrc = self._gdst = self._eid = None
        self.err_flag = None

    # ---- DGLGraph-like queries -------------------------------------------
    def number_of_nodes(self):
        return int(sum(self._node_counts))

    num_nodes = number_of_nodes

    def number_of_edges(self):
        return int(sum(self._edge_counts))

    num_edges = number_of_edges

    def batch_num_nodes(self):
        return torch.as_tensor(self._node_counts, dtype=torch.int64)

    def batch_num_edges(self):
        return torch.as_tensor(self._edge_counts, dtype=torch.int64)

    @property
    def batch_size(self):
        return len(self._node_counts)

    def in_degrees(self):
        if self._csr is not None:
            indptr = self._csr[0]
            return (indptr[1:] - indptr[:-1]).long()
        d = np.bincount(self._global_edges_host()[1], minlength=self.number_of_nodes())
        return torch.as_tensor(d, dtype=torch.int64)

    def has_zero_in_degree(self):
        """Cached (one device->host read per graph): GATConv's DGLError condition."""
        if getattr(self, "_zero_in_deg", None) is None:
            n = self.number_of_nodes()
            self._zero_in_deg = bool(n > 0 and int(self.in_degrees().min()) == 0)
        return self._zero_in_deg

    def has_duplicate_edges(self):
        """True if some (src, dst) pair occurs more than once inside a graph (a multigraph).  Graphs that come from the
        reference's networkx ``Graph`` objects never do.  The deterministic arg-max backward (gts_segmax_bwd_det walks
        the out-edge lists) counts a duplicated edge once per copy, the atomic form once: it needs a simple graph."""
        if self._src is None:
            raise ops._lib.GtsError("has_duplicate_edges needs the host edge lists")
        s, d = self._global_edges_host()
        key = s * np.int64(max(self.number_of_nodes(), 1)) + d
        return bool(np.unique(key).size != key.size)

    def _global_edges_host(self):
        off = np.repeat(self._node_off.numpy()[:-1].astype(np.int64), self._edge_counts)
        return self._src.numpy().astype(np.int64) + off, self._dst.numpy().astype(np.int64) + off

    def edges(self):
        """Global (src, dst) in edge-id order (dgl.batch order), int64 CPU tensors."""
        s, d = self._global_edges_host()
        return torch.as_tensor(s), torch.as_tensor(d)

    # ---- device residency -------------------------------------------------
    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            if self.device.type == "cuda":
                raise ops._lib.GtsError("BatchedGraph: moving a device graph back to the host is not supported")
            return self
        if self.device == device and self._csr is not None:
            return self
        g = BatchedGraph.__new__(BatchedGraph)
        g.__dict__.update(self.__dict__)
        g.device = device
        g.ndata = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.ndata.items()}
        with torch.cuda.device(device):
            src = self._src.to(device, non_blocking=True)
            dst = self._dst.to(device, non_blocking=True)
            noff = self._node_off.to(device, non_blocking=True)
            eoff = self._edge_off.to(device, non_blocking=True)
            g._build_on_device(src, dst, noff, eoff)
        return g

    def _build_on_device(self, src, dst, noff, eoff):
        """Device side of dgl.batch + the lazy COO->CSR of DGL: global edge ids, then the in-edge CSR (K6)."""
        device = src.device
        if len(self._node_counts) > 1:
            gsrc, gdst = ops.batch_edges(src, dst, eoff, noff)
        else:
            gsrc, gdst = src, dst
        indptr, indices, eid = ops.csr_build(gdst, gsrc, self.number_of_nodes(), want_eid=True)
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=device)
        self._gsrc, self._gdst, self._eid = gsrc, gdst, eid
        self._csr = (indptr, indices)
        self._csc = None

    @classmethod
    def from_device_edges(cls, src_local, dst_local, node_off, edge_off, node_counts, edge_counts):
        """Batched graph whose per-graph LOCAL edge lists (int32) and offsets already live on the device (static
        input buffers of a CUDA-graphed step, trainer.GraphedStep): only the device build runs, nothing is copied."""
        g = cls.__new__(cls)
        g._src = g._dst = None                       # no host copy
        g._node_counts = [int(n) for n in node_counts]
        g._edge_counts = [int(e) for e in edge_counts]
        g._node_off, g._edge_off = node_off, edge_off
        g.device = src_local.device
        g.ndata = {}
        g._csr = g._csc = None
        g._gsrc = g._gdst = g._eid = None
        g.err_flag = None
        with torch.cuda.device(g.device):
            g._build_on_device(src_local, dst_local, node_off, edge_off)
        return g

    def cuda(self):
        return self.to("cuda")

    def device_tensors(self):
        """Every device tensor this graph owns (CSR, edge lists, lazily built CSC, error flag, ndata)."""
        out = [t for t in (getattr(self, "_gsrc", None), getattr(self, "_gdst", None), getattr(self, "_eid", None),
                           getattr(self, "err_flag", None)) if torch.is_tensor(t) and t.is_cuda]
        for grp in (self._csr, getattr(self, "_csc", None)):
            if grp is not None:
                out += [t for t in grp if torch.is_tensor(t) and t.is_cuda]
        out += [t for t in self.ndata.values() if torch.is_tensor(t) and t.is_cuda]
        return out

    def record_stream(self, stream):
        """Tell the caching allocator that ``stream`` uses this graph's tensors (they were staged on another stream:
        data_loader.DevicePrefetcher)."""
        for t in self.device_tensors():
            t.record_stream(stream)
        return self

    @property
    def csr(self):
        """(indptr int32[N+1], indices int32[E]) — in-edges by destination."""
        if self._csr is None:
            raise ops._lib.GtsError("BatchedGraph is on the host: call .to('cuda') first (no CPU compute path)")
        return self._csr

    @property
    def csr_eid(self):
        self.csr
        return self._eid

    @property
    def csc(self):
        """(indptr, indices = destinations, csc2csr) — out-edges by source; built on first use."""
        if self._csc is None:
            self.csr
            with torch.cuda.device(self.device):
                cptr, cidx, ceid = ops.csr_build(self._gsrc, self._gdst, self.number_of_nodes(), want_eid=True)
                c2r = ops.edge_perm_compose(self._eid, ceid)
            self._csc = (cptr, cidx, c2r)
        return self._csc


def from_edge_list(src, dst, n_nodes, pin=False):
    """Graph from a directed edge list (edge id = position)."""
    src = np.asarray(src)
    dst = np.asarray(dst)
    if src.shape != dst.shape:
        raise ValueError("src and dst must have the same length")
    if src.size and (src.min() < 0 or dst.min() < 0 or src.max() >= n_nodes or dst.max() >= n_nodes):
        raise ValueError("edge endpoint outside [0, n_nodes)")
    return BatchedGraph(src, dst, [n_nodes], [src.shape[0]], pin=pin)


def from_networkx(nx_graph):
    """dgl.from_networkx (data_processing/data_loader.py:72): undirected input
    becomes both directions, edge ids follow ``to_directed().edges()`` order,
    self-loops once, nodes relabelled to sorted consecutive integers."""
    g = nx_graph.to_directed() if not nx_graph.is_directed() else nx_graph
    nodes = sorted(g.nodes())
    n = len(nodes)
    identity = nodes == list(range(n))
    e = np.asarray(list(g.edges()), dtype=np.int64).reshape(-1, 2)
    if not identity and e.size:
        remap = {v: i for i, v in enumerate(nodes)}
        e = np.asarray([(remap[u], remap[v]) for u, v in e.tolist()], dtype=np.int64).reshape(-1, 2)
    return from_edge_list(e[:, 0], e[:, 1], n)


def batch(graphs, pin=False):
    """dgl.batch (data_loader.py:168): block-diagonal union in sample order."""
    graphs = list(graphs)
    if not graphs:
        raise ValueError("batch() needs at least one graph")
    for g in graphs:
        if g.device.type != "cpu":
            raise ValueError("batch() takes host graphs; batch first, then .to('cuda')")
    src = np.concatenate([g._src.numpy() for g in graphs])
    dst = np.concatenate([g._dst.numpy() for g in graphs])
    nc = [n for g in graphs for n in g._node_counts]
    ec = [e for g in graphs for e in g._edge_counts]
    return BatchedGraph(src, dst, nc, ec, pin=pin)


def minibatch_graphs(samples):
    """Collate function with the reference's signature
    (data_processing/data_loader.py:165-169): samples of
    (mri_id, graph, features, labels) -> (ids, batched graph, FloatTensor, LongTensor)."""
    mri_ids, graphs, features, labels = map(list, zip(*samples))
    batched_graph = batch(graphs)
    return (mri_ids, batched_graph, torch.FloatTensor(np.concatenate(features)),
            torch.LongTensor(np.concatenate(labels)))
