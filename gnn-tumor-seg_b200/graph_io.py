"""On-disk graph format -> host edge lists, without networkx or DGL.

The reference stores every MRI's region-adjacency graph as node-link JSON
(``save_networkx_graph`` / ``load_networkx_graph``, data_processing/graph_io.py:27-37)
and rebuilds a networkx graph and a DGLGraph from it for every sample of every
epoch (``ImageGraphDataset.get_graph``, data_processing/data_loader.py:67-83).
Here the JSON is parsed once into exactly the arrays the device CSR build
consumes, and cached next to it as a small binary file.

What ``dgl.from_networkx(load_networkx_graph(fp))`` sees (SURVEY.md Appendix A.4;
pinned by tests/golden/reference_kat*.npz, which hold the reference's own JSON
text and the edge order its loader produced):
* nodes in JSON order; ``features`` / ``label`` read per node in that order;
* the graph is undirected (``"directed": false``): every link (s, t) puts t in
  adj[s] and s in adj[t], in link order, duplicates ignored; the directed edge list is
  ``for u in nodes: for v in adj[u]`` — a self-loop appears once;
* node ids are relabelled to 0..N-1 in sorted order (identity for the files the
  preprocessing writes).
Both the ``"links"`` key (networkx 2.x, the reference's era) and ``"edges"``
(networkx >= 3.4 default) are accepted.
"""
from __future__ import annotations

import json
import os

import numpy as np

from .graph import BatchedGraph, from_edge_list
from .project import project_nodes_to_img  # noqa: F401  (data_processing/graph_io.py:21-24 lives in this module in the reference)

CACHE_SUFFIX = ".gtscache.npz"
CACHE_VERSION = 1


def parse_node_link_json(text):
    """node-link JSON text -> (src int32[E], dst int32[E], n_nodes, features float64[N,F] | None, labels int64[N] | None)."""
    d = json.loads(text)
    if d.get("multigraph", False):
        raise ValueError("multigraph node-link files are not produced by the preprocessing and are not supported")
    nodes = d["nodes"]
    links = d["links"] if "links" in d else d["edges"]
    ids = [n["id"] for n in nodes]
    order = {v: i for i, v in enumerate(ids)}              # insertion order = JSON order
    src_key = "source"
    dst_key = "target"
    directed = bool(d.get("directed", False))
    # adjacency in insertion order (dict keeps first insertion, like networkx's adjacency dicts)
    adj = [dict() for _ in ids]
    for l in links:
        s, t = l[src_key], l[dst_key]
        for v in (s, t):
            if v not in order:                              # node_link_graph adds endpoints it has not seen
                order[v] = len(ids)
                ids.append(v)
                adj.append(dict())
        si, ti = order[s], order[t]
        adj[si].setdefault(ti, None)
        if not directed:
            adj[ti].setdefault(si, None)
    n = len(ids)
    counts = np.fromiter((len(a) for a in adj), dtype=np.int64, count=n)
    src = np.repeat(np.arange(n, dtype=np.int64), counts)
    dst = np.fromiter((v for a in adj for v in a), dtype=np.int64, count=int(counts.sum()))
    # dgl.from_networkx relabels to sorted consecutive integers
    try:
        sorted_ids = sorted(ids)
    except TypeError:
        sorted_ids = ids
    if sorted_ids != ids:
        rank = {v: i for i, v in enumerate(sorted_ids)}
        remap = np.fromiter((rank[v] for v in ids), dtype=np.int64, count=n)
        src, dst = remap[src], remap[dst]
    feats = labels = None
    if nodes and "features" in nodes[0]:
        feats = np.asarray([nd["features"] for nd in nodes], dtype=np.float64)
    if nodes and "label" in nodes[0]:
        labels = np.asarray([nd["label"] for nd in nodes], dtype=np.int64)
    return src.astype(np.int32), dst.astype(np.int32), n, feats, labels


def load_graph_json(fp, use_cache=True):
    """``{mri_id}_nxgraph.json`` -> (host BatchedGraph, features, labels) — the triple
    ``ImageGraphDataset.get_graph`` returns (data_loader.py:67-83), features float64 [N,F]
    exactly as ``np.array([...])`` there.  The parsed arrays are cached beside the JSON
    (``<fp>.gtscache.npz``) and re-used while the JSON is not newer."""
    cache = fp + CACHE_SUFFIX
    if use_cache and os.path.exists(cache) and os.path.getmtime(cache) >= os.path.getmtime(fp):
        try:
            z = np.load(cache)
            if int(z["version"]) == CACHE_VERSION:
                feats = z["features"] if "features" in z.files else None
                labels = z["labels"] if "labels" in z.files else None
                return from_edge_list(z["src"], z["dst"], int(z["n_nodes"])), feats, labels
        except Exception:
            pass                                            # unreadable cache: fall through to the JSON
    with open(fp, "r") as f:
        src, dst, n, feats, labels = parse_node_link_json(f.read())
    if use_cache:
        arrays = {"version": CACHE_VERSION, "src": src, "dst": dst, "n_nodes": n}
        if feats is not None:
            arrays["features"] = feats
        if labels is not None:
            arrays["labels"] = labels
        try:
            tmp = cache + ".tmp.npz"
            np.savez(tmp, **arrays)
            os.replace(tmp, cache)
        except OSError:
            pass                                            # read-only dataset directory: no cache
    return from_edge_list(src, dst, n), feats, labels


def save_graph_json(src, dst, n_nodes, features, labels, fp, edges_key="links"):
    """Writer twin of ``save_networkx_graph`` for undirected graphs given as a directed
    (both directions present) edge list: emits each undirected link once, in first-seen order."""
    seen = set()
    links = []
    for s, t in zip(np.asarray(src).tolist(), np.asarray(dst).tolist()):
        key = (s, t) if s <= t else (t, s)
        if key not in seen:
            seen.add(key)
            links.append({"source": s, "target": t})
    nodes = []
    for i in range(int(n_nodes)):
        nd = {"id": i}
        if features is not None:
            nd["features"] = [float(x) for x in np.asarray(features[i]).tolist()]
        if labels is not None:
            nd["label"] = int(labels[i])
        nodes.append(nd)
    with open(fp, "w") as f:
        f.write(json.dumps({"directed": False, "multigraph": False, "graph": {}, "nodes": nodes, edges_key: links}))


__all__ = ["parse_node_link_json", "load_graph_json", "save_graph_json", "project_nodes_to_img", "BatchedGraph"]
