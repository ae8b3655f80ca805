"""Deterministic synthetic supervoxel region-adjacency graphs.

There is no dataset on the GPU box, so tests, smoke() and bench.py all draw
their inputs from here.  The generator mimics what the reference's offline
preprocessing produces for ``-k 0`` (true region adjacency):

* partition: int16 supervoxel map over the BraTS volume (240,240,155), -1 =
  background, ids 0..N-1 (reference: mri2graph/graphgen.py:71-90,243 — SLIC
  followed by discard_empty_svs; here a Voronoi partition of an ellipsoidal
  "brain", SLIC needs skimage which is absent); ids are numbered in raster
  order of a coarse grid like SLIC's grid-initialised clusters;
* graph: 6-connectivity region adjacency with a self-loop on every node, as
  mri2graph/graphgen.py:161-196 (find_adjacent_nodes) builds it, returned as
  the directed edge list dgl.from_networkx would see (lexicographic (src,dst),
  self-loop once; data_processing/data_loader.py:72);
* features fp32 [N,20] ~ N(0,1), labels categorical p=(.90,.05,.03,.02).

This is input synthesis, not the hot path: plain numpy/scipy on the host.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

BRATS_SHAPE = (240, 240, 155)
ELLIPSOID_CENTRE = (120.0, 120.0, 77.5)
ELLIPSOID_SEMI_AXES = (70.0, 85.0, 70.0)
LABEL_P = (0.90, 0.05, 0.03, 0.02)


@dataclass
class SynthGraph:
    """One synthetic MRI: graph + node data (+ optional voxel partition)."""
    mri_id: str
    n_nodes: int
    src: np.ndarray          # int32 [E], lexicographic (src,dst)
    dst: np.ndarray          # int32 [E]
    features: np.ndarray     # float32 [N,20]
    labels: np.ndarray       # int64 [N]
    svs: np.ndarray | None = None      # int16 cropped partition [X,Y,Z], -1 background
    crop: tuple | None = None          # np.ix_-style index arrays into BRATS_SHAPE

    @property
    def n_edges(self) -> int:
        return int(self.src.shape[0])


def region_adjacency_edges(svs: np.ndarray, n_nodes: int):
    """Directed edge list of the 6-connectivity region-adjacency graph.

    Same result as ``np.where(find_adjacent_nodes(svs, n_nodes, as_mat=True))``
    (reference mri2graph/graphgen.py:161-196) without the dense
    (n+1)x(n+1) bool matrix: symmetric, self-loop on every node, background
    (-1) excluded, pairs sorted lexicographically by (row, col).
    """
    n1 = np.int64(n_nodes + 1)
    keys = []
    for axis in range(svs.ndim):
        lo = [slice(None)] * svs.ndim
        hi = [slice(None)] * svs.ndim
        lo[axis] = slice(None, -1)
        hi[axis] = slice(1, None)
        a = svs[tuple(lo)]
        b = svs[tuple(hi)]
        m = (a != b) & (a >= 0) & (b >= 0)
        a64 = a[m].astype(np.int64)
        b64 = b[m].astype(np.int64)
        keys.append(np.unique(a64 * n1 + b64))
        keys.append(np.unique(b64 * n1 + a64))
    diag = np.arange(n_nodes, dtype=np.int64)
    keys.append(diag * n1 + diag)
    k = np.unique(np.concatenate(keys))
    return (k // n1).astype(np.int32), (k % n1).astype(np.int32)


def voronoi_partition(seed: int, n_seeds: int = 15000, shape=BRATS_SHAPE,
                      centre=ELLIPSOID_CENTRE, semi_axes=ELLIPSOID_SEMI_AXES):
    """int16 supervoxel map over ``shape``: nearest of ``n_seeds`` uniform seeds
    inside the ellipsoid; outside = -1; empty cells dropped and ids compacted
    (the effect of discard_empty_svs, graphgen.py:71-90)."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    c = np.asarray(centre)
    r = np.asarray(semi_axes)
    gx, gy, gz = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    inside = (((gx - c[0]) / r[0]) ** 2 + ((gy - c[1]) / r[1]) ** 2 + ((gz - c[2]) / r[2]) ** 2) <= 1.0
    pts = np.stack([gx[inside], gy[inside], gz[inside]], axis=1)
    # uniform seeds in the ellipsoid by rejection from the bounding box
    seeds = np.empty((0, 3))
    while seeds.shape[0] < n_seeds:
        cand = rng.uniform(-1.0, 1.0, size=(2 * n_seeds, 3))
        cand = cand[(cand ** 2).sum(1) <= 1.0]
        seeds = np.concatenate([seeds, cand * r + c])
    seeds = seeds[:n_seeds]
    # SLIC (mri2graph/graphgen.py:243) initialises its clusters on a regular grid and numbers
    # them in C (raster) order, so real supervoxel ids are spatially coherent: consecutive ids
    # are neighbours in space.  Give the synthetic ids the same property: order the seeds by
    # the raster index of a coarse grid cell (x, then y, then z fastest).  This only relabels
    # nodes — edge count and degree statistics are unchanged.
    cell = 6.0
    key = (np.floor(seeds[:, 0] / cell) * 4096 + np.floor(seeds[:, 1] / cell)) * 4096 + np.floor(seeds[:, 2] / cell)
    seeds = seeds[np.argsort(key, kind="stable")]
    _, owner = cKDTree(seeds).query(pts, workers=-1)
    uniq, compact = np.unique(owner, return_inverse=True)
    vol = np.full(shape, -1, dtype=np.int16)
    vol[inside] = compact.astype(np.int16)
    return vol, int(uniq.shape[0])


def brain_crop(vol: np.ndarray):
    """np.ix_ crop of all planes that hold at least one supervoxel (what
    determine_brain_crop, data_processing/image_processing.py:31-41, yields
    for a non-empty brain mask)."""
    mask = vol >= 0
    return np.ix_(mask.any(axis=(1, 2)), mask.any(axis=(0, 2)), mask.any(axis=(0, 1)))


def node_data(g: int, n_nodes: int, in_feats: int = 20):
    import torch
    gen = torch.Generator().manual_seed(1000 + g)
    feats = torch.randn(n_nodes, in_feats, generator=gen, dtype=torch.float32).numpy()
    rng = np.random.default_rng(2000 + g)
    labels = rng.choice(len(LABEL_P), size=n_nodes, p=LABEL_P).astype(np.int64)
    return feats, labels


def make_graph(g: int, n_seeds: int = 15000, with_partition: bool = False,
               cache_dir: str | None = None) -> SynthGraph:
    """Synthetic graph ``g`` (partition seed = g, node-data seed = 1000+g)."""
    cache = None
    if cache_dir is None:
        cache_dir = os.environ.get("GTS_SYNTH_CACHE") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), ".synth_cache")
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        cache = os.path.join(cache_dir, f"rag_raster_s{g}_n{n_seeds}.npz")
    if cache and os.path.exists(cache):
        z = np.load(cache)
        n_nodes, src, dst = int(z["n_nodes"]), z["src"], z["dst"]
        svs_c = z["svs"] if with_partition else None
        crop = tuple(z[f"crop{i}"] for i in range(3)) if with_partition else None
    else:
        vol, n_nodes = voronoi_partition(g, n_seeds)
        src, dst = region_adjacency_edges(vol, n_nodes)
        crop = brain_crop(vol)
        svs_c = np.ascontiguousarray(vol[crop])
        if cache:
            tmp = cache + f".{os.getpid()}.tmp.npz"
            np.savez_compressed(tmp, n_nodes=n_nodes, src=src, dst=dst, svs=svs_c,
                                crop0=crop[0], crop1=crop[1], crop2=crop[2])
            os.replace(tmp, cache)
        if not with_partition:
            svs_c, crop = None, None
    feats, labels = node_data(g, n_nodes)
    return SynthGraph(f"synth_{g:05d}", n_nodes, src, dst, feats, labels, svs_c, crop)


def make_small_graph(g: int, n_nodes: int = 200, avg_deg: int = 6, in_feats: int = 20,
                     self_loops: bool = True, isolated: int = 0) -> SynthGraph:
    """Small random symmetric graph for unit tests (no voxel partition).
    ``isolated`` trailing nodes get no edges at all (zero in-degree case)."""
    rng = np.random.default_rng(g)
    live = n_nodes - isolated
    m = max(1, live * avg_deg // 2)
    a = rng.integers(0, live, size=m)
    b = rng.integers(0, live, size=m)
    keep = a != b
    a, b = a[keep], b[keep]
    n1 = np.int64(n_nodes + 1)
    keys = [a.astype(np.int64) * n1 + b, b.astype(np.int64) * n1 + a]
    if self_loops:
        d = np.arange(live, dtype=np.int64)
        keys.append(d * n1 + d)
    k = np.unique(np.concatenate(keys))
    src, dst = (k // n1).astype(np.int32), (k % n1).astype(np.int32)
    import torch
    gen = torch.Generator().manual_seed(1000 + g)
    feats = torch.randn(n_nodes, in_feats, generator=gen, dtype=torch.float32).numpy()
    labels = rng.choice(len(LABEL_P), size=n_nodes, p=LABEL_P).astype(np.int64)
    return SynthGraph(f"small_{g:05d}", n_nodes, src, dst, feats, labels)
