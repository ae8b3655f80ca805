"""Pin the oracle against fixtures produced by RUNNING the reference's own
functions (tests/golden/make_golden.py -> reference_kat.npz).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import graph_ref, project_ref
from gnn_tumor_seg_b200 import synth

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_kat.npz"))


def test_project_docstring_partition():
    out = project_ref.project_nodes_to_img_ref(GOLD["part"], np.array([3, 1, 2]))
    assert np.array_equal(out, GOLD["proj"])
    assert np.array_equal(out, [[0, 0, 0, 0], [3, 3, 1, 0], [3, 1, 1, 0], [2, 2, 1, 0]])
    assert out.dtype == np.dtype(str(GOLD["proj_dtype"]))


def test_region_adjacency_matches_find_adjacent_nodes():
    r, c = synth.region_adjacency_edges(GOLD["part"], 3)
    assert np.array_equal(r, GOLD["rows"]) and np.array_equal(c, GOLD["cols"])
    assert GOLD["adj"].all()                                   # all-ones 3x3 incl. self loops
    n3 = int(GOLD["n3"])
    r3, c3 = synth.region_adjacency_edges(GOLD["part3"], n3)
    assert np.array_equal(r3, GOLD["r3"]) and np.array_equal(c3, GOLD["c3"])


def test_networkx_edge_order_and_node_data():
    import networkx as nx
    n3 = int(GOLD["n3"])
    G = nx.Graph()
    G.add_nodes_from(range(n3))
    G.add_edges_from(zip(GOLD["r3"].tolist(), GOLD["c3"].tolist()))
    s, d, n = graph_ref.edges_from_networkx(G)
    assert n == n3
    assert np.array_equal(np.stack([s, d], 1), GOLD["nx_edges"])
    # lexicographic (src,dst), self-loop once == np.where order of the adjacency matrix
    assert np.array_equal(s, GOLD["r3"]) and np.array_equal(d, GOLD["c3"])


def test_project3d_and_chain():
    assert np.array_equal(project_ref.project_nodes_to_img_ref(GOLD["part3"], GOLD["labels3"]), GOLD["proj3"])
    crop = (GOLD["crop_ix0"], GOLD["crop_ix1"], GOLD["crop_ix2"])
    vox = project_ref.project_nodes_to_img_ref(GOLD["svs_c"], GOLD["pred_nodes"])
    assert np.array_equal(vox, GOLD["vox"])
    full = project_ref.save_voxel_preds_ref(GOLD["logits"], GOLD["svs_c"], crop)
    assert full.shape == (240, 240, 155) and full.dtype == np.int16 == np.dtype(str(GOLD["vox_brats_dtype"]))
    assert hashlib.sha256(full.tobytes()).digest() == GOLD["vox_brats_sha256"].tobytes()
    nz = np.flatnonzero(full)
    assert np.array_equal(nz, GOLD["vox_brats_nonzero_idx"])
    assert np.array_equal(full.reshape(-1)[nz], GOLD["vox_brats_nonzero_val"])
    assert np.array_equal(np.argmax(GOLD["logits"], 1), GOLD["pred_nodes"])   # first maximum on ties


def test_logit_projection_and_relabel():
    out = project_ref.save_voxel_logits_ref(GOLD["logits"], GOLD["svs_c"])
    assert np.array_equal(out, GOLD["vox_logits"])
    assert np.array_equal(project_ref.swap_labels_to_brats_ref(np.array([0, 1, 2, 3, 3, 0])), GOLD["swapped"])
    assert np.array_equal(GOLD["swapped"], [0, 2, 1, 4, 4, 0])
    with pytest.raises(RuntimeError, match=str(GOLD["swap_error"])):
        project_ref.swap_labels_to_brats_ref(np.array([0, 5]))


def test_synthetic_rag_statistics():
    """seed 0 -> N=15000, E=223512, in-degree 4/14.90/15/30."""
    g = synth.make_graph(0)
    assert g.n_nodes == 15000 and g.n_edges == 223512
    deg = np.bincount(g.dst, minlength=g.n_nodes)
    assert deg.min() == 4 and deg.max() == 30 and abs(deg.mean() - 14.9008) < 1e-3
    key = g.src.astype(np.int64) * g.n_nodes + g.dst
    assert np.all(np.diff(key) > 0)                            # lexicographic, no duplicates
    assert np.array_equal(np.sort(key), np.sort(g.dst.astype(np.int64) * g.n_nodes + g.src))   # symmetric
