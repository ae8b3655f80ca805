"""Host logic of the on-disk graph format: the node-link JSON parser against the JSON text and
the edge order produced by the REFERENCE's own save_networkx_graph / load_networkx_graph
(tests/golden/make_golden.py), the binary cache, and the dataset twin.  CPU only (no compute calls)."""
import json
import os

import numpy as np
import pytest

from gnn_tumor_seg_b200 import data_loader, graph, graph_io

HERE = os.path.dirname(__file__)
GOLD = np.load(os.path.join(HERE, "golden", "reference_kat.npz"))
GOLD_CROP = np.load(os.path.join(HERE, "golden", "reference_kat_crop.npz"))
JSON_TEXT = GOLD_CROP["json_text"].tobytes().decode()


def test_parser_matches_reference_loader_edge_order():
    src, dst, n, feats, labels = graph_io.parse_node_link_json(JSON_TEXT)
    assert n == int(GOLD["n3"])
    assert np.array_equal(np.stack([src, dst], 1), GOLD["nx_edges"])     # what to_directed().edges() gave the reference
    assert np.array_equal(feats, GOLD["feats_rt"]) and feats.dtype == np.float64
    assert np.array_equal(labels, GOLD["labels_rt"])


def test_links_and_edges_keys_and_networkx_agreement():
    import networkx as nx
    d = json.loads(JSON_TEXT)
    key = "links" if "links" in d else "edges"
    other = "edges" if key == "links" else "links"
    d2 = dict(d)
    d2[other] = d2.pop(key)
    a = graph_io.parse_node_link_json(json.dumps(d2))
    b = graph_io.parse_node_link_json(JSON_TEXT)
    assert all(np.array_equal(x, y) for x, y in zip(a[:2], b[:2]))
    # an irregular graph: shuffled link order, self loops, a node that only appears in a link
    rng = np.random.default_rng(3)
    G = nx.Graph()
    G.add_nodes_from(range(12))
    pairs = [(int(u), int(v)) for u, v in rng.integers(0, 12, size=(40, 2))]
    G.add_edges_from(pairs)
    data = nx.node_link_data(G, edges="links")
    s, t, n, _, _ = graph_io.parse_node_link_json(json.dumps(data))
    G2 = nx.node_link_graph(data, edges="links")
    ref = graph.from_networkx(G2)
    assert n == ref.number_of_nodes()
    assert np.array_equal(s, ref._src.numpy()) and np.array_equal(t, ref._dst.numpy())


def test_load_graph_json_cache_and_dataset(tmp_path):
    root = tmp_path / "ds"
    ids = ["BraTS_001", "BraTS_002"]
    for k, mid in enumerate(ids):
        d = root / mid
        d.mkdir(parents=True)
        (d / f"{mid}_nxgraph.json").write_text(JSON_TEXT)
        np.save(d / f"{mid}_crop.npy", np.array([np.arange(3), np.arange(4), np.arange(5)], dtype=object), allow_pickle=True)
    fp = str(root / ids[0] / f"{ids[0]}_nxgraph.json")
    g1, f1, l1 = graph_io.load_graph_json(fp)
    assert os.path.exists(fp + graph_io.CACHE_SUFFIX)
    g2, f2, l2 = graph_io.load_graph_json(fp)                     # served from the binary cache
    assert np.array_equal(g1._src.numpy(), g2._src.numpy()) and np.array_equal(g1._dst.numpy(), g2._dst.numpy())
    assert np.array_equal(f1, f2) and np.array_equal(l1, l2)
    assert g1.number_of_edges() == len(GOLD["nx_edges"])

    ds = data_loader.ImageGraphDataset(str(root) + os.sep, "BraTS", read_image=False, read_graph=True, read_label=True)
    assert sorted(ds.all_ids) == ids and len(ds) == 2
    mid, g, feats, labels = ds[0]
    assert isinstance(g, graph.BatchedGraph) and feats.shape == (int(GOLD["n3"]), 2) and labels.shape == (int(GOLD["n3"]),)
    # the collate function of the reference (data_loader.py:165-169): ids, batched graph, FloatTensor, LongTensor
    mris, bg, bf, bl = data_loader.minibatch_graphs([ds[0], ds[1]])
    assert bg.number_of_nodes() == 2 * int(GOLD["n3"]) and bf.dtype.is_floating_point and bl.dtype.__str__() == "torch.int64"
    assert len(ds.get_crop(ids[0])) == 3
    with pytest.raises(Exception, match="NIfTI"):
        ds.get_supervoxel_partitioning(ids[0])
    ds_nolabel = data_loader.ImageGraphDataset(str(root) + os.sep, "BraTS", read_image=False, read_label=False)
    assert len(ds_nolabel[1]) == 3                                # (id, graph, features)


def test_writer_round_trip(tmp_path):
    src, dst, n, feats, labels = graph_io.parse_node_link_json(JSON_TEXT)
    fp = str(tmp_path / "x_nxgraph.json")
    graph_io.save_graph_json(src, dst, n, feats, labels, fp)
    g, f, l = graph_io.load_graph_json(fp, use_cache=False)
    assert np.array_equal(g._src.numpy(), src) and np.array_equal(g._dst.numpy(), dst)
    assert np.array_equal(f, feats) and np.array_equal(l, labels)


def test_oracle_tumor_crop_pinned():
    from oracle import project_ref
    for case in ("pred", "healthy", "border", "single"):
        ix = project_ref.determine_tumor_crop_ref(GOLD_CROP[f"tcrop_{case}_vol"])
        for a in range(3):
            assert np.array_equal(np.asarray(ix[a]).reshape(-1), GOLD_CROP[f"tcrop_{case}_{a}"])
