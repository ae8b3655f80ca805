"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN FUNCTIONS.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The fixtures are committed; nothing at test/bench time reads /root/reference.

Importable reference pieces used (DGL-dependent modules cannot be imported —
``import dgl`` fails in this image):
  data_processing.graph_io.project_nodes_to_img / save_networkx_graph / load_networkx_graph
  data_processing.image_processing.uncrop_to_brats_size / determine_brain_crop
  scripts.preprocess_dataset.swap_labels_to_brats      (nibabel stubbed: only nifti I/O needs it)
  mri2graph.graphgen.find_adjacent_nodes               (skimage stubbed: only slic needs it)
The save_voxel_preds / save_voxel_logits chains (scripts/generate_gnn_predictions.py:55-73)
are reproduced by calling those reference functions in the script's order,
because the script itself imports DGL through model.gnn_model.
"""
import hashlib
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def main():
    sys.path.insert(0, REF)
    sk = _stub("skimage")
    sk.segmentation = _stub("skimage.segmentation", slic=None)
    _stub("nibabel")
    import networkx as nx
    from data_processing import graph_io, image_processing
    from mri2graph import graphgen
    from scripts.preprocess_dataset import swap_labels_to_brats

    rng = np.random.default_rng(20211018)

    # 1. docstring partition (mri2graph/graphgen.py:227-230) through project_nodes_to_img
    part = np.array([[-1, -1, -1, -1], [0, 0, 1, -1], [0, 1, 1, -1], [2, 2, 1, -1]], dtype=np.int16)
    proj = graph_io.project_nodes_to_img(part, np.array([3, 1, 2]))
    adj = graphgen.find_adjacent_nodes(part.copy(), 3, as_mat=True)
    rows, cols = graphgen.find_adjacent_nodes(part.copy(), 3)

    # 2. a seeded 3-D partition: 40 regions in 14x12x10, background shell
    shape = (14, 12, 10)
    seeds = rng.uniform(0, 1, size=(40, 3)) * np.array(shape)
    gx, gy, gz = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    pts = np.stack([gx, gy, gz], -1).reshape(-1, 3).astype(np.float64)
    owner = np.argmin(((pts[:, None, :] - seeds[None]) ** 2).sum(-1), axis=1)
    _, owner = np.unique(owner, return_inverse=True)
    part3 = owner.reshape(shape).astype(np.int16)
    part3[0], part3[-1], part3[:, 0], part3[:, :, -1] = -1, -1, -1, -1
    _, comp = np.unique(part3[part3 >= 0], return_inverse=True)
    part3[part3 >= 0] = comp.astype(np.int16)
    n3 = int(part3.max()) + 1
    r3, c3 = graphgen.find_adjacent_nodes(part3.copy(), n3)
    labels3 = rng.integers(0, 4, size=n3)
    proj3 = graph_io.project_nodes_to_img(part3, labels3)

    # 3. node-link JSON round trip -> directed edge order seen by from_networkx
    G = nx.Graph()
    G.add_nodes_from(range(n3))
    for u, v in zip(r3.tolist(), c3.tolist()):
        G.add_edge(u, v)
    for n in G.nodes:
        G.nodes[n]["features"] = [float(n)] * 2
        G.nodes[n]["label"] = int(labels3[n])
    with tempfile.TemporaryDirectory() as td:
        fp = os.path.join(td, "g_nxgraph.json")
        graph_io.save_networkx_graph(G, fp)
        G2 = graph_io.load_networkx_graph(fp)
    e = np.asarray(list(G2.to_directed().edges()), dtype=np.int64)
    feats_rt = np.array([G2.nodes[n]["features"] for n in G2.nodes])
    labels_rt = np.array([G2.nodes[n]["label"] for n in G2.nodes])

    # 4. crop / uncrop / relabel chain on a brain box inside the BraTS volume
    box = np.zeros((240, 240, 155), np.float32)
    box[60:60 + shape[0], 100:100 + shape[1], 30:30 + shape[2]] = 1.0
    box[60 + 3, :, :] = 0.0      # a black plane: crop indices are NOT contiguous
    ix = image_processing.determine_brain_crop(box)
    crop_shape = tuple(int(a.size) for a in ix)
    svs_c = np.ascontiguousarray(np.delete(part3, 3, axis=0))
    assert svs_c.shape == crop_shape, (svs_c.shape, crop_shape)
    logits = rng.normal(size=(n3, 4)).astype(np.float32)
    logits[5] = [0.5, 0.5, 0.5, 0.5]          # tie -> first maximum
    logits[6] = [-1.0, 2.0, 2.0, 0.0]
    pred_nodes = np.argmax(logits, axis=1)     # torch.max(logits,1) picks the first maximum too
    vox = graph_io.project_nodes_to_img(svs_c, pred_nodes)
    vox_full = image_processing.uncrop_to_brats_size(ix, vox)
    vox_brats = swap_labels_to_brats(vox_full)
    DEFAULT_BACKGROUND_NODE_LOGITS = [[1.0, -1.0, -1.0, -1.0]]
    vox_logits = np.concatenate([logits, DEFAULT_BACKGROUND_NODE_LOGITS])[svs_c]
    swapped = swap_labels_to_brats(np.array([0, 1, 2, 3, 3, 0], dtype=np.int64))
    try:
        swap_labels_to_brats(np.array([0, 5]))
        raised = ""
    except RuntimeError as ex:
        raised = str(ex)

    np.savez_compressed(
        os.path.join(OUT, "reference_kat.npz"),
        part=part, proj=proj, adj=adj, rows=rows, cols=cols,
        part3=part3, n3=n3, r3=r3, c3=c3, labels3=labels3, proj3=proj3,
        nx_edges=e, feats_rt=feats_rt, labels_rt=labels_rt,
        crop_ix0=ix[0], crop_ix1=ix[1], crop_ix2=ix[2], svs_c=svs_c, logits=logits,
        pred_nodes=pred_nodes, vox=vox, vox_brats_nonzero_idx=np.flatnonzero(vox_brats).astype(np.int64),
        vox_brats_nonzero_val=vox_brats.reshape(-1)[np.flatnonzero(vox_brats)],
        vox_brats_sha256=np.frombuffer(hashlib.sha256(vox_brats.tobytes()).digest(), dtype=np.uint8),
        vox_brats_dtype=str(vox_brats.dtype), vox_dtype=str(vox.dtype), proj_dtype=str(proj.dtype),
        vox_logits=vox_logits, swapped=swapped, swap_error=raised,
    )
    # 5. determine_tumor_crop (data_processing/image_processing.py:8-17) on the projected predictions, on an
    #    all-healthy volume, and on tumour voxels touching the volume border; and the node-link JSON text the
    #    reference's save_networkx_graph writes (the on-disk graph format, data_processing/graph_io.py:27-37)
    crops = {}
    cases = {"pred": vox, "healthy": np.zeros_like(vox)}
    edge = np.zeros((9, 7, 5), dtype=np.int64)
    edge[0, 3, 2] = 2
    edge[8, 6, 4] = 1
    edge[4, 0, 0] = 3
    cases["border"] = edge
    sparse = np.zeros((12, 10, 8), dtype=np.int64)
    sparse[5, 4, 3] = 1
    cases["single"] = sparse
    for name, vol in cases.items():
        cx = image_processing.determine_tumor_crop(vol)
        for a in range(3):
            crops[f"tcrop_{name}_{a}"] = np.asarray(cx[a]).reshape(-1)
        crops[f"tcrop_{name}_vol"] = vol
    with tempfile.TemporaryDirectory() as td:
        fp = os.path.join(td, "g_nxgraph.json")
        graph_io.save_networkx_graph(G, fp)
        json_text = open(fp).read()
    np.savez_compressed(os.path.join(OUT, "reference_kat_crop.npz"), json_text=np.frombuffer(json_text.encode(), dtype=np.uint8),
                        **crops)
    print("wrote", os.path.join(OUT, "reference_kat_crop.npz"), {k: v.shape for k, v in crops.items() if k.endswith("_0")})

    print("wrote", os.path.join(OUT, "reference_kat.npz"), "n3 =", n3, "edges =", len(r3),
          "nonzero voxels =", int((vox_brats != 0).sum()), "dtypes", vox.dtype, vox_brats.dtype, vox_logits.dtype)


if __name__ == "__main__":
    main()
