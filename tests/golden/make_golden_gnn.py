"""Generate tests/golden/reference_gnn.npz by RUNNING THE REFERENCE'S OWN model/gnn_model.py (class GNN).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_gnn.py
The fixture is committed; nothing at test/bench time reads /root/reference.

What runs unmodified from /root/reference: ``model.gnn_model.GNN`` (__init__, run_epoch, evaluate,
calculate_all_metrics_for_brain, save_weights), ``model.networks.init_graph_net / GraphSage``, ``model.evaluation``,
``data_processing.data_loader.ImageGraphDataset / minibatch_graphs``, ``data_processing.graph_io`` (the dataset's graphs
are written with its save_networkx_graph), ``utils.hyperparam_helpers.FullParamSet``.  What is stubbed, because DGL and
nibabel cannot be installed here: ``dgl.from_networkx`` / ``dgl.batch`` (oracle/graph_ref.py: edge order pinned to the
reference's own edge list in tests/test_oracle_golden.py), ``dgl.nn.pytorch`` conv classes (the oracle's per-layer
modules: "parity unpinned" arithmetic), ``nibabel.load`` (reads the ``.npy`` twin of a ``.nii.gz`` path), and
``ExponentialLR`` is wrapped to accept the ``verbose`` keyword current torch no longer takes.

What the fixture pins to the reference's code: the training loop (loss = CrossEntropyLoss(weight), zero_grad / backward
/ AdamW(lr, weight_decay = hp.w_decay) per batch, ExponentialLR once per epoch, the epoch's return value = mean of the
batch losses), ``evaluate`` (per-brain loss + the ten metric slots, means over brains, summed label counts) and the
checkpoint file ``save_weights`` writes.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(OUT))
sys.path.insert(0, ROOT)
sys.path.insert(0, OUT)

from make_golden_stack import install_dgl_stub          # noqa: E402  (conv classes over the oracle's layers)
from oracle import graph_ref                            # noqa: E402


class StubGraph:
    """Stands in for a DGLGraph: the calls the reference makes on one (data_loader.py:72-79, gnn_model.py:38,58)."""

    def __init__(self, src, dst, n):
        self.src, self.dst, self.n = np.asarray(src, np.int64), np.asarray(dst, np.int64), int(n)
        self.ndata = {}
        self.indptr, self.indices, _ = graph_ref.csr_by_dst_ref(self.src, self.dst, self.n)

    def number_of_edges(self):
        return int(self.src.size)

    def number_of_nodes(self):
        return self.n

    def in_degrees(self):
        return torch.as_tensor(np.bincount(self.dst, minlength=self.n))

    def to(self, device):
        return self

    def __iter__(self):                                   # the stub SAGEConv unpacks (indptr, indices)
        return iter((self.indptr, self.indices))


def install_io_stubs():
    import dgl                                            # the stub package make_golden_stack installed
    dgl.from_networkx = lambda g: StubGraph(*graph_ref.edges_from_networkx(g))
    dgl.batch = lambda graphs: StubGraph(*graph_ref.batch_graphs_ref([(g.src, g.dst, g.n) for g in graphs])[:3])

    class _Img:
        def __init__(self, arr):
            self.dataobj = arr
    # torch >= 2.7 dropped the ``verbose`` keyword the reference passes (gnn_model.py:29): accept and ignore it
    base = torch.optim.lr_scheduler.ExponentialLR

    class ExponentialLR(base):
        def __init__(self, optimizer, gamma, last_epoch=-1, verbose=False):
            super().__init__(optimizer, gamma, last_epoch)
    torch.optim.lr_scheduler.ExponentialLR = ExponentialLR
    nib = types.ModuleType("nibabel")
    nib.load = lambda fp: _Img(np.load(fp.replace(".nii.gz", ".npy")))
    sys.modules["nibabel"] = nib


def make_dataset(root, n_graphs, ref_graph_io):
    """n_graphs small synthetic brains in the reference's on-disk layout.  Returns the arrays the test needs to
    rebuild the same dataset without the reference."""
    import networkx as nx
    from gnn_tumor_seg_b200 import synth
    from oracle import project_ref
    rng = np.random.default_rng(11)
    rec = {}
    for s in range(n_graphs):
        g = synth.make_small_graph(80 + s, n_nodes=120 + 15 * s, avg_deg=7)
        mri = f"BraTS_{s:03d}"
        os.makedirs(os.path.join(root, mri))
        G = nx.Graph()
        G.add_nodes_from(range(g.n_nodes))
        G.add_edges_from(zip(g.src.tolist(), g.dst.tolist()))        # symmetric list incl. self-loops -> undirected graph
        for n in G.nodes:
            G.nodes[n]["features"] = [float(x) for x in g.features[n]]
            G.nodes[n]["label"] = int(g.labels[n])
        ref_graph_io.save_networkx_graph(G, os.path.join(root, mri, f"{mri}_nxgraph.json"))
        svs = rng.integers(-1, g.n_nodes, size=(12, 11, 10)).astype(np.int16)
        truth = project_ref.project_nodes_to_img_ref(svs, g.labels).astype(np.int16)
        flip = rng.random(truth.shape) < 0.03
        truth[flip] = rng.integers(0, 4, size=int(flip.sum()))
        np.save(os.path.join(root, mri, f"{mri}_supervoxels.npy"), svs)
        np.save(os.path.join(root, mri, f"{mri}_label.npy"), truth)
        rec[f"{mri}/json"] = np.frombuffer(open(os.path.join(root, mri, f"{mri}_nxgraph.json"), "rb").read(), dtype=np.uint8)
        rec[f"{mri}/svs"], rec[f"{mri}/truth"] = svs, truth
    return rec


def main():
    record = []
    install_dgl_stub(record)
    install_io_stubs()
    sys.path.insert(0, REF)
    from data_processing import graph_io as ref_graph_io
    from data_processing.data_loader import ImageGraphDataset, minibatch_graphs
    from model import gnn_model as ref_gnn
    from utils.hyperparam_helpers import FullParamSet

    out = {}
    n_graphs, n_epochs = 8, 3
    hp = FullParamSet(n_epochs, 20, 4, 1e-3, 0.9, 1e-2, [0.1, 1.0, 2.0, 2.0], [32, 16], 0, None, None)
    with tempfile.TemporaryDirectory() as td:
        root = td + os.sep
        out.update(make_dataset(root, n_graphs, ref_graph_io))
        ds = ImageGraphDataset(root, "BraTS", read_image=False, read_graph=True, read_label=True)
        ds.all_ids = sorted(ds.all_ids)                   # glob order is file-system order: fix it
        torch.manual_seed(4)
        model = ref_gnn.GNN("GSpool", hp, ds)
        assert str(model.device) == "cpu"
        # the reference shuffles (DataLoader(shuffle=True), gnn_model.py:31); same loader without the shuffle, so that
        # the batch composition is part of the fixture rather than of torch's RNG stream
        model.train_loader = torch.utils.data.DataLoader(ds, batch_size=ref_gnn.BATCH_SIZE, shuffle=False, num_workers=0,
                                                         collate_fn=minibatch_graphs)
        for k, v in model.net.state_dict().items():
            out["init_sd/" + k] = v.clone().numpy()
        losses, lrs = [], []
        for _ in range(n_epochs):
            losses.append(model.run_epoch())
            lrs.append(model.optimizer.param_groups[0]["lr"])
        out["epoch_losses"], out["epoch_lrs"] = np.array(losses, np.float64), np.array(lrs, np.float64)
        for k, v in model.net.state_dict().items():
            out["final_sd/" + k] = v.clone().numpy()
        subset = torch.utils.data.Subset(ds, [0, 2, 3, 5, 7])
        avg, counts = model.evaluate(subset)
        out["eval_subset"], out["eval_avg"], out["eval_counts"] = np.array([0, 2, 3, 5, 7]), avg, counts
        model.save_weights(root, "ckpt")
        sd = torch.load(root + "ckpt.pt")
        out["ckpt_keys"] = np.array(sorted(sd.keys()))
        out["ckpt_equals_net"] = np.array(all(torch.equal(sd[k], v) for k, v in model.net.state_dict().items()))
    out["hp"] = np.array(repr(tuple(hp)))
    out["batch_size"] = np.int64(ref_gnn.BATCH_SIZE)
    out["ids"] = np.array(ds.all_ids)
    out["layers"] = np.array([repr(r) for r in record])
    np.savez_compressed(os.path.join(OUT, "reference_gnn.npz"), **out)
    print("epoch losses", losses, "lrs", lrs)
    print("evaluate avg", avg, "counts", counts)
    print("layers", record)


if __name__ == "__main__":
    main()
