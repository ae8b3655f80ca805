"""GNN.evaluate / save_weights / reference-layout checkpoints through the reference's entry points
(model/gnn_model.py:51-90) on an on-disk dataset in the reference's layout
({root}/{id}/{id}_nxgraph.json + voxel volumes), against the CPU oracle + the evaluation functions pinned to the
reference's model/evaluation.py."""
import os
from collections import namedtuple

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gnn_tumor_seg_b200 import evaluation, graph_io, networks, synth
from gnn_tumor_seg_b200.data_loader import ImageGraphDataset
from oracle import graph_ref, project_ref, sage_ref

pytestmark = pytest.mark.gpu
HP = namedtuple("HP", "in_feats out_classes layer_sizes gat_heads gat_residuals class_weights lr w_decay lr_decay")


def _write_dataset(root, n_graphs=3):
    rng = np.random.default_rng(3)
    ids = []
    for s in range(n_graphs):
        g = synth.make_small_graph(40 + s, n_nodes=150 + 20 * s, avg_deg=7)
        mri = f"BraTS_{s:03d}"
        os.makedirs(os.path.join(root, mri))
        graph_io.save_graph_json(g.src, g.dst, g.n_nodes, g.features, g.labels, os.path.join(root, mri, f"{mri}_nxgraph.json"))
        svs = rng.integers(-1, g.n_nodes, size=(14, 13, 12)).astype(np.int16)
        np.save(os.path.join(root, mri, f"{mri}_supervoxels.npy"), svs)
        # voxel ground truth = projected node labels with a few voxels flipped (so Dice < 1 and HD95 > 0)
        truth = project_ref.project_nodes_to_img_ref(svs, g.labels).astype(np.int16)
        flip = rng.random(truth.shape) < 0.02
        truth[flip] = rng.integers(0, 4, size=int(flip.sum()))
        np.save(os.path.join(root, mri, f"{mri}_label.npy"), truth)
        ids.append(mri)
    return ids


def test_evaluate_and_save_weights_round_trip(cuda_dev, tmp_path):
    from gnn_tumor_seg_b200.gnn_model import GNN
    root = str(tmp_path / "data") + os.sep
    os.makedirs(root)
    _write_dataset(root)
    ds = ImageGraphDataset(root, "BraTS", read_image=False, read_graph=True, read_label=True)
    assert len(ds) == 3
    subset = torch.utils.data.Subset(ds, list(range(len(ds))))           # the reference evaluates Subsets (gnn_model.py:52)
    hp = HP(20, 4, [64, 64], None, None, [0.1, 1.0, 2.0, 2.0], 1e-3, 1e-4, 0.98)
    torch.manual_seed(0)
    model = GNN("GSpool", hp, ds)
    model.run_epoch()
    avg, counts = model.evaluate(subset)
    assert avg.shape == (10,) and counts.shape == (8,)

    # oracle: same weights on the CPU, the reference's loop (gnn_model.py:58-74) with the pinned metric functions
    ref = sage_ref.GraphSageRef(20, [64, 64], 4)
    ref.load_state_dict({k: v.cpu() for k, v in model.net.state_dict().items()})
    w = torch.tensor([0.1, 1.0, 2.0, 2.0])
    rows, cnts = [], []
    for mri_id, G, feats, labels in ds:
        s, d = G.edges()
        csr = graph_ref.csr_by_dst_ref(s.numpy(), d.numpy(), G.number_of_nodes())[:2]
        with torch.no_grad():
            rl = ref(csr, torch.as_tensor(feats, dtype=torch.float32))
            loss = F.cross_entropy(rl, torch.as_tensor(labels), weight=w).item()
        pred = rl.argmax(1).numpy()
        vox = project_ref.project_nodes_to_img_ref(ds.get_supervoxel_partitioning(mri_id), pred)
        rows.append([loss] + evaluation.calculate_node_dices(pred, labels)
                    + evaluation.calculate_brats_metrics(vox, ds.get_voxel_labels(mri_id)))
        cnts.append(np.concatenate([evaluation.count_node_labels(pred), evaluation.count_node_labels(labels)]))
    ref_avg, ref_counts = np.mean(np.array(rows, dtype=np.float64), axis=0), np.sum(cnts, axis=0)
    assert np.array_equal(counts, ref_counts)                       # identical arg-max classes on every node
    assert abs(avg[0] - ref_avg[0]) < 1e-4 * max(1.0, abs(ref_avg[0]))
    assert np.array_equal(avg[1:], ref_avg[1:])                     # integer-derived metrics: exact
    assert np.isfinite(avg).all()

    # save_weights writes f"{folder}{name}.pt" (gnn_model.py:89-90); the file loads into the oracle and into a new net
    folder = str(tmp_path) + os.sep
    model.save_weights(folder, "ckpt")
    sd = torch.load(folder + "ckpt.pt", map_location="cpu")
    assert set(sd.keys()) == set(model.net.state_dict().keys())
    net2 = networks.GraphSage(20, [64, 64], 4, "pool", 0)
    net2.load_state_dict(sd)
    for (k, a), (_, b) in zip(sorted(sd.items()), sorted(net2.state_dict().items())):
        assert torch.equal(a, b), k


def test_reference_layout_checkpoint_loads(cuda_dev):
    """A DGL-era checkpoint (provided_gnn_weights.pt layout: fc_pool / fc_self / fc_neigh with the DGL<=0.7 bias
    placement) loads and gives the oracle's logits."""
    torch.manual_seed(1)
    ref = sage_ref.GraphSageRef(20, [32, 32], 4)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    # DGL <= 0.7 keeps ONE bias per layer under 'layers.i.bias' instead of fc_self.bias (Appendix A.3)
    legacy = {}
    for k, v in sd.items():
        legacy[k.replace("fc_self.bias", "bias")] = v
    net = networks.GraphSage(20, [32, 32], 4, "pool", 0)
    net.load_state_dict(legacy)
    net.to(cuda_dev).eval()
    from gnn_tumor_seg_b200 import graph as G
    g = synth.make_small_graph(5, n_nodes=200, avg_deg=6)
    bg = G.from_edge_list(g.src, g.dst, g.n_nodes)
    s, d = bg.edges()
    csr = graph_ref.csr_by_dst_ref(s.numpy(), d.numpy(), g.n_nodes)[:2]
    with torch.no_grad():
        out = net(bg.to(cuda_dev), torch.as_tensor(g.features).to(cuda_dev)).cpu()
        rl = ref(csr, torch.as_tensor(g.features))
    assert (out - rl).abs().max().item() <= 1e-4 * rl.abs().max().item()


def test_run_epoch_matches_oracle_training_loop(cuda_dev):
    """GNN.run_epoch (model/gnn_model.py:34-48) against the reference's loop restated on the CPU: same batches in the same
    order (shuffle off), oracle forward/backward + torch.optim.AdamW + ExponentialLR.  Two epochs: the mean epoch losses
    agree to 1e-4 and the learning rate decays like the reference's scheduler."""
    from gnn_tumor_seg_b200 import graph as G
    from gnn_tumor_seg_b200.gnn_model import GNN
    from gnn_tumor_seg_b200.graph import minibatch_graphs
    graphs = [synth.make_small_graph(60 + s, n_nodes=300 + 21 * s, avg_deg=8) for s in range(8)]
    samples = [(g.mri_id, G.from_edge_list(g.src, g.dst, g.n_nodes), g.features, g.labels) for g in graphs]
    hp = HP(20, 4, [64, 64], None, None, [0.1, 1.0, 2.0, 2.0], 1e-3, 1e-4, 0.9)
    torch.manual_seed(3)
    model = GNN("GSpool", hp, samples)
    assert model.trainer is not None                    # the one-call step is the path under test
    model.train_loader = torch.utils.data.DataLoader(samples, batch_size=6, shuffle=False, num_workers=0, collate_fn=minibatch_graphs)
    ref = sage_ref.GraphSageRef(20, [64, 64], 4)
    ref.load_state_dict({k: v.detach().cpu().clone() for k, v in model.net.state_dict().items()})
    ropt = torch.optim.AdamW(ref.parameters(), lr=hp.lr, weight_decay=hp.w_decay)
    rsched = torch.optim.lr_scheduler.ExponentialLR(ropt, hp.lr_decay, last_epoch=-1)
    w = torch.tensor(hp.class_weights)
    for epoch in range(2):
        got = model.run_epoch()
        losses = []
        for lo in range(0, len(samples), 6):
            _, bg, feats, labels = minibatch_graphs(samples[lo:lo + 6])
            s, d = bg.edges()
            csr = graph_ref.csr_by_dst_ref(s.numpy(), d.numpy(), bg.number_of_nodes())[:2]
            loss = F.cross_entropy(ref(csr, feats), labels, weight=w)
            losses.append(loss.item())
            ropt.zero_grad(); loss.backward(); ropt.step()
        rsched.step()
        assert abs(got - float(np.mean(losses))) <= 1e-4 * abs(float(np.mean(losses))), (epoch, got, float(np.mean(losses)))
        assert abs(model.optimizer.param_groups[0]["lr"] - ropt.param_groups[0]["lr"]) < 1e-12
    # after two epochs (4 AdamW steps) the parameters still agree on average (Adam normalises rounding-level gradients)
    for (n, p), (_, q) in zip(model.net.named_parameters(), ref.named_parameters()):
        assert (p.detach().cpu() - q.detach()).abs().mean().item() < 2e-5, n
