"""The training / evaluation loop of model/gnn_model.py pinned to the REFERENCE'S OWN CLASS: tests/golden/reference_gnn.npz
holds what /root/reference/model/gnn_model.py::GNN (run unmodified on the CPU over stubbed DGL / nibabel, see
tests/golden/make_golden_gnn.py) returned from run_epoch x 3, evaluate and save_weights on an 8-brain dataset written
with the reference's own save_networkx_graph.

Here (no GPU): the loop as this repository restates it — oracle network + torch.optim.AdamW(lr, weight_decay=hp.w_decay)
+ ExponentialLR per epoch + the metric functions of gnn_tumor_seg_b200.evaluation, the SAME restatement
tests/test_gpu_gnn_api.py holds the CUDA ``GNN`` class against — reproduces the reference's numbers, reading the dataset
through the product's own ImageGraphDataset / minibatch_graphs.  Reference GNN == restated loop (this file) and
restated loop == CUDA GNN (test_gpu_gnn_api.py) together pin SURVEY §8 rows a10-a12 to the reference's code."""
import ast
import os

import numpy as np
import torch
import torch.nn.functional as F

from gnn_tumor_seg_b200 import evaluation
from gnn_tumor_seg_b200.data_loader import ImageGraphDataset
from gnn_tumor_seg_b200.graph import minibatch_graphs
from oracle import graph_ref, project_ref, sage_ref

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_gnn.npz"))
HP_FIELDS = "n_epochs in_feats out_classes lr lr_decay w_decay class_weights layer_sizes feature_dropout gat_heads gat_residuals".split()


def _dataset(root):
    for mri in [str(x) for x in GOLD["ids"]]:
        os.makedirs(os.path.join(root, mri))
        with open(os.path.join(root, mri, f"{mri}_nxgraph.json"), "wb") as f:
            f.write(GOLD[f"{mri}/json"].tobytes())                   # the text the reference's save_networkx_graph wrote
        np.save(os.path.join(root, mri, f"{mri}_supervoxels.npy"), GOLD[f"{mri}/svs"])
        np.save(os.path.join(root, mri, f"{mri}_label.npy"), GOLD[f"{mri}/truth"])
    ds = ImageGraphDataset(root + os.sep, "BraTS", read_image=False, read_graph=True, read_label=True)
    ds.all_ids = sorted(ds.all_ids)
    return ds


def _csr(g):
    s, d = g.edges()
    return graph_ref.csr_by_dst_ref(s.numpy(), d.numpy(), g.number_of_nodes())[:2]


def test_restated_training_and_evaluation_loop_reproduces_the_reference_gnn_class(tmp_path):
    hp = dict(zip(HP_FIELDS, ast.literal_eval(str(GOLD["hp"]))))
    ds = _dataset(str(tmp_path / "data"))
    assert ds.all_ids == [str(x) for x in GOLD["ids"]] and len(ds) == 8
    assert [ast.literal_eval(str(r))[1:3] for r in GOLD["layers"]] == [(20, 32), (32, 16), (16, 4)]

    ref = sage_ref.GraphSageRef(hp["in_feats"], hp["layer_sizes"], hp["out_classes"])
    ref.load_state_dict({k[len("init_sd/"):]: torch.as_tensor(GOLD[k]) for k in GOLD.files if k.startswith("init_sd/")})
    opt = torch.optim.AdamW(ref.parameters(), lr=hp["lr"], weight_decay=hp["w_decay"])
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, hp["lr_decay"], last_epoch=-1)
    w = torch.tensor(hp["class_weights"], dtype=torch.float32)
    samples = [ds[i] for i in range(len(ds))]
    bs = int(GOLD["batch_size"])
    assert bs == 6                                                       # model/gnn_model.py:11
    for epoch in range(hp["n_epochs"]):
        losses = []
        for lo in range(0, len(samples), bs):
            _, bg, feats, labels = minibatch_graphs(samples[lo:lo + bs])
            loss = F.cross_entropy(ref(_csr(bg), feats), labels, weight=w)
            losses.append(loss.item())
            opt.zero_grad()
            loss.backward()
            opt.step()
        sched.step()
        want = float(GOLD["epoch_losses"][epoch])
        assert abs(float(np.mean(losses)) - want) <= 1e-5 * abs(want), (epoch, float(np.mean(losses)), want)
        assert abs(opt.param_groups[0]["lr"] - float(GOLD["epoch_lrs"][epoch])) < 1e-15
    for k, v in ref.state_dict().items():
        assert torch.allclose(v, torch.as_tensor(GOLD["final_sd/" + k]), rtol=1e-4, atol=1e-6), k

    # evaluate (model/gnn_model.py:51-74) on the same Subset
    ref.eval()
    rows, cnts = [], []
    for i in [int(x) for x in GOLD["eval_subset"]]:
        mri_id, g, feats, labels = ds[i]
        with torch.no_grad():
            logits = ref(_csr(g), torch.as_tensor(feats, dtype=torch.float32))
            loss = F.cross_entropy(logits, torch.as_tensor(labels), weight=w).item()
        pred = logits.argmax(1).numpy()
        vox = project_ref.project_nodes_to_img_ref(ds.get_supervoxel_partitioning(mri_id), pred)
        rows.append([loss] + list(evaluation.calculate_node_dices(pred, labels))
                    + list(evaluation.calculate_brats_metrics(vox, ds.get_voxel_labels(mri_id))))
        cnts.append(np.concatenate([evaluation.count_node_labels(pred), evaluation.count_node_labels(labels)]))
    avg, counts = np.mean(np.array(rows, dtype=np.float64), axis=0), np.sum(cnts, axis=0)
    assert np.array_equal(counts, GOLD["eval_counts"])
    assert abs(avg[0] - GOLD["eval_avg"][0]) <= 1e-5 * abs(GOLD["eval_avg"][0])
    assert np.allclose(avg[1:], GOLD["eval_avg"][1:], rtol=0, atol=1e-12)

    # save_weights: f"{folder}{name}.pt" holding net.state_dict() with DGL's keys
    assert bool(GOLD["ckpt_equals_net"])
    assert [str(k) for k in GOLD["ckpt_keys"]] == sorted(ref.state_dict().keys())
