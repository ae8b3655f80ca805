"""Training / evaluation wrapper with the reference's entry points
(model/gnn_model.py:21-90): ``GNN(model_type, hyperparameters, train_dataset)``
with ``.net .optimizer .lr_decay .loss_fcn .train_loader .device``,
``run_epoch()``, ``evaluate(dataset)``, ``save_weights(folder, name)``.

The dataset contract is the reference's ImageGraphDataset one: iterating yields
``(mri_id, graph, features, labels)`` where ``graph`` is a host
``BatchedGraph`` (gnn_tumor_seg_b200.graph) instead of a DGLGraph.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import DataLoader

from . import evaluation, ops
from ._lib import GtsError
from .data_loader import DevicePrefetcher
from .graph import minibatch_graphs
from .networks import init_graph_net
from .project import project_nodes_to_img
from .trainer import FusedAdamW, SageTrainer

BATCH_SIZE = 6          # model/gnn_model.py:12


class WeightedCrossEntropy(torch.nn.Module):
    """torch.nn.CrossEntropyLoss(weight=w) computed by the fused K8 kernel."""

    def __init__(self, weight):
        super().__init__()
        self.register_buffer("weight", weight)

    def forward(self, logits, labels):
        return ops.weighted_cross_entropy(logits, labels, self.weight)


class GNN:
    def __init__(self, model_type, hyperparameters, train_dataset):
        if not torch.cuda.is_available():
            raise GtsError("gnn_tumor_seg_b200.GNN needs a CUDA device (B200); there is no CPU path")
        self.device = torch.device('cuda')
        print("Using device", self.device)
        class_weights = torch.FloatTensor(hyperparameters.class_weights).to(self.device)
        self.net = init_graph_net(model_type, hyperparameters)
        self.net.to(self.device)
        # AdamW as one kernel over a flat parameter arena (gts_adamw_step_dev); a torch.optim.Optimizer, so the
        # reference's ExponentialLR drives it unchanged (model/gnn_model.py:28-29).  GraphSage('pool') without
        # dropout trains through the one-call step (trainer.SageTrainer: forward + CE + backward in gts_sage_step);
        # every other network through autograd + the same optimiser.
        self.trainer = None
        try:
            self.trainer = SageTrainer(self.net, class_weights, lr=hyperparameters.lr, weight_decay=hyperparameters.w_decay)
            self.optimizer = self.trainer.optimizer
        except GtsError:
            self.optimizer = FusedAdamW(self.net.parameters(), lr=hyperparameters.lr, weight_decay=hyperparameters.w_decay)
        self.lr_decay = torch.optim.lr_scheduler.ExponentialLR(self.optimizer, hyperparameters.lr_decay, last_epoch=-1)
        self.loss_fcn = WeightedCrossEntropy(class_weights)
        self.train_loader = DataLoader(train_dataset, batch_size=BATCH_SIZE, shuffle=True, num_workers=0,
                                       collate_fn=minibatch_graphs) if train_dataset is not None else None

    def run_epoch(self):
        self.net.train()
        losses = []
        # self.prefetch = True stages batch i+1 (H2D + CSR build) on a side stream while batch i computes; the default
        # keeps the reference's in-stream .to(device) (model/gnn_model.py:38-40)
        loader = DevicePrefetcher(self.train_loader, self.device) if getattr(self, "prefetch", False) else self.train_loader
        for batch_mris, batch_graphs, batch_features, batch_labels in loader:
            batch_graphs = batch_graphs.to(self.device)
            batch_features = batch_features.to(self.device)
            batch_labels = batch_labels.to(self.device)
            if self.trainer is not None and ops.use_stack_path():
                loss = self.trainer.step(batch_graphs, batch_features, batch_labels)
            else:
                logits = self.net(batch_graphs, batch_features)
                loss = self.loss_fcn(logits, batch_labels)
                self.optimizer.zero_grad()
                loss.backward()
                self.optimizer.step()
            losses.append(loss.detach())          # read back once per epoch, not per step
        self.lr_decay.step()
        return float(torch.stack(losses).mean().item()) if losses else float("nan")

    # must be a Subset of an ImageGraphDataset (or any object with the same protocol)
    def evaluate(self, dataset):
        base = getattr(dataset, "dataset", dataset)
        assert getattr(base, "read_label", True) == True
        self.net.eval()
        # metrics: loss, 3 node dices (WT, CT, ET), 3 voxel dices, 3 voxel HD95
        metrics = np.zeros((len(dataset), 10))
        counts = np.zeros((len(dataset), 8))
        i = 0
        for curr_id, curr_graph, curr_feats, curr_labels in dataset:
            curr_graph = curr_graph.to(self.device)
            curr_feats = torch.FloatTensor(curr_feats).to(self.device)
            curr_labels = torch.LongTensor(curr_labels).to(self.device)
            with torch.no_grad():
                logits = self.net(curr_graph, curr_feats)
                loss = self.loss_fcn(logits, curr_labels)
            _, predicted_classes = torch.max(logits, dim=1)
            predicted_classes = predicted_classes.detach().cpu().numpy()
            metrics[i][0] = loss.item()
            ct, res = self.calculate_all_metrics_for_brain(curr_id, dataset, predicted_classes,
                                                           curr_labels.detach().cpu().numpy())
            metrics[i][1:] = res
            counts[i] = ct
            i += 1
        avg_metrics = np.mean(metrics, axis=0)
        total_counts = np.sum(counts, axis=0)
        return avg_metrics, total_counts

    def calculate_all_metrics_for_brain(self, mri_id, dataset, node_preds, node_labels):
        """model/gnn_model.py:76-87: label counts (predicted, true), node-wise WT/CT/ET Dice, voxel-wise WT/CT/ET
        Dice and HD95 with the reference's definitions (gnn_tumor_seg_b200.evaluation, pinned to outputs of the
        reference's model/evaluation.py).  The projection to voxels runs on the device (K7).  A dataset without
        voxel accessors (in-memory lists of graphs) yields nan in the six voxel slots instead of raising."""
        label_counts = np.concatenate([evaluation.count_node_labels(node_preds), evaluation.count_node_labels(node_labels)])
        node_dices = evaluation.calculate_node_dices(node_preds, node_labels)
        voxel_metrics = np.full(6, np.nan)
        base = getattr(dataset, "dataset", dataset)
        if hasattr(base, "get_supervoxel_partitioning") and hasattr(base, "get_voxel_labels"):
            sv_partitioning = base.get_supervoxel_partitioning(mri_id)
            true_voxels = base.get_voxel_labels(mri_id)
            pred_voxels = project_nodes_to_img(sv_partitioning, node_preds)
            voxel_metrics = evaluation.calculate_brats_metrics(pred_voxels, true_voxels)
        return label_counts, np.concatenate([node_dices, voxel_metrics])

    def save_weights(self, folder, name):
        torch.save(self.net.state_dict(), f"{folder}{name}.pt")
