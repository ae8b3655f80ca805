"""Oracle: weighted-mean cross entropy (model/gnn_model.py:30,42).
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
import torch


def weighted_ce_ref(logits, labels, class_weights):
    """L = sum_i w[y_i] * (-log_softmax(z_i)[y_i]) / sum_i w[y_i]  — what
    torch.nn.CrossEntropyLoss(weight=w) computes."""
    logp = torch.log_softmax(logits, dim=1)
    w = class_weights[labels]
    nll = -logp.gather(1, labels.view(-1, 1)).squeeze(1)
    return (w * nll).sum() / w.sum()
