"""Golden vectors for gnn_tumor_seg_b200.evaluation, produced by RUNNING the reference's own functions
(/root/reference/model/evaluation.py) in this container.  Run once: python tests/golden/make_golden_eval.py
-> tests/golden/reference_eval.npz."""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from model import evaluation as ev  # noqa: E402

rng = np.random.default_rng(7)
out = {}
cases = []
# node-level: random label vectors with and without missing classes
for i, (n, p) in enumerate([(50, [0.7, 0.1, 0.1, 0.1]), (200, [0.9, 0.1, 0.0, 0.0]), (30, [1.0, 0, 0, 0]), (64, [0.25] * 4),
                            (5, None)]):
    if p is None:
        pr, la = np.array([0, 1, 2, 3, 3]), np.array([0, 1, 3, 3, 0])          # SURVEY §8c golden 6
    else:
        pr, la = rng.choice(4, size=n, p=p), rng.choice(4, size=n, p=p[::-1] if i == 3 else p)
    out[f"node_pred_{i}"], out[f"node_true_{i}"] = pr, la
    out[f"node_dice_{i}"] = np.array(ev.calculate_node_dices(pr, la), dtype=np.float64)
    out[f"node_cnt_{i}"] = ev.count_node_labels(pr)
out["n_node"] = 5
# voxel-level: blobs, disjoint blobs, an absent class in one / both volumes, 2-D case
def blobs(shape, specs):
    v = np.zeros(shape, dtype=np.int64)
    for lab, sl in specs:
        v[sl] = lab
    return v
vox = [
    (blobs((12, 11, 10), [(1, np.s_[2:9, 2:9, 2:8]), (2, np.s_[3:7, 3:7, 3:6]), (3, np.s_[4:6, 4:6, 4:5])]),
     blobs((12, 11, 10), [(1, np.s_[3:10, 2:8, 2:8]), (2, np.s_[4:8, 3:7, 3:7]), (3, np.s_[5:7, 4:6, 4:6])])),
    (blobs((10, 10, 10), [(1, np.s_[0:3, 0:3, 0:3])]), blobs((10, 10, 10), [(3, np.s_[6:9, 6:9, 6:9])])),
    (blobs((9, 9, 9), [(1, np.s_[2:6, 2:6, 2:6])]), blobs((9, 9, 9), [(1, np.s_[2:6, 2:6, 2:6])])),
    (np.zeros((6, 6, 6), dtype=np.int64), np.zeros((6, 6, 6), dtype=np.int64)),
    (blobs((16, 16), [(2, np.s_[3:9, 3:9]), (3, np.s_[5:7, 5:7])]), blobs((16, 16), [(2, np.s_[4:11, 2:9]), (1, np.s_[12:15, 12:15])])),
    (rng.choice(4, size=(8, 9, 7), p=[0.6, 0.2, 0.1, 0.1]), rng.choice(4, size=(8, 9, 7), p=[0.6, 0.2, 0.1, 0.1])),
]
for i, (a, b) in enumerate(vox):
    out[f"vox_pred_{i}"], out[f"vox_true_{i}"] = a, b
    out[f"vox_metrics_{i}"] = np.array(ev.calculate_brats_metrics(a, b), dtype=np.float64)
    out[f"vox_acc_{i}"] = np.array([ev.compute_accuracy(a, b, True)] + ([ev.compute_accuracy(a, b, False)] if (b != 0).any() else []))
out["n_vox"] = len(vox)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_eval.npz"), **out)
print("wrote reference_eval.npz", {k: v.tolist() for k, v in out.items() if k.startswith("vox_metrics")})
