"""BraTS region metrics with the reference's definitions (model/evaluation.py:22-106):
node / voxel Dice of the three nested regions WT (label != 0), CT (labels {2,3}) and ET (label 3) on the
reference's internal labels (0 healthy, 1 edema, 2 non-enhancing, 3 enhancing), the 95th-percentile symmetric
Hausdorff distance of the same regions, and the per-class node counts.

Host-side numpy / scipy like the reference (these are evaluation statistics, not the hot path); pinned to outputs
of the reference's own functions by tests/golden/make_golden_eval.py -> tests/test_evaluation.py.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

HEALTHY, EDEMA, NET, ET = 0, 1, 2, 3
HD_BOTH_EMPTY = 0        # region absent from prediction and ground truth: the prediction was right (evaluation.py:88-89)
HD_ONE_EMPTY = 300       # region absent from exactly one of them: maximal distance (evaluation.py:91-92)


def count_node_labels(preds_or_labels):
    """[#healthy, #edema, #net, #et] as float64 (model/evaluation.py:22-26)."""
    return np.bincount(np.asarray(preds_or_labels).reshape(-1).astype(np.int64), minlength=4)[:4].astype(np.float64)


def _regions(x):
    x = np.asarray(x)
    return x != HEALTHY, (x == NET) | (x == ET), x == ET


def dice_from_masks(pred, truth):
    """2TP / (2TP + FP + FN); 1 when the region is absent from both (model/evaluation.py:98-106)."""
    tp = np.count_nonzero(pred & truth)
    fp = np.count_nonzero(pred & ~truth)
    fn = np.count_nonzero(~pred & truth)
    if tp + fp + fn == 0:
        return 1
    return (2 * tp) / (2 * tp + fp + fn)


def calculate_node_dices(preds, labels):
    """[WT, CT, ET] Dice over the nodes of one brain (model/evaluation.py:30-45)."""
    return [dice_from_masks(p, t) for p, t in zip(_regions(preds), _regions(labels))]


def _surface_distances(a, b, connectivity=1):
    """Distances from the border voxels of a to the nearest border voxel of b (the reference's copy of medpy's
    __surface_distances, model/evaluation.py:150-182)."""
    footprint = ndimage.generate_binary_structure(a.ndim, connectivity)
    a_border = a ^ ndimage.binary_erosion(a, structure=footprint, iterations=1)
    b_border = b ^ ndimage.binary_erosion(b, structure=footprint, iterations=1)
    dt = ndimage.distance_transform_edt(~b_border)
    return dt[a_border]


def hd95_from_masks(pred, truth):
    """95th percentile of the symmetric surface distances, with the reference's empty-region conventions
    (model/evaluation.py:82-95,111-145)."""
    pred = np.atleast_1d(np.asarray(pred, dtype=bool))
    truth = np.atleast_1d(np.asarray(truth, dtype=bool))
    if not pred.any() or not truth.any():
        return HD_BOTH_EMPTY if not pred.any() and not truth.any() else HD_ONE_EMPTY
    return np.percentile(np.hstack((_surface_distances(pred, truth), _surface_distances(truth, pred))), 95)


def calculate_brats_metrics(predicted_voxels, true_voxels):
    """[WT, CT, ET Dice, WT, CT, ET HD95] over the voxels of one brain (model/evaluation.py:63-80)."""
    pr, tr = _regions(predicted_voxels), _regions(true_voxels)
    return [dice_from_masks(p, t) for p, t in zip(pr, tr)] + [hd95_from_masks(p, t) for p, t in zip(pr, tr)]


def compute_accuracy(supervoxel_labelling, ground_truth, include_healthy=True):
    """Fraction of voxels predicted correctly, optionally over the non-healthy ground truth only
    (model/evaluation.py:49-58)."""
    a, g = np.asarray(supervoxel_labelling), np.asarray(ground_truth)
    assert g.shape == a.shape
    if include_healthy:
        return np.sum(g == a) / g.size
    mask = g != 0
    return np.sum((g == a) & mask) / np.sum(mask)
