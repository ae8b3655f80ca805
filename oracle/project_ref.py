"""Oracle: node -> voxel reprojection chain (numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED against fixtures made by
running the reference's own functions (tests/golden/make_golden.py).
"""
import numpy as np

BRATS_SHAPE = (240, 240, 155)
DEFAULT_BACKGROUND_NODE_LOGITS = [[1.0, -1.0, -1.0, -1.0]]     # utils/hyperparam_helpers.py:25
LABEL_MAP = {4: 3, 2: 1, 1: 2}                                  # scripts/preprocess_dataset.py:15


def project_nodes_to_img_ref(svs, node_labels):
    """data_processing/graph_io.py:21-24: append a 0 and fancy-index, so the
    background id -1 wraps onto the appended healthy label."""
    node_labels = np.append(node_labels, 0)
    return node_labels[svs]


def uncrop_to_brats_size_ref(crop, voxel_preds):
    """data_processing/image_processing.py:21-25."""
    out = np.zeros(BRATS_SHAPE, dtype=np.int16)
    out[crop] = voxel_preds
    return out


def swap_labels_to_brats_ref(label_data):
    """scripts/preprocess_dataset.py:159-169: 3->4, 1->2, 2->1; raises
    RuntimeError('unexpected label') on anything outside {0,1,2,3}."""
    for u in np.unique(label_data):
        if u not in [0, 1, 2, 3]:
            raise RuntimeError("unexpected label")
    new = np.zeros_like(label_data, dtype=np.int16)
    new[label_data == LABEL_MAP[4]] = 4
    new[label_data == LABEL_MAP[2]] = 2
    new[label_data == LABEL_MAP[1]] = 1
    return new


def save_voxel_preds_ref(node_logits, svs, crop):
    """scripts/generate_gnn_predictions.py:64-73 up to (not including) the
    NIfTI write: argmax (first maximum, as torch.max) -> project -> uncrop ->
    BraTS relabel.  Returns int16 (240,240,155)."""
    pred = np.argmax(np.asarray(node_logits), axis=1)
    vox = project_nodes_to_img_ref(svs, pred)
    vox = uncrop_to_brats_size_ref(crop, vox)
    return swap_labels_to_brats_ref(vox)


def save_voxel_logits_ref(node_logits, svs):
    """scripts/generate_gnn_predictions.py:55-62: background row appended,
    fancy-indexed by the supervoxel map -> float [X,Y,Z,C]."""
    node_logits = np.concatenate([np.asarray(node_logits), DEFAULT_BACKGROUND_NODE_LOGITS])
    return node_logits[svs]


def determine_tumor_crop_ref(preds):
    """data_processing/image_processing.py:8-17: dilate the predicted-tumour mask by one 3-D cross step
    (scipy.ndimage.binary_dilation defaults) and keep the planes that hold a voxel of it; the whole volume
    when nothing is predicted tumorous."""
    from scipy import ndimage
    mask = np.asarray(preds) != 0
    mask = ndimage.binary_dilation(mask)
    if np.all(~mask):
        mask = ~mask
    return np.ix_(mask.any(axis=(1, 2)), mask.any(axis=(0, 2)), mask.any(axis=(0, 1)))
