"""gnn_tumor_seg_b200.evaluation against outputs of the reference's own model/evaluation.py
(tests/golden/make_golden_eval.py ran it; /root/reference is not needed at test time)."""
import os

import numpy as np

from gnn_tumor_seg_b200 import evaluation as ev

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_eval.npz"))


def test_node_dices_and_counts_match_reference():
    for i in range(int(G["n_node"])):
        pr, la = G[f"node_pred_{i}"], G[f"node_true_{i}"]
        assert np.array_equal(np.array(ev.calculate_node_dices(pr, la), dtype=np.float64), G[f"node_dice_{i}"]), i
        assert np.array_equal(ev.count_node_labels(pr), G[f"node_cnt_{i}"]), i


def test_survey_golden_vector_6():
    assert np.allclose(ev.calculate_node_dices([0, 1, 2, 3, 3], [0, 1, 3, 3, 0]), [0.8571428571428571, 0.8, 0.5], rtol=0, atol=1e-15)
    assert ev.count_node_labels([0, 1, 2, 3, 3]).tolist() == [1, 1, 1, 2]


def test_brats_metrics_match_reference_including_empty_regions():
    for i in range(int(G["n_vox"])):
        a, b = G[f"vox_pred_{i}"], G[f"vox_true_{i}"]
        got = np.array(ev.calculate_brats_metrics(a, b), dtype=np.float64)
        assert np.array_equal(got, G[f"vox_metrics_{i}"]), (i, got, G[f"vox_metrics_{i}"])
        acc = G[f"vox_acc_{i}"]
        assert ev.compute_accuracy(a, b, True) == acc[0]
        if len(acc) > 1:
            assert ev.compute_accuracy(a, b, False) == acc[1]


def test_empty_conventions():
    z = np.zeros((4, 4, 4), dtype=np.int64)
    one = z.copy(); one[1, 1, 1] = 3
    assert ev.calculate_brats_metrics(z, z) == [1, 1, 1, 0, 0, 0]
    m = ev.calculate_brats_metrics(one, z)
    assert m[:3] == [0.0, 0.0, 0.0] and m[3:] == [300, 300, 300]
