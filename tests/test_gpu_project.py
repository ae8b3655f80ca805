"""GPU parity: K7 reprojection — bit-exact against the reference-generated
golden fixture and the oracle, plus full-size properties."""
import os

import numpy as np
import pytest
import torch

from gnn_tumor_seg_b200 import project, synth
from oracle import project_ref

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_kat.npz"))


def test_project_nodes_golden(cuda_dev):
    out = project.project_nodes_to_img(GOLD["part"], np.array([3, 1, 2]))
    assert isinstance(out, np.ndarray) and out.dtype == np.int64
    assert np.array_equal(out, GOLD["proj"])
    assert np.array_equal(project.project_nodes_to_img(GOLD["part3"], GOLD["labels3"]), GOLD["proj3"])
    t = project.project_nodes_to_img(torch.as_tensor(GOLD["part3"]).to(cuda_dev), torch.as_tensor(GOLD["labels3"]).to(cuda_dev))
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), GOLD["proj3"])


def test_project_nodes_out_of_range_raises(cuda_dev):
    with pytest.raises(IndexError):
        project.project_nodes_to_img(np.array([[0, 5, -1]], dtype=np.int16), np.array([1, 2]))
    empty = project.project_nodes_to_img(np.zeros((0, 3), np.int16), np.array([1]))
    assert empty.shape == (0, 3)


def test_save_voxel_preds_chain_golden(cuda_dev):
    crop = (GOLD["crop_ix0"], GOLD["crop_ix1"], GOLD["crop_ix2"])      # non-contiguous np.ix_ crop
    vol = project.project_labels_to_brats(torch.as_tensor(GOLD["logits"]).to(cuda_dev), GOLD["svs_c"], crop)
    project.check_projection(vol)
    full = vol.cpu().numpy()
    assert full.dtype == np.int16 and full.shape == (240, 240, 155)
    nz = np.flatnonzero(full)
    assert np.array_equal(nz, GOLD["vox_brats_nonzero_idx"])
    assert np.array_equal(full.reshape(-1)[nz], GOLD["vox_brats_nonzero_val"])
    # classes instead of logits
    vol2 = project.project_labels_to_brats(torch.as_tensor(GOLD["pred_nodes"]).to(cuda_dev), GOLD["svs_c"], crop)
    assert torch.equal(vol, vol2)
    # a class outside the relabel table -> the reference's RuntimeError('unexpected label')
    bad = GOLD["pred_nodes"].copy(); bad[int(GOLD["svs_c"].max())] = 7
    with pytest.raises(RuntimeError, match="unexpected label"):
        project.check_projection(project.project_labels_to_brats(torch.as_tensor(bad).to(cuda_dev), GOLD["svs_c"], crop))


def test_save_voxel_logits_golden(cuda_dev):
    out = project.project_logits_to_img(torch.as_tensor(GOLD["logits"]).to(cuda_dev), GOLD["svs_c"])
    assert out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy().astype(np.float64), GOLD["vox_logits"])   # reference emits float64 copies


def test_full_size_synthetic_volume(cuda_dev):
    """Config-5 size: 15k-node partition -> (240,240,155) int16, vs the oracle, plus
    structural properties (outside-crop zeros, label histogram through the LUT)."""
    g = synth.make_graph(0, with_partition=True)
    rng = np.random.default_rng(0)
    logits = rng.normal(size=(g.n_nodes, 4)).astype(np.float32)
    vol = project.check_projection(project.project_labels_to_brats(torch.as_tensor(logits).to(cuda_dev), g.svs, g.crop)).cpu().numpy()
    ref = project_ref.save_voxel_preds_ref(logits, g.svs, g.crop)
    assert np.array_equal(vol, ref)
    assert set(np.unique(vol)) <= {0, 1, 2, 4}
    inside = np.zeros((240, 240, 155), bool); inside[g.crop] = True
    assert not vol[~inside].any()
    cls = logits.argmax(1)
    counts = np.bincount(g.svs[g.svs >= 0].astype(np.int64), minlength=g.n_nodes)
    for c, lab in enumerate(project.BRATS_LABEL_LUT):
        if lab:
            assert (vol == lab).sum() == counts[cls == c].sum()


GOLD_CROP = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_kat_crop.npz"))


@pytest.mark.parametrize("case", ["pred", "healthy", "border", "single"])
def test_determine_tumor_crop_golden(cuda_dev, case):
    """Plane-occupancy kernel vs the crops the reference's determine_tumor_crop returned
    (tests/golden/make_golden.py): every voxel is its own supervoxel, its class = the predicted label."""
    vol = GOLD_CROP[f"tcrop_{case}_vol"]
    flat = vol.reshape(-1)
    svs = np.arange(flat.size, dtype=np.int64).reshape(vol.shape)
    if flat.size > 32767:
        pytest.skip("fixture volume exceeds the int16 supervoxel id range")
    ix = project.determine_tumor_crop(svs.astype(np.int16), flat.astype(np.int64))
    for a in range(3):
        assert np.array_equal(np.asarray(ix[a]).reshape(-1), GOLD_CROP[f"tcrop_{case}_{a}"])
        assert ix[a].ndim == 3                                   # np.ix_ shape, usable as vol[ix]


def test_determine_tumor_crop_full_size_vs_oracle(cuda_dev):
    """Config-5 size: 15k-node partition, random node logits; crop == the oracle's crop of the voxel predictions."""
    g = synth.make_graph(1, with_partition=True)
    rng = np.random.default_rng(5)
    logits = rng.normal(size=(g.n_nodes, 4)).astype(np.float32)
    logits[:, 0] += 3.0                                          # mostly healthy, a few tumour supervoxels
    ix = project.determine_tumor_crop(g.svs, torch.as_tensor(logits).to(cuda_dev))
    ref = project_ref.determine_tumor_crop_ref(project_ref.project_nodes_to_img_ref(g.svs, logits.argmax(1)))
    for a in range(3):
        assert np.array_equal(ix[a], ref[a])
    # nothing tumorous -> the whole volume
    ix0 = project.determine_tumor_crop(g.svs, np.zeros(g.n_nodes, dtype=np.int64))
    assert tuple(a.size for a in ix0) == g.svs.shape
    with pytest.raises(IndexError):
        project.determine_tumor_crop(np.full((2, 2, 2), 9, np.int16), np.array([1, 0]))
