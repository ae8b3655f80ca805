"""Node -> voxel reprojection on the device (K7).

Drop-ins for data_processing/graph_io.py:21-24 (``project_nodes_to_img``) and
for the call sequence of scripts/generate_gnn_predictions.py:55-73
(``save_voxel_preds`` / ``save_voxel_logits`` minus the NIfTI write).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

BRATS_SHAPE = (240, 240, 155)                       # data_processing/image_processing.py:24
DEFAULT_BACKGROUND_NODE_LOGITS = [1.0, -1.0, -1.0, -1.0]   # utils/hyperparam_helpers.py:25
# swap_labels_to_brats (scripts/preprocess_dataset.py:15,159-169): 0->0, 1->2, 2->1, 3->4
BRATS_LABEL_LUT = (0, 2, 1, 4)


def _device():
    if not torch.cuda.is_available():
        raise _lib.GtsError("gnn_tumor_seg_b200.project needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


_LUT_CACHE = {}


def _lut_on(device, label_lut):
    """The relabelling table as a cached int16 device tensor (one H2D copy per device and table, not per call)."""
    key = (str(device), tuple(int(v) for v in label_lut))
    t = _LUT_CACHE.get(key)
    if t is None:
        t = torch.as_tensor(np.asarray(label_lut, dtype=np.int16)).to(device)
        _LUT_CACHE[key] = t
    return t


def _as_dev(x, dtype, device):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x)).to(device=device, dtype=dtype, non_blocking=True)


def project_nodes_to_img(svs, node_labels):
    """reference data_processing/graph_io.py:21-24: every voxel gets the label of
    its supervoxel; background (-1) gets 0.  numpy in -> numpy int64 out (as the
    reference); CUDA tensors in -> CUDA int64 tensor out (no host round trip).
    Raises IndexError on a supervoxel id outside [-1, len(node_labels))."""
    lib = _lib.load()
    numpy_out = not torch.is_tensor(svs)
    dev = svs.device if (torch.is_tensor(svs) and svs.is_cuda) else _device()
    if numpy_out:
        big = np.asarray(svs)
        if big.size and (big.max() > 32767 or big.min() < -32768):
            raise IndexError("supervoxel ids must fit int16 (mri2graph/graphgen.py:77,243)")
    svs_d = _as_dev(svs, torch.int16, dev)
    lab_d = _as_dev(node_labels, torch.int64, dev)
    out = torch.empty(svs_d.shape, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.gts_project_nodes(ptr(svs_d), svs_d.numel(), ptr(lab_d), lab_d.numel(), ptr(out), ptr(err),
                                    stream_ptr()), "gts_project_nodes")
    if numpy_out:
        res = out.cpu().numpy()
        if int(err.item()):
            raise IndexError(f"supervoxel id out of bounds for {lab_d.numel()} node labels")
        return res
    return out


def crop_inverse_maps(crop, full_shape=BRATS_SHAPE, device=None):
    """Three int32 device vectors: crop coordinate of every full-volume plane, or -1.
    ``crop`` is the np.ix_ tuple the preprocessor saves (``_crop.npy``)."""
    device = device or _device()
    maps = []
    for axis, (ix, n) in enumerate(zip(crop, full_shape)):
        ix = np.asarray(ix).reshape(-1)
        if ix.dtype == np.bool_:
            ix = np.flatnonzero(ix)
        inv = np.full(n, -1, dtype=np.int32)
        inv[ix] = np.arange(ix.size, dtype=np.int32)
        maps.append(torch.as_tensor(inv).to(device, non_blocking=True))
    return tuple(maps)


def project_labels_to_brats(node_logits_or_classes, svs, crop, full_shape=BRATS_SHAPE,
                            label_lut=BRATS_LABEL_LUT, out=None, inv_maps=None):
    """save_voxel_preds (scripts/generate_gnn_predictions.py:64-73) without the
    file write: argmax (first maximum) -> project through the cropped int16 map
    -> paste into the (240,240,155) volume -> BraTS relabel.  Returns an int16
    CUDA tensor of ``full_shape``; ``.cpu().numpy()`` it to save.
    Raises RuntimeError('unexpected label') like swap_labels_to_brats when a
    class falls outside the relabelling table."""
    lib = _lib.load()
    x = node_logits_or_classes
    dev = x.device if (torch.is_tensor(x) and x.is_cuda) else _device()
    with torch.cuda.device(dev):
        st = stream_ptr()
        if torch.is_tensor(x) and x.dim() == 2 or (not torch.is_tensor(x) and np.asarray(x).ndim == 2):
            logits = _as_dev(x, torch.float32, dev)
            n_nodes, n_cls = logits.shape
            cls = torch.empty(n_nodes, dtype=torch.int32, device=dev)
            check(lib.gts_argmax_rows(ptr(logits), logits.stride(0), n_nodes, n_cls, ptr(cls), st), "gts_argmax_rows")
        else:
            cls = _as_dev(x, torch.int32, dev)
            n_nodes = cls.numel()
        svs_d = _as_dev(svs, torch.int16, dev)
        X, Y, Z = svs_d.shape
        if inv_maps is None:
            inv_maps = crop_inverse_maps(crop, full_shape, dev)
        lut = _lut_on(dev, label_lut)
        if out is None:
            out = torch.empty(full_shape, dtype=torch.int16, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        check(lib.gts_project_labels(ptr(svs_d), X, Y, Z, ptr(inv_maps[0]), ptr(inv_maps[1]), ptr(inv_maps[2]),
                                     ptr(cls), n_nodes, ptr(lut), lut.numel(), ptr(out),
                                     full_shape[0], full_shape[1], full_shape[2], ptr(err), st), "gts_project_labels")
    out._gts_err = err      # checked lazily by callers that want the reference's exception
    return out


def check_projection(out):
    """Raise the reference's RuntimeError('unexpected label') if the projection
    kernel flagged an id/class outside its tables (one device->host read)."""
    err = getattr(out, "_gts_err", None)
    if err is not None and int(err.item()):
        raise RuntimeError("unexpected label")
    return out


def project_logits_to_img(node_logits, svs, background=DEFAULT_BACKGROUND_NODE_LOGITS):
    """save_voxel_logits gather (scripts/generate_gnn_predictions.py:55-62; the
    on-device twin at scripts/generate_joint_predictions.py:64-66): fp32
    [X,Y,Z,C], background voxels get ``background``."""
    lib = _lib.load()
    dev = node_logits.device if (torch.is_tensor(node_logits) and node_logits.is_cuda) else _device()
    logits = _as_dev(node_logits, torch.float32, dev)
    svs_d = _as_dev(svs, torch.int16, dev)
    n_nodes, n_cls = logits.shape
    bg = torch.as_tensor(np.asarray(background, dtype=np.float32).reshape(-1)).to(dev)
    if bg.numel() != n_cls:
        raise ValueError("background row width must equal the number of classes")
    out = torch.empty(tuple(svs_d.shape) + (n_cls,), dtype=torch.float32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.gts_project_logits(ptr(svs_d), svs_d.numel(), ptr(logits), logits.stride(0), n_nodes, n_cls, ptr(bg),
                                     ptr(out), ptr(err), stream_ptr()), "gts_project_logits")
    out._gts_err = err
    return out


def determine_tumor_crop(svs, node_classes_or_logits):
    """determine_tumor_crop (data_processing/image_processing.py:8-17) for the GNN's own predictions, without
    building the voxel prediction volume: the crop the reference computes from
    ``node_logits[svs].argmax(-1)`` (scripts/generate_joint_predictions.py:64-67; PredLogitDataset.get_crop,
    data_processing/data_loader.py:146-151) depends only on which planes of the supervoxel map hold a
    predicted-tumour voxel.  Returns the same ``np.ix_`` tuple of int64 index arrays."""
    lib = _lib.load()
    x = node_classes_or_logits
    dev = x.device if (torch.is_tensor(x) and x.is_cuda) else _device()
    svs_d = _as_dev(svs, torch.int16, dev)
    if svs_d.dim() != 3:
        raise ValueError("svs must be the 3-D supervoxel map")
    X, Y, Z = (int(v) for v in svs_d.shape)
    with torch.cuda.device(dev):
        st = stream_ptr()
        xt = x if torch.is_tensor(x) else torch.as_tensor(np.ascontiguousarray(x))
        if xt.dim() == 2:
            logits = xt.to(device=dev, dtype=torch.float32).contiguous()
            cls = torch.empty(logits.shape[0], dtype=torch.int32, device=dev)
            check(lib.gts_argmax_rows(ptr(logits), logits.stride(0), logits.shape[0], logits.shape[1], ptr(cls), st),
                  "gts_argmax_rows")
        else:
            cls = xt.to(device=dev, dtype=torch.int32).contiguous()
        occ = torch.empty(X + Y + Z, dtype=torch.int32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        check(lib.gts_tumor_plane_occupancy(ptr(svs_d), X, Y, Z, ptr(cls), cls.numel(), ptr(occ), ptr(occ[X:]),
                                            ptr(occ[X + Y:]), ptr(err), st), "gts_tumor_plane_occupancy")
    host = torch.cat([occ, err]).cpu().numpy()          # one device->host read: 3 short flag vectors + the error flag
    if host[-1]:
        raise IndexError(f"supervoxel id out of bounds for {cls.numel()} nodes")
    flags = []
    for a, b in ((0, X), (X, X + Y), (X + Y, X + Y + Z)):
        o = host[a:b].astype(bool)
        d = o.copy()                                      # one step of binary dilation along the axis, border 0
        d[1:] |= o[:-1]
        d[:-1] |= o[1:]
        flags.append(d)
    if not any(f.any() for f in flags):                   # nothing predicted tumorous: the whole (uncropped) volume
        flags = [np.ones(n, dtype=bool) for n in (X, Y, Z)]
    return np.ix_(*flags)


class VolumeDownloader:
    """Ring of (device volume, pinned host volume) pairs with a private copy stream.

    ``generate_gnn_predictions`` (scripts/generate_gnn_predictions.py:43-52,64-73) produces one (240,240,155) int16
    label volume per MRI and hands it to the host (NIfTI write).  17.9 MB per volume over PCIe costs as much as the
    whole eval forward of the graph, so copying on the compute stream halves the throughput of bulk inference; here the
    device->host copy of volume i runs on its own stream while graph i+1 computes.

        dl = VolumeDownloader(depth=3)
        for ...:
            slot, vol = dl.acquire()                       # compute stream waits until the slot's last copy is done
            project_labels_to_brats(..., out=vol, ...)
            dl.submit(slot)                                # D2H on the copy stream, after the kernels enqueued so far
            ...
            host = dl.wait(slot)                           # pinned host tensor, valid until the slot is acquired again
    """

    def __init__(self, depth=3, shape=BRATS_SHAPE, dtype=torch.int16, device=None):
        self.device = torch.device(device) if device is not None else _device()
        self.depth = max(1, int(depth))
        self.dev = [torch.empty(shape, dtype=dtype, device=self.device) for _ in range(self.depth)]
        self.host = [torch.empty(shape, dtype=dtype).pin_memory() for _ in range(self.depth)]
        self.stream = torch.cuda.Stream(device=self.device)
        self._done = [None] * self.depth
        self._next = 0

    def acquire(self):
        slot = self._next
        self._next = (self._next + 1) % self.depth
        if self._done[slot] is not None:
            torch.cuda.current_stream(self.device).wait_event(self._done[slot])
        return slot, self.dev[slot]

    def submit(self, slot):
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.host[slot].copy_(self.dev[slot], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        self._done[slot] = done

    def wait(self, slot):
        if self._done[slot] is not None:
            self._done[slot].synchronize()
        return self.host[slot]

    def drain(self):
        self.stream.synchronize()

