"""Dataset side of the hot path with the reference's entry points
(data_processing/data_loader.py:39-169): ``ImageGraphDataset`` (graph part),
``PredLogitDataset.get_crop`` and ``minibatch_graphs``.

``get_graph`` returns ``(graph, features, labels)`` like the reference, with a
host ``BatchedGraph`` in place of the DGLGraph: the node-link JSON is parsed
once per file into the arrays the device CSR build consumes and cached in
binary form (graph_io.load_graph_json) instead of being re-parsed into networkx
and DGL objects for every sample of every epoch.

NIfTI files (images, voxel labels, supervoxel maps) are outside the hot path
(nibabel is absent from this image): the accessors read ``.npy``
twins when they exist and raise otherwise.
"""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from . import graph_io
from ._lib import GtsError
from .graph import minibatch_graphs  # noqa: F401  (collate_fn, data_loader.py:165-169)
from .project import determine_tumor_crop


class ImageGraphDataset(torch.utils.data.Dataset):
    """data_processing/data_loader.py:39-120."""

    def __init__(self, dataset_root_dir, mri_start_string, read_image=True, read_graph=True, read_label=True):
        self.dataset_root_dir = dataset_root_dir
        self.all_ids = self.get_all_mris_in_dataset(dataset_root_dir, mri_start_string)
        self.read_image = read_image
        self.read_graph = read_graph
        self.read_label = read_label
        assert (self.read_graph or self.read_image)

    def get_all_mris_in_dataset(self, dataset_root_dir, mri_start_string):
        mri_folders = glob.glob(f"{dataset_root_dir}**/{mri_start_string}*/", recursive=True)
        mri_ids = [fp.split(os.sep)[-2] for fp in mri_folders]
        print(f"Found {len(mri_folders)} MRIs")
        return mri_ids

    def get_one(self, mri_id):
        if self.read_graph and not self.read_image:
            return (mri_id, *self.get_graph(mri_id))
        elif self.read_image and not self.read_graph:
            return (mri_id, *self.get_image(mri_id))
        elif self.read_image and self.read_graph:
            return (mri_id, *self.get_graph(mri_id), *self.get_image(mri_id))
        else:
            print("Invalid combination of flags")

    def _path(self, mri_id, suffix):
        return f"{self.dataset_root_dir}{os.sep}{mri_id}{os.sep}{mri_id}{suffix}"

    def get_graph(self, mri_id):
        """data_loader.py:67-83.  The 'norm' node field the reference computes there is never read by
        any network and is not produced."""
        G, features, labels = graph_io.load_graph_json(self._path(mri_id, "_nxgraph.json"))
        if self.read_label:
            return G, features, labels
        return G, features

    def _read_volume(self, mri_id, stem, dtype):
        fp = self._path(mri_id, stem + ".npy")
        if os.path.exists(fp):
            return np.load(fp).astype(dtype, copy=False)
        raise GtsError(f"{self._path(mri_id, stem + '.nii.gz')}: NIfTI I/O is outside the hot path of this package "
                       f"(nibabel is not available); provide {fp}")

    def get_voxel_labels(self, mri_id):
        return self._read_volume(mri_id, "_label", np.int16)

    def get_image(self, mri_id):
        img = self._read_volume(mri_id, "_input", np.float32)
        if self.read_label:
            return img, self.get_voxel_labels(mri_id)
        return (img,)

    def get_supervoxel_partitioning(self, mri_id):
        return self._read_volume(mri_id, "_supervoxels", np.int16)

    def get_crop(self, mri_id):
        return tuple(np.load(self._path(mri_id, "_crop.npy"), allow_pickle=True))

    def __iter__(self):
        for mri_id in self.all_ids:
            yield self.get_one(mri_id)

    def __getitem__(self, index):
        return self.get_one(self.all_ids[index])

    def __len__(self):
        return len(self.all_ids)


class PredLogitDataset:
    """data_processing/data_loader.py:138-165: cached tumour crops of saved GNN logits.  ``get_crop`` here takes
    the supervoxel map and node logits / classes and runs the plane-occupancy kernel; the reference computes the
    same crop from the saved voxel-logit volume."""

    def __init__(self, root_dir=None):
        self.root_dir = root_dir
        self.mri_crops = {}

    def get_crop(self, mri_id, svs=None, node_logits=None):
        if mri_id in self.mri_crops:
            return self.mri_crops[mri_id]
        if svs is None or node_logits is None:
            raise GtsError("PredLogitDataset.get_crop: pass the supervoxel map and the node logits on first use "
                           "(saved NIfTI logit volumes are outside the hot path)")
        crop_idxs = determine_tumor_crop(svs, node_logits)
        self.mri_crops[mri_id] = crop_idxs
        return crop_idxs


class DevicePrefetcher:
    """Stage batch i+1 on a side stream while batch i computes.

    The reference moves every batch with ``.to(device)`` inside the step (model/gnn_model.py:38-40, DataLoader with
    num_workers=0): host->device copies, and here the device CSR build, sit on the critical path of every step.
    This iterator wraps any iterable of batches (tuples / lists whose members are ``BatchedGraph``s, tensors or
    anything else, e.g. the ``(mri_ids, graph, features, labels)`` tuples of ``minibatch_graphs``): members are moved
    with ``.to(device)`` on a private CUDA stream ``depth`` batches ahead (pinned host tensors make the copies
    asynchronous), the consumer's stream waits on the staging event, and the allocator is told about the hand-over
    (``record_stream``), so the yielded device objects can be used like the result of a plain ``.to(device)``.
    """

    def __init__(self, batches, device, depth=1):
        if not torch.cuda.is_available():
            raise GtsError("DevicePrefetcher needs a CUDA device (no CPU path)")
        self.batches = batches
        self.device = torch.device(device)
        self.depth = max(1, int(depth))
        self.stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return len(self.batches)

    def _stage(self, batch):
        from .graph import BatchedGraph
        with torch.cuda.stream(self.stream):
            out = []
            for m in (batch if isinstance(batch, (tuple, list)) else (batch,)):
                if isinstance(m, BatchedGraph):
                    out.append(m.to(self.device))
                elif torch.is_tensor(m):
                    out.append(m.to(self.device, non_blocking=True))
                elif isinstance(m, (tuple, list)) and m and all(torch.is_tensor(t) for t in m):
                    out.append(type(m)(t.to(self.device, non_blocking=True) for t in m))
                else:
                    out.append(m)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return (tuple(out) if isinstance(batch, (tuple, list)) else out[0]), ev

    def _hand_over(self, staged, ev):
        from .graph import BatchedGraph
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for m in (staged if isinstance(staged, tuple) else (staged,)):
            if isinstance(m, BatchedGraph):
                m.record_stream(cur)
            elif torch.is_tensor(m) and m.is_cuda:
                m.record_stream(cur)
            elif isinstance(m, (tuple, list)):
                for t in m:
                    if torch.is_tensor(t) and t.is_cuda:
                        t.record_stream(cur)
        return staged

    def __iter__(self):
        """Batch i+depth is staged before batch i is handed to the consumer."""
        from collections import deque
        self._it = iter(self.batches)
        self._q = deque()
        self._explicit = False
        while True:
            if not self._explicit:
                self.stage_next(self.depth + 1 - len(self._q))
            if not self._q:
                self.stage_next(1)
                if not self._q:
                    return
            yield self._hand_over(*self._q.popleft())

    def stage_next(self, n=None):
        """Stage up to ``n`` more batches now (default: fill the queue to ``depth``).  A consumer that synchronises with
        the device every step (``loss.item()``) calls this right AFTER enqueueing the step's kernels and BEFORE the
        sync: the host-side staging work then overlaps the device's compute instead of delaying the step's first
        kernel; once called, the iterator stops staging ahead on its own."""
        if n is None:
            self._explicit = True
            n = self.depth - len(self._q)
        for _ in range(max(0, n)):
            try:
                b = next(self._it)
            except StopIteration:
                return
            self._q.append(self._stage(b))


class StagedBatch:
    """Static device buffers for ONE batch signature (node / edge counts per graph, tensor shapes) with copy-stream
    staging — the input side of ``trainer.GraphedStep`` as a stand-alone helper for inference loops
    (scripts/generate_gnn_predictions.py:43-52 moves every graph in-stream with ``.to(device)``).

        sb.load_async(host_graph, tensors, copy_stream)     # H2D on the copy stream, overlaps the previous batch
        g, dev_tensors = sb.take()                           # compute stream waits for the copies, device CSR build
        ... enqueue the consumer ...
        sb.release()                                         # the next load_async waits for this point

    Nothing is allocated after construction except the CSR arrays of ``take`` (freed and reused by the caching
    allocator on the consumer's own stream: no cross-stream hand-over of allocator blocks, which is what made
    ``DevicePrefetcher`` unstable in time)."""

    def __init__(self, host_graph, tensors, device):
        if not torch.cuda.is_available():
            raise GtsError("StagedBatch needs a CUDA device (no CPU path)")
        dev = torch.device(device)
        self.signature = self.signature_of(host_graph, tensors)
        self.src = torch.empty(host_graph._src.shape, dtype=host_graph._src.dtype, device=dev)
        self.dst = torch.empty(host_graph._dst.shape, dtype=host_graph._dst.dtype, device=dev)
        self.node_off = host_graph._node_off.to(dev)
        self.edge_off = host_graph._edge_off.to(dev)
        self._node_counts, self._edge_counts = list(host_graph._node_counts), list(host_graph._edge_counts)
        self.tensors = [torch.empty(tuple(t.shape), dtype=t.dtype, device=dev) for t in tensors]
        self._loaded = self._done = None

    @staticmethod
    def signature_of(host_graph, tensors):
        return (tuple(host_graph._node_counts), tuple(host_graph._edge_counts), tuple((tuple(t.shape), t.dtype) for t in tensors))

    def load_async(self, host_graph, tensors, stream):
        if self.signature_of(host_graph, tensors) != self.signature:
            raise GtsError("StagedBatch: batch signature differs from the one the buffers were sized for")
        if self._done is not None:
            stream.wait_event(self._done)
        with torch.cuda.stream(stream):
            self.src.copy_(host_graph._src, non_blocking=True)
            self.dst.copy_(host_graph._dst, non_blocking=True)
            for d, t in zip(self.tensors, tensors):
                d.copy_(t, non_blocking=True)
        self._loaded = torch.cuda.Event()
        self._loaded.record(stream)

    def take(self):
        from .graph import BatchedGraph
        if self._loaded is None:
            raise GtsError("StagedBatch.take before load_async")
        torch.cuda.current_stream().wait_event(self._loaded)
        self._loaded = None
        g = BatchedGraph.from_device_edges(self.src, self.dst, self.node_off, self.edge_off, self._node_counts,
                                           self._edge_counts)
        return g, self.tensors

    def release(self):
        self._done = torch.cuda.Event()
        self._done.record(torch.cuda.current_stream())
