"""Data-parallel gradient exchange over NVLink peer memory, host side of csrc/peer.cu.

The reference trains on one device (model/gnn_model.py:23,41-47).  Sharded by whole graphs, the only cross-rank step
is the sum of the flat gradient arena (5 MB for the 7x256 stack).  ``PeerExchange`` gives every rank one
IPC-exportable buffer in its own HBM, maps the other ranks' buffers (CUDA IPC, NVLink / NVSwitch peer access) and
drives the two launches of a step — ``gts_peer_publish`` and ``gts_peer_allreduce_adamw`` — which carry their own
cross-GPU flags and therefore sit INSIDE the captured CUDA graph of the training step: no NCCL launch, no graph
segmentation, the optimiser fused into the pass that reads the peers' gradients.

``try_create`` is collective and agrees across ranks: where peer mapping is unavailable (no peer path between the
devices, IPC blocked) or the start-up self-test does not reproduce the exact sums, every rank gets ``None`` and the
trainer keeps the bucketed NCCL all-reduce.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import GtsError, check, ptr, stream_ptr


def enabled() -> bool:
    """GTS_DP_PEER=0 keeps the NCCL all-reduce (A/B runs)."""
    return os.environ.get("GTS_DP_PEER", "1") != "0"


class PeerExchange:
    def __init__(self, n_floats: int, group=None):
        """Collective over ``group``.  Raises GtsError on THIS rank's failure only after every rank has taken part in
        the handle exchange — use ``try_create`` for the agreed outcome."""
        if not (dist.is_available() and dist.is_initialized()):
            raise GtsError("PeerExchange needs an initialised torch.distributed process group")
        lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.MAX_PEERS:
            raise GtsError(f"PeerExchange: at most {_lib.MAX_PEERS} ranks (one box)")
        self.n = (int(n_floats) + 3) // 4 * 4
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._local = C.c_void_p()
        self._opened = {}
        self.comm = _lib.PeerComm()
        self.comm.rank, self.comm.world, self.comm.n = self.rank, self.world, self.n
        err = None
        handle = C.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        rc = lib.gts_peer_alloc(lib.gts_peer_buffer_bytes(self.n), C.byref(self._local), handle)
        if rc != _lib.GTS_OK:
            err = "gts_peer_alloc: " + (lib.gts_last_error() or b"").decode()
        mine = (self.rank, handle.raw if err is None else b"", os.getpid())
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        if err is None and any(len(h) != _lib.PEER_HANDLE_BYTES for _, h, _ in everyone):
            err = "a peer could not allocate its exchange buffer"
        if err is None:
            for r, h, _pid in everyone:
                if r == self.rank:
                    self.comm.base[r] = self._local.value
                    continue
                p = C.c_void_p()
                rc = lib.gts_peer_open(h, C.byref(p))
                if rc != _lib.GTS_OK:
                    err = f"gts_peer_open(rank {r}): " + (lib.gts_last_error() or b"").decode()
                    break
                self._opened[r] = p
                self.comm.base[r] = p.value
        self.error = err

    # ---- collective construction with agreement ---------------------------------------------------------------
    @classmethod
    def try_create(cls, n_floats: int, group=None):
        """Returns a working exchange on EVERY rank, or None on every rank (the caller keeps NCCL)."""
        if not enabled() or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
            return None
        if dist.get_backend(group) != "nccl" or not torch.cuda.is_available():
            return None
        try:
            ex = cls(n_floats, group)
        except GtsError as e:                       # before the exchange (too many ranks): same outcome on every rank
            cls.last_failure = str(e)
            return None
        if not ex._agree(ex.error is None):
            cls.last_failure = ex.error or "a peer rank could not map the exchange buffers"
            ex.close(collective=False)
            return None
        ok = False
        try:
            ok = ex.self_test()
        except GtsError as e:
            ex.error = str(e)
        if not ex._agree(ok):
            cls.last_failure = ex.error or "peer self-test failed on some rank"
            ex.close(collective=False)
            return None
        return ex

    last_failure = None

    def _agree(self, ok: bool) -> bool:
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return bool(int(t.item()))

    # ---- the two launches of a step -----------------------------------------------------------------------
    def publish(self, src: torch.Tensor):
        if src.numel() != self.n or src.dtype != torch.float32 or not src.is_cuda or not src.is_contiguous():
            raise GtsError(f"PeerExchange.publish: expected a contiguous CUDA fp32 vector of {self.n} elements")
        check(_lib.load().gts_peer_publish(C.byref(self.comm), ptr(src), stream_ptr()), "gts_peer_publish")

    def allreduce(self, grads: torch.Tensor):
        """Sum of every rank's published vector into ``grads`` (rank order: identical bits on every rank)."""
        if grads.numel() != self.n or grads.dtype != torch.float32 or not grads.is_cuda or not grads.is_contiguous():
            raise GtsError(f"PeerExchange.allreduce: expected a contiguous CUDA fp32 vector of {self.n} elements")
        check(_lib.load().gts_peer_allreduce_adamw(C.byref(self.comm), ptr(grads), 0, None, None, None, None, -1, 0,
                                                   stream_ptr()), "gts_peer_allreduce_adamw")

    def allreduce_adamw(self, grads, n_params, params, exp_avg, exp_avg_sq, hyper, denom_index):
        """The same sum, and AdamW on sum / (summed element ``denom_index``) for the first n_params elements."""
        if grads.numel() != self.n or not grads.is_contiguous():
            raise GtsError(f"PeerExchange.allreduce_adamw: expected a contiguous gradient vector of {self.n} elements")
        check(_lib.load().gts_peer_allreduce_adamw(C.byref(self.comm), ptr(grads), int(n_params), ptr(params),
                                                   ptr(exp_avg), ptr(exp_avg_sq), ptr(hyper), int(denom_index), 1,
                                                   stream_ptr()), "gts_peer_allreduce_adamw")

    def status(self):
        """(completed epochs, error word) — synchronises the device."""
        torch.cuda.synchronize(self.device)
        e, err = C.c_uint32(), C.c_uint32()
        check(_lib.load().gts_peer_status(C.byref(self.comm), C.byref(e), C.byref(err)), "gts_peer_status")
        return int(e.value), int(err.value)

    def self_test(self) -> bool:
        """Two exchanges (both staging buffers) of integer-valued vectors whose sums are exact in fp32."""
        base = (torch.arange(self.n, device=self.device, dtype=torch.float32) % 251.0) + 1.0
        tri = self.world * (self.world + 1) // 2
        out = torch.empty(self.n, dtype=torch.float32, device=self.device)
        ok = True
        for k in (1, 3):
            self.publish(base * float(k * (self.rank + 1)))
            out.zero_()
            self.allreduce(out)
            ok = ok and bool(torch.equal(out, base * float(k * tri)))
        _, err = self.status()
        if err:
            self.error = "a peer's flag timed out during the self-test"
        return ok and err == 0

    def close(self, collective: bool = True):
        """Unmap the peers' buffers and free the local one.  collective=True places a barrier between the two (a
        buffer must not be freed while a peer still has it mapped)."""
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        for p in self._opened.values():
            lib.gts_peer_close(p)
        self._opened = {}
        if collective and dist.is_initialized():
            dist.barrier(group=self.group)
        if self._local:
            if collective:
                lib.gts_peer_free(self._local)
            # non-collective: the buffer is left to process teardown (a peer may still have it mapped)
            self._local = C.c_void_p()
