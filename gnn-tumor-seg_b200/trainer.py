"""One-call training step for GraphSage('pool') behind the reference's loop body
(model/gnn_model.py:36-47: to(device), forward, weighted CE, zero_grad, backward, AdamW step):

* ``ParamArena``   — every parameter of a module re-pointed into ONE flat fp32 buffer, every ``.grad`` a view of a
                     second one (+ two trailing floats for [sum w*nll, sum w]); layer-major, so the gradients of the
                     top layers are one contiguous bucket.
* ``FusedAdamW``   — a ``torch.optim.Optimizer`` (so ``ExponentialLR`` and ``param_groups`` work unchanged) whose
                     ``step()`` is ONE ``gts_adamw_step_dev`` launch over the arena; hyper-parameters and the step
                     count live on the device, so the launch can be replayed from a CUDA graph.
* ``SageTrainer``  — forward + CE + backward as ONE library call (``gts_sage_step``) on caller-owned workspace and
                     arenas (nothing allocated per step), the optimiser launch behind it; data-parallel: the
                     gradient arena is exchanged over NVLink peer memory and summed, normalised and applied by ONE
                     kernel (``peer.PeerExchange``: gts_peer_publish + gts_peer_allreduce_adamw, both inside the
                     captured step); where the ranks cannot map each other's memory the backward is split in two
                     ranges and the finished range's bucket is all-reduced (NCCL, async) while the other range
                     runs, the loss denominator folded into the AdamW launch;
                     ``GraphedStep`` captures the whole step (device CSR build included) into a CUDA graph for
                     fixed-shape batches.

There is no CPU path: everything here requires CUDA tensors and libgts.so.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import GtsError, check, ptr, stream_ptr


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class ParamArena:
    """Flat parameter / gradient buffers.  ``groups``: list of lists of parameters; each group (a layer) is
    contiguous, every slice starts on a 16-byte boundary.  After construction ``p.data`` and ``p.grad`` of every
    parameter are views into ``self.params`` / ``self.grads``; ``self.extra`` = the two floats behind the gradients."""

    def __init__(self, groups, n_extra: int = 2):
        flat = [p for g in groups for p in g]
        if not flat:
            raise ValueError("ParamArena needs at least one parameter")
        dev = flat[0].device
        if dev.type != "cuda":
            raise GtsError("ParamArena: parameters must live on a CUDA device (no CPU path)")
        self.offsets, self.group_off = {}, []
        off = 0
        for g in groups:
            self.group_off.append(off)
            for p in g:
                if p.dtype != torch.float32:
                    raise GtsError("ParamArena: fp32 parameters only")
                self.offsets[id(p)] = off
                off += _pad4(p.numel())
        self.group_off.append(off)
        self.total = off
        self.param_list = flat
        self.params = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(off + _pad4(n_extra), dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p in flat:
                o, n = self.offsets[id(p)], p.numel()
                self.params[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.params[o:o + n].view_as(p)
                p.grad = self.grads[o:o + n].view_as(p)
        self.extra = self.grads[off:off + n_extra]

    def grad_view(self, p):
        o = self.offsets[id(p)]
        return self.grads[o:o + p.numel()].view_as(p)

    def grads_in_place(self) -> bool:
        """True when every parameter's .grad still is its arena view (the fused step writes there directly)."""
        base = self.grads.data_ptr()
        for p in self.param_list:
            g = p.grad
            if g is None or g.data_ptr() != base + 4 * self.offsets[id(p)] or not g.is_contiguous():
                return False
        return True


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction; model/gnn_model.py:28) as one
    ``gts_adamw_step_dev`` launch over a ParamArena.  ``groups`` (optional) = the layer grouping of the arena."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, groups=None):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise GtsError("FusedAdamW: one parameter group (the reference uses one, model/gnn_model.py:28)")
        plist = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        self.arena = ParamArena(groups if groups is not None else [plist])
        dev = self.arena.params.device
        self.exp_avg = torch.zeros_like(self.arena.params)
        self.exp_avg_sq = torch.zeros_like(self.arena.params)
        b1, b2 = betas
        # device-side hyper-parameter block {lr, beta1, beta2, eps, weight_decay, step, 0, 0}
        self.hyper = torch.tensor([lr, b1, b2, eps, weight_decay, 0.0, 0.0, 0.0], dtype=torch.float32, device=dev)
        self._dev_lr = float(lr)

    def _sync_lr(self):
        lr = float(self.param_groups[0]["lr"])
        if lr != self._dev_lr:                      # ExponentialLR rewrote it (once per epoch): one fill kernel, no sync
            self.hyper[0:1].fill_(lr)
            self._dev_lr = lr

    @torch.no_grad()
    def step(self, closure=None, grad_denom=None):
        """grad_denom: optional 1-element device tensor; gradients are divided by it inside the launch (the global
        loss denominator of data-parallel training)."""
        loss = closure() if closure is not None else None
        a = self.arena
        if not a.grads_in_place():                  # generic autograd produced its own tensors: pack them once
            for p in a.param_list:
                if p.grad is not None:
                    a.grad_view(p).copy_(p.grad)
                else:
                    a.grad_view(p).zero_()
                p.grad = a.grad_view(p)
        self._sync_lr()
        lib = _lib.load()
        check(lib.gts_adamw_step_dev(ptr(a.params), ptr(a.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq), a.total,
                                     ptr(self.hyper), 1.0, ptr(grad_denom), stream_ptr()), "gts_adamw_step_dev")
        ops._count(2)
        return loss

    def zero_grad(self, set_to_none: bool = True):
        """Gradients stay views of the arena (the fused backward overwrites every element); autograd users get the
        usual set-to-None behaviour so that the next backward assigns fresh tensors."""
        if set_to_none:
            for p in self.arena.param_list:
                p.grad = None
        else:
            self.arena.grads.zero_()
            for p in self.arena.param_list:
                p.grad = self.arena.grad_view(p)


def sage_layer_groups(net):
    """[fc_pool.weight, fc_pool.bias, fc_self.weight, fc_self.bias, fc_neigh.weight, fc_neigh.bias] per layer."""
    groups = []
    for l in net.layers:
        g = [l.fc_pool.weight, l.fc_pool.bias, l.fc_self.weight]
        if l.fc_self.bias is not None:
            g.append(l.fc_self.bias)
        g.append(l.fc_neigh.weight)
        if l.fc_neigh.bias is not None:
            g.append(l.fc_neigh.bias)
        groups.append(g)
    return groups


def plan_buckets(group_off, n_layers: int, n_buckets: int = 2):
    """Split of the backward for the data-parallel overlap: returns [(layer_hi, layer_lo, grad_lo, grad_hi), ...] in
    execution order (top layers first).  The last bucket's gradient range is extended by the caller to cover the two
    loss sums behind the arena.  Pure host logic (tests/test_trainer_host.py)."""
    n_buckets = max(1, min(n_buckets, n_layers))
    cuts = [round(n_layers * i / n_buckets) for i in range(n_buckets + 1)]      # 0 = bottom layer
    out = []
    for b in range(n_buckets, 0, -1):
        lo, hi = cuts[b - 1], cuts[b]
        out.append((hi, lo, group_off[lo], group_off[hi]))
    return out


class SageTrainer:
    """Fused training step for ``networks.GraphSage(..., 'pool', dropout=0)``.

    ``step(graph, feats, labels)`` = forward + weighted-mean CE + backward + AdamW, returns the loss as a 0-d device
    tensor (no host sync).  With a process group of more than one rank the step is data parallel over whole graphs
    (SURVEY.md §8e): un-normalised gradients, bucketed all-reduce overlapped with the rest of the backward, global
    denominator applied inside the optimiser launch — R ranks x B graphs equal one device on the union batch.
    ``data_parallel=False`` keeps a trainer local even when torch.distributed is initialised.
    """

    def __init__(self, net, class_weights, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8,
                 process_group=None, n_buckets=2, optimizer=None, data_parallel=None, peer=None):
        from .networks import GraphSage
        if not isinstance(net, GraphSage) or not all(l._aggre_type == "pool" for l in net.layers):
            raise GtsError("SageTrainer: a GraphSage('pool') network is required")
        if any(l.feat_drop.p != 0 for l in net.layers):
            raise GtsError("SageTrainer: feature dropout is not part of the fused step (use the autograd path)")
        if any(l.norm is not None for l in net.layers) or net.layers[-1].activation is not None:
            raise GtsError("SageTrainer: unsupported layer configuration")
        self.net = net
        self.class_weights = class_weights.detach().float().contiguous()
        ops.require_cuda(self.class_weights, *net.parameters())
        self.groups = sage_layer_groups(net)
        self.optimizer = optimizer or FusedAdamW(net.parameters(), lr=lr, betas=betas, eps=eps,
                                                 weight_decay=weight_decay, groups=self.groups)
        self.arena = self.optimizer.arena
        self.pg = process_group
        self.world_size = (dist.get_world_size(process_group)
                           if data_parallel is not False and dist.is_available() and dist.is_initialized() else 1)
        self.L = len(net.layers)
        # Data parallel: the gradient exchange over NVLink peer memory fused with AdamW (peer.PeerExchange, both
        # launches inside the captured step) where the ranks can map each other's buffers; otherwise — or with
        # peer=False / GTS_DP_PEER=0 — the bucketed NCCL all-reduce beside the backward.  Collective decision.
        self.peer = None
        if self.world_size > 1 and peer is not False and isinstance(self.optimizer, FusedAdamW):
            from .peer import PeerExchange
            self.peer = PeerExchange.try_create(self.arena.grads.numel(), process_group)
        use_buckets = self.world_size > 1 and self.peer is None
        self.buckets = plan_buckets(self.arena.group_off, self.L, n_buckets if use_buckets else 1)
        self._ws = None
        self._logits = None
        self._bias_scratch = None
        self._layers = (_lib.SageLayer * self.L)()
        self._grads = (_lib.SageLayerGrads * self.L)()
        self._fill_layer_structs()
        self.logits = None          # logits of the last step (view of an internal buffer)

    # ---- C structs over the arenas -----------------------------------------------------------------------
    def _fill_layer_structs(self):
        a = self.arena
        dev = a.params.device
        for i, l in enumerate(self.net.layers):
            ly, g = self._layers[i], self._grads[i]
            ly.din, ly.dout, ly.relu = l._in_feats, l._out_feats, int(l.activation is not None)
            ly.Wp, ly.bp, ly.Ws, ly.Wn = ptr(l.fc_pool.weight), ptr(l.fc_pool.bias), ptr(l.fc_self.weight), ptr(l.fc_neigh.weight)
            ly.b, ly.b2 = ptr(l.fc_self.bias), ptr(l.fc_neigh.bias)
            g.dWp, g.dbp = ptr(a.grad_view(l.fc_pool.weight)), ptr(a.grad_view(l.fc_pool.bias))
            g.dWs, g.dWn = ptr(a.grad_view(l.fc_self.weight)), ptr(a.grad_view(l.fc_neigh.weight))
            if l.fc_self.bias is not None:
                g.db = ptr(a.grad_view(l.fc_self.bias))
                g.db2 = ptr(a.grad_view(l.fc_neigh.bias)) if l.fc_neigh.bias is not None else None
            elif l.fc_neigh.bias is not None:
                g.db, g.db2 = ptr(a.grad_view(l.fc_neigh.bias)), None
            else:                                   # bias=False modules: the kernel still emits the column sums
                if self._bias_scratch is None:
                    self._bias_scratch = torch.empty(max(x._out_feats for x in self.net.layers), dtype=torch.float32, device=dev)
                g.db, g.db2 = ptr(self._bias_scratch), None

    def _buffers(self, n_nodes: int, mode: int):
        lib = _lib.load()
        need = lib.gts_sage_workspace_bytes(self._layers, self.L, n_nodes, 1, mode)
        if need == 0:
            raise GtsError("gts_sage_workspace_bytes: inconsistent layer dimensions")
        dev = self.arena.params.device
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=dev)
        n_cls = self.net.layers[-1]._out_feats
        if self._logits is None or self._logits.shape[0] < n_nodes:
            self._logits = torch.empty((n_nodes, n_cls), dtype=torch.float32, device=dev)
        return self._ws, self._logits[:n_nodes]

    def _step_args(self, graph, feats, labels, normalize: bool, bwd_layer_lo: int):
        ops.require_cuda(feats, labels)
        if not self.arena.grads_in_place():
            for p in self.arena.param_list:         # someone re-assigned .grad (autograd use in between): restore views
                p.grad = self.arena.grad_view(p)
        feats = ops._row_major_2d(feats)
        if labels.dtype != torch.int64 or not labels.is_contiguous():
            labels = labels.long().contiguous()
        n = feats.shape[0]
        mode = ops._gemm_mode
        ws, logits = self._buffers(n, mode)
        indptr, indices = graph.csr
        a = _lib.SageStepArgs()
        a.layers, a.grads, a.n_layers, a.n_nodes = self._layers, self._grads, self.L, n
        a.indptr, a.indices = ptr(indptr), ptr(indices)
        if ops.deterministic_backward():
            cptr, cidx, _ = graph.csc
            a.csc_indptr, a.csc_indices = ptr(cptr), ptr(cidx)
        a.feats, a.ldf = ptr(feats), ops._ld(feats)
        a.labels, a.class_w = ptr(labels), ptr(self.class_weights)
        a.sums = ptr(self.arena.extra)
        a.logits, a.ldl = ptr(logits), logits.shape[1]
        a.workspace, a.workspace_bytes = ptr(ws), ws.numel()
        a.mode, a.normalize, a.bwd_layer_lo = mode, int(normalize), bwd_layer_lo
        self.logits = logits
        self._keep = (feats, labels, graph)         # alive until the next step (the launches are asynchronous)
        return a

    # ---- the step ----------------------------------------------------------------------------------------
    def forward_backward(self, graph, feats, labels):
        """Forward + CE + backward (+ gradient exchange): gradients land in the arena; returns the loss (0-d device tensor).
        Single device: gradients of the weighted mean.  Data parallel: un-normalised sums, divided by the GLOBAL
        denominator inside the optimiser launch (``optimizer.step(grad_denom=trainer.denominator)``)."""
        lib = _lib.load()
        L = self.L
        if self.world_size == 1:
            a = self._step_args(graph, feats, labels, True, 0)
            check(lib.gts_sage_step(C.byref(a), stream_ptr()), "gts_sage_step")
            ops._count(23 * L + 3)
            return self.arena.extra[0] / self.arena.extra[1]
        if self.peer is not None:
            # data parallel over peer memory: whole backward, publish the arena, sum every rank's copy in rank order
            a = self._step_args(graph, feats, labels, False, 0)
            check(lib.gts_sage_step(C.byref(a), stream_ptr()), "gts_sage_step")
            self.peer.publish(self.arena.grads)
            self.peer.allreduce(self.arena.grads)
            ops._count(23 * L + 5)
            return self.arena.extra[0] / self.arena.extra[1]
        # data parallel: top bucket's layers first, its all-reduce runs while the lower layers back-propagate
        hi0, lo0, _, _ = self.buckets[0]
        a = self._step_args(graph, feats, labels, False, lo0)
        check(lib.gts_sage_step(C.byref(a), stream_ptr()), "gts_sage_step")
        works = []
        total = self.arena.total
        for i, (hi, lo, g_lo, g_hi) in enumerate(self.buckets):
            if i > 0:
                check(lib.gts_sage_step_backward_rest(C.byref(a), hi, lo, stream_ptr()), "gts_sage_step_backward_rest")
            end = total + 2 if i == 0 else g_hi        # the loss sums ride in the first (top) bucket
            works.append(dist.all_reduce(self.arena.grads[g_lo:end], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        for w in works:
            w.wait()                                    # stream-level wait: the host does not block
        ops._count(23 * L + 3)
        return self.arena.extra[0] / self.arena.extra[1]

    # ---- the data-parallel step as phases (GraphedStep captures each into its own CUDA graph; the NCCL calls between
    #      them stay eager) ---------------------------------------------------------------------------------------
    def dp_phase_first(self, graph, feats, labels):
        """forward + CE + backward of the top bucket's layers."""
        hi0, lo0, _, _ = self.buckets[0]
        self._dp_args = self._step_args(graph, feats, labels, False, lo0)
        check(_lib.load().gts_sage_step(C.byref(self._dp_args), stream_ptr()), "gts_sage_step")
        ops._count(23 * self.L + 3)

    def dp_phase_rest(self, i):
        """backward of bucket i (i >= 1)."""
        hi, lo, _, _ = self.buckets[i]
        check(_lib.load().gts_sage_step_backward_rest(C.byref(self._dp_args), hi, lo, stream_ptr()), "gts_sage_step_backward_rest")

    def dp_all_reduce(self, i):
        """async NCCL all-reduce of bucket i's gradient range (bucket 0 carries the two loss sums)."""
        _, _, g_lo, g_hi = self.buckets[i]
        end = self.arena.total + 2 if i == 0 else g_hi
        return dist.all_reduce(self.arena.grads[g_lo:end], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def check_exchange(self):
        """Raise GtsError when a flag wait of the peer-memory exchange ran into its bound (a rank fell more than 10 s
        behind or died: that step summed a stale staging buffer).  Synchronises the device — call it where the host
        reads results anyway (end of an epoch, before a checkpoint).  No-op without a peer exchange."""
        if self.peer is None:
            return
        epochs, err = self.peer.status()
        if err:
            raise GtsError(f"peer-memory gradient exchange: a rank's flag timed out (after {epochs} completed exchanges); "
                           "the parameters of this rank are no longer in sync — restart from the last checkpoint "
                           "(GTS_DP_PEER=0 selects the NCCL all-reduce)")

    @property
    def denominator(self):
        """sum of the class weights of the (global) batch — 1-element device tensor."""
        return self.arena.extra[1:2]

    def step(self, graph, feats, labels):
        if self.world_size > 1 and self.peer is not None:
            # forward + CE + backward, then TWO launches: publish the arena, and one pass that reads every rank's
            # copy over NVLink, sums in rank order, divides by the global loss denominator and applies AdamW
            opt, ar = self.optimizer, self.arena
            a = self._step_args(graph, feats, labels, False, 0)
            check(_lib.load().gts_sage_step(C.byref(a), stream_ptr()), "gts_sage_step")
            opt._sync_lr()
            self.peer.publish(ar.grads)
            self.peer.allreduce_adamw(ar.grads, ar.total, ar.params, opt.exp_avg, opt.exp_avg_sq, opt.hyper, ar.total + 1)
            ops._count(23 * self.L + 5)
            return ar.extra[0] / ar.extra[1]
        loss = self.forward_backward(graph, feats, labels)
        self.optimizer.step(grad_denom=self.denominator if self.world_size > 1 else None)
        return loss


class GraphedStep:
    """A whole training step for ONE batch signature (graphs, nodes, edges) captured into a CUDA graph:
    static device input buffers <- (async H2D copies, outside the graph) <- pinned host batch; graph = batched edge
    list + CSR build (K6) + forward + CE + backward + AdamW + loss.  Replay costs one launch on the host side."""

    def __init__(self, trainer: SageTrainer, host_graph, feats, labels, capture: bool = True):
        """capture=False keeps the static input buffers and the copy-stream staging but runs the step eagerly on
        replay.  With a data-parallel trainer the step is captured as segments around the eager NCCL all-reduces."""
        from .graph import BatchedGraph
        self.trainer = trainer
        self._BatchedGraph = BatchedGraph
        self.graph = None
        self.segments = None
        dev = trainer.arena.params.device
        self.signature = self.signature_of(host_graph, feats)
        self.src = torch.empty_like(host_graph._src, device=dev)
        self.dst = torch.empty_like(host_graph._dst, device=dev)
        self.node_off = host_graph._node_off.to(dev)
        self.edge_off = host_graph._edge_off.to(dev)
        self.feats = torch.empty(tuple(feats.shape), dtype=torch.float32, device=dev)
        self.labels = torch.empty(tuple(labels.shape), dtype=torch.int64, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self._node_counts, self._edge_counts = list(host_graph._node_counts), list(host_graph._edge_counts)
        self._load(host_graph, feats, labels)
        if not capture:
            return
        # warm-up on a side stream (lazy initialisation: function attributes, tensor-map entry point), then capture
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        n0 = ops.launch_counter["n"]
        with torch.cuda.stream(s):
            self._body(BatchedGraph)
        self.launches_per_replay = ops.launch_counter["n"] - n0        # kernels one replay stands for
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        if trainer.world_size == 1 or trainer.peer is not None:
            # (data parallel over peer memory: the exchange kernels carry their own cross-GPU flags — one graph)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._body(BatchedGraph)
            return
        # Data parallel: the NCCL all-reduces stay OUTSIDE the graphs (capturing the async collectives hung the ranks on
        # 2 x B200, torch 2.11 / NCCL 2.28).  The step becomes captured segments with the eager collectives between
        # them: [CSR build + forward + CE + backward of the top bucket] | all-reduce | [backward of bucket i] | all-reduce
        # ... | [AdamW + loss]; every rank replays the same sequence.
        pool = torch.cuda.graph_pool_handle()
        self.segments = []
        g0 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g0, pool=pool):
            self.device_graph = BatchedGraph.from_device_edges(self.src, self.dst, self.node_off, self.edge_off,
                                                               self._node_counts, self._edge_counts)
            trainer.dp_phase_first(self.device_graph, self.feats, self.labels)
        self.segments.append(g0)
        for i in range(1, len(trainer.buckets)):
            gi = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gi, pool=pool):
                trainer.dp_phase_rest(i)
            self.segments.append(gi)
        self._dp_args = trainer._dp_args                  # the phases' argument block (pointers into static buffers)
        gl = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gl, pool=pool):
            trainer.optimizer.step(grad_denom=trainer.denominator)
            self.loss.copy_(trainer.arena.extra[0] / trainer.arena.extra[1])
        self.segments.append(gl)

    @staticmethod
    def signature_of(host_graph, feats):
        return (tuple(host_graph._node_counts), tuple(host_graph._edge_counts), tuple(feats.shape))

    def _body(self, BatchedGraph):
        g = BatchedGraph.from_device_edges(self.src, self.dst, self.node_off, self.edge_off, self._node_counts,
                                           self._edge_counts)
        self.device_graph = g
        self.loss.copy_(self.trainer.step(g, self.feats, self.labels))

    def _load(self, host_graph, feats, labels):
        self.src.copy_(host_graph._src, non_blocking=True)
        self.dst.copy_(host_graph._dst, non_blocking=True)
        self.feats.copy_(feats, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)

    def load_async(self, host_graph, feats, labels, stream):
        """Stage the next batch with this signature on ``stream`` (a copy stream) while other work runs on the compute
        stream: the copies wait for the previous replay of THIS object (the last reader of its static buffers)."""
        if self.signature_of(host_graph, feats) != self.signature:
            raise GtsError("GraphedStep: batch signature differs from the captured one")
        if getattr(self, "_done", None) is not None:
            stream.wait_event(self._done)
        with torch.cuda.stream(stream):
            self._load(host_graph, feats, labels)
        self._loaded = torch.cuda.Event()
        self._loaded.record(stream)

    def replay(self):
        """Run the captured step on the current stream once the staged inputs have landed; returns the loss (0-d device
        tensor, overwritten by the next replay of this object)."""
        cur = torch.cuda.current_stream()
        if getattr(self, "_loaded", None) is not None:
            cur.wait_event(self._loaded)
            self._loaded = None
        self.trainer.optimizer._sync_lr()
        if self.graph is not None:
            self.graph.replay()
        elif self.segments is not None:
            tr = self.trainer
            tr._dp_args = self._dp_args
            works = []
            for i in range(len(tr.buckets)):
                self.segments[i].replay()
                works.append(tr.dp_all_reduce(i))        # NCCL stream; the next segment runs beside it
            for w in works:
                w.wait()
            self.segments[-1].replay()
        else:
            self._body(self._BatchedGraph)
        self._done = torch.cuda.Event()
        self._done.record(cur)
        return self.loss

    def __call__(self, host_graph, feats, labels):
        """One step on a batch with this signature; returns the loss (0-d device tensor, overwritten by the next
        replay)."""
        if self.signature_of(host_graph, feats) != self.signature:
            raise GtsError("GraphedStep: batch signature differs from the captured one")
        self._load(host_graph, feats, labels)
        return self.replay()
