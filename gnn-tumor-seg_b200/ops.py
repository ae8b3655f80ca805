"""Tensor-level wrappers over the C-ABI (include/gts.h) and the autograd
Functions built from them.

PyTorch is plumbing here: it owns device memory (outputs, workspaces), the
current stream, and the autograd tape; all arithmetic on the hot path runs in
libgts.so kernels.  Every wrapper requires CUDA tensors and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import ACT_MASK_BITS, ACT_MASK_POS, ACT_MASK_POS_SCATTER, ACT_NONE, ACT_RELU, GEMM_MODES, GemmNtArgs, check, ptr, require_cuda, stream_ptr

# ---------------------------------------------------------------------------
# arithmetic mode of the dense contractions (stated per run in bench/tests)
# ---------------------------------------------------------------------------
_gemm_mode = GEMM_MODES[os.environ.get("GTS_GEMM_MODE", "tf32x3")]


def set_gemm_mode(mode: str) -> None:
    """'fp32' (SIMT FFMA), 'tf32' (tcgen05 1xTF32) or 'tf32x3' (tcgen05 3xTF32)."""
    global _gemm_mode
    _gemm_mode = GEMM_MODES[mode]


def get_gemm_mode() -> str:
    return {v: k for k, v in GEMM_MODES.items()}[_gemm_mode]


_deterministic = os.environ.get("GTS_DETERMINISTIC", "0") == "1"


def set_deterministic_backward(flag: bool) -> None:
    """True: neighbour-max backward uses the transposed gather over the out-edge
    CSC (no float atomics, bit-reproducible); False: atomic scatter (faster)."""
    global _deterministic
    _deterministic = bool(flag)


def deterministic_backward() -> bool:
    return _deterministic


_stack_path = os.environ.get("GTS_STACK_PATH", "1") == "1"


def set_stack_path(flag: bool) -> None:
    """True (default): GraphSage.forward runs the whole layer stack through
    gts_sage_forward/backward; False: one autograd Function per layer."""
    global _stack_path
    _stack_path = bool(flag)


def use_stack_path() -> bool:
    return _stack_path


# counts kernels launched through this module (bench.py's gpu_launches claim)
launch_counter = {"n": 0}


def _count(n: int = 1) -> None:
    launch_counter["n"] += n


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _row_major_2d(t: torch.Tensor) -> torch.Tensor:
    """fp32 2-D with unit column stride (row stride arbitrary >= cols)."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D tensor, got shape {tuple(t.shape)}")
    if t.shape[1] > 1 and t.stride(1) != 1:
        t = t.contiguous()
    if t.shape[0] > 1 and t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------------------
# K6 graph build
# ---------------------------------------------------------------------------
def batch_edges(src_local, dst_local, edge_off, node_off):
    require_cuda(src_local, dst_local, edge_off, node_off)
    lib = _lib.load()
    E = src_local.numel()
    src_g = torch.empty(E, dtype=torch.int32, device=src_local.device)
    dst_g = torch.empty(E, dtype=torch.int32, device=src_local.device)
    check(lib.gts_batch_edges(ptr(src_local), ptr(dst_local), E, ptr(edge_off), ptr(node_off),
                              edge_off.numel() - 1, ptr(src_g), ptr(dst_g), stream_ptr()), "gts_batch_edges")
    _count()
    return src_g, dst_g


def csr_build(row, col, n_nodes, want_eid=False):
    """Canonical CSR keyed by `row` (int32 CUDA tensors)."""
    require_cuda(row, col)
    lib = _lib.load()
    E = row.numel()
    dev = row.device
    indptr = torch.empty(n_nodes + 1, dtype=torch.int32, device=dev)
    indices = torch.empty(E, dtype=torch.int32, device=dev)
    eid = torch.empty(E, dtype=torch.int32, device=dev) if want_eid else None
    ws = _workspace(lib.gts_csr_build_workspace_bytes(E, n_nodes), dev)
    check(lib.gts_csr_build(ptr(row), ptr(col), E, n_nodes, ptr(indptr), ptr(indices), ptr(eid),
                            ptr(ws), ws.numel(), stream_ptr()), "gts_csr_build")
    _count(7)
    return indptr, indices, eid


def edge_perm_compose(eid_csr, eid_csc):
    require_cuda(eid_csr, eid_csc)
    lib = _lib.load()
    E = eid_csr.numel()
    scratch = torch.empty(E, dtype=torch.int32, device=eid_csr.device)
    out = torch.empty(E, dtype=torch.int32, device=eid_csr.device)
    check(lib.gts_edge_perm_compose(ptr(eid_csr), ptr(eid_csc), E, ptr(scratch), ptr(out), stream_ptr()),
          "gts_edge_perm_compose")
    _count(2)
    return out


# ---------------------------------------------------------------------------
# dense contractions
# ---------------------------------------------------------------------------
def gemm_nt(A1, B1, A2=None, B2=None, bias=None, act=ACT_NONE, aux=None, mode=None, out=None, bias2=None,
            relu_bits_out=None, aux_bits=None, zero_fill=None):
    """act(A1 @ B1.T + A2 @ B2.T + bias); B* in nn.Linear layout [out,in].
    zero_fill: an unrelated contiguous tensor the call also clears (gts_gemm_nt_args.zero_fill: spare warps of the
    256-wide kernel, a memset otherwise).
    relu_bits_out (int32 [M, N/32], with ACT_RELU): receives the bit matrix of (C > 0); aux_bits (with ACT_MASK_BITS): the
    mask as such a bit matrix (layout: include/gts.h GTS_ACT_MASK_BITS) — 256-wide tensor-core path only."""
    require_cuda(A1, B1, A2, B2, bias, aux)
    lib = _lib.load()
    A1 = _row_major_2d(A1)
    B1 = _row_major_2d(B1)
    M, K1 = A1.shape
    N = B1.shape[0]
    assert B1.shape[1] == K1, (A1.shape, B1.shape)
    a = GemmNtArgs()
    a.A1, a.lda1, a.K1 = ptr(A1), _ld(A1), K1
    a.B1, a.ldb1 = ptr(B1), _ld(B1)
    if A2 is not None:
        A2 = _row_major_2d(A2)
        B2 = _row_major_2d(B2)
        assert A2.shape[0] == M and B2.shape[0] == N and A2.shape[1] == B2.shape[1]
        a.A2, a.lda2, a.K2 = ptr(A2), _ld(A2), A2.shape[1]
        a.B2, a.ldb2 = ptr(B2), _ld(B2)
    else:
        a.A2, a.lda2, a.K2, a.B2, a.ldb2 = None, 0, 0, None, 0
    if bias is not None:
        bias = _f32c(bias)
        assert bias.numel() == N
    a.bias = ptr(bias)
    if bias2 is not None:
        bias2 = _f32c(bias2)
        assert bias2.numel() == N
    a.bias2 = ptr(bias2)
    if aux is not None:
        aux = _row_major_2d(aux)
        assert tuple(aux.shape) == (M, N)
        a.aux, a.ldaux = ptr(aux), _ld(aux)
    else:
        a.aux, a.ldaux = None, 0
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A1.device)
    a.C, a.ldc = ptr(out), _ld(out)
    if relu_bits_out is not None:
        require_cuda(relu_bits_out)
        assert relu_bits_out.dtype == torch.int32 and tuple(relu_bits_out.shape) == (M, N // 32) and relu_bits_out.is_contiguous()
        a.relu_bits_out, a.ld_bits_out = ptr(relu_bits_out), N // 32
    if aux_bits is not None:
        require_cuda(aux_bits)
        assert aux_bits.dtype == torch.int32 and tuple(aux_bits.shape) == (M, N // 32) and aux_bits.is_contiguous()
        a.aux_bits, a.ld_aux_bits = ptr(aux_bits), N // 32
    if zero_fill is not None:
        require_cuda(zero_fill)
        assert zero_fill.is_contiguous()
        a.zero_fill, a.zero_fill_bytes = ptr(zero_fill), zero_fill.numel() * zero_fill.element_size()
    a.M, a.N, a.act = M, N, act
    a.mode = _gemm_mode if mode is None else (GEMM_MODES[mode] if isinstance(mode, str) else mode)
    check(lib.gts_gemm_nt(C.byref(a), stream_ptr()), "gts_gemm_nt")
    _count()
    return out


def gemm_nt_scatter(A, B, aux, idx, n_out_rows, mode=None):
    """Backward of the neighbour max fused into the GEMM that feeds it (GTS_ACT_MASK_POS_SCATTER):
    ``v = (A @ B.T) * (aux > 0)`` is never stored; ``out[idx[m,n], n] += v[m,n]`` for idx >= 0.
    Equals ``segmax_bwd(gemm_nt(A, B, act=ACT_MASK_POS, aux=aux), idx, n_out_rows)`` up to the order of the
    fp32 additions.  Returns the zero-initialised-then-accumulated [n_out_rows, N] tensor."""
    require_cuda(A, B, aux, idx)
    lib = _lib.load()
    A = _row_major_2d(A)
    B = _row_major_2d(B)
    aux = _row_major_2d(aux)
    M, K = A.shape
    N = B.shape[0]
    assert B.shape[1] == K and tuple(aux.shape) == (M, N) and tuple(idx.shape) == (M, N) and idx.dtype == torch.int32
    idx = idx.contiguous()
    out = torch.zeros((n_out_rows, N), dtype=torch.float32, device=A.device)
    a = GemmNtArgs()
    a.A1, a.lda1, a.K1 = ptr(A), _ld(A), K
    a.B1, a.ldb1 = ptr(B), _ld(B)
    a.aux, a.ldaux = ptr(aux), _ld(aux)
    a.C, a.ldc = None, N
    a.M, a.N, a.act = M, N, ACT_MASK_POS_SCATTER
    a.mode = _gemm_mode if mode is None else (GEMM_MODES[mode] if isinstance(mode, str) else mode)
    a.scatter_idx, a.ld_idx = ptr(idx), N
    a.scatter_out, a.ld_out = ptr(out), N
    check(lib.gts_gemm_nt(C.byref(a), stream_ptr()), "gts_gemm_nt")
    _count(2)
    return out


def gemm_tn(A, B, mode=None):
    """A.T @ B for A [K,Mo], B [K,No] (weight gradients)."""
    require_cuda(A, B)
    lib = _lib.load()
    A = _row_major_2d(A)
    B = _row_major_2d(B)
    K, Mo = A.shape
    No = B.shape[1]
    assert B.shape[0] == K
    m = _gemm_mode if mode is None else (GEMM_MODES[mode] if isinstance(mode, str) else mode)
    out = torch.empty((Mo, No), dtype=torch.float32, device=A.device)
    ws = _workspace(lib.gts_gemm_tn_workspace_bytes(Mo, No, K, m), A.device)
    check(lib.gts_gemm_tn(ptr(A), _ld(A), ptr(B), _ld(B), ptr(out), No, Mo, No, K, m, ptr(ws), ws.numel(),
                          stream_ptr()), "gts_gemm_tn")
    _count(2)
    return out


def gemm_tn_colsum(A, B, mode=None):
    """(A.T @ B, A.sum(0)) — a weight gradient and the bias gradient sharing its A operand.
    In the tf32x3 mode the column sums are a by-product of the kernel that stages A through
    tensor memory; otherwise gts_gemm_tn + gts_colsum behind the same entry point."""
    require_cuda(A, B)
    lib = _lib.load()
    A = _row_major_2d(A)
    B = _row_major_2d(B)
    K, Mo = A.shape
    No = B.shape[1]
    assert B.shape[0] == K
    m = _gemm_mode if mode is None else (GEMM_MODES[mode] if isinstance(mode, str) else mode)
    out = torch.empty((Mo, No), dtype=torch.float32, device=A.device)
    cs = torch.empty(Mo, dtype=torch.float32, device=A.device)
    ws = _workspace(lib.gts_gemm_tn_colsum_workspace_bytes(Mo, No, K, m), A.device)
    check(lib.gts_gemm_tn_colsum(ptr(A), _ld(A), ptr(B), _ld(B), ptr(out), No, Mo, No, K, m, ptr(cs), ptr(ws),
                                 ws.numel(), stream_ptr()), "gts_gemm_tn_colsum")
    _count(3)
    return out, cs


def gemm_tn2_colsum(A, B1, B2, mode=None):
    """(A.T @ B1, A.T @ B2, A.sum(0)) in one pass over A (tf32x3: the CTA-pair kernel feeds both products from the
    same split A tile in tensor memory).  The two products are the halves of ONE [2, Mo, No] buffer."""
    require_cuda(A, B1, B2)
    lib = _lib.load()
    A = _row_major_2d(A)
    B1 = _row_major_2d(B1)
    B2 = _row_major_2d(B2)
    K, Mo = A.shape
    No = B1.shape[1]
    assert B1.shape == B2.shape and B1.shape[0] == K
    m = _gemm_mode if mode is None else (GEMM_MODES[mode] if isinstance(mode, str) else mode)
    out = torch.empty((2, Mo, No), dtype=torch.float32, device=A.device)
    cs = torch.empty(Mo, dtype=torch.float32, device=A.device)
    ws = _workspace(lib.gts_gemm_tn2_colsum_workspace_bytes(Mo, No, K, m), A.device)
    check(lib.gts_gemm_tn2_colsum(ptr(A), _ld(A), ptr(B1), _ld(B1), ptr(B2), _ld(B2), ptr(out[0]), ptr(out[1]), No,
                                  Mo, No, K, m, ptr(cs), ptr(ws), ws.numel(), stream_ptr()), "gts_gemm_tn2_colsum")
    _count(2)
    return out[0], out[1], cs


def colsum(A):
    require_cuda(A)
    lib = _lib.load()
    A = _row_major_2d(A)
    rows, cols = A.shape
    out = torch.empty(cols, dtype=torch.float32, device=A.device)
    ws = _workspace(lib.gts_colsum_workspace_bytes(rows, cols), A.device)
    check(lib.gts_colsum(ptr(A), _ld(A), rows, cols, ptr(out), ptr(ws), ws.numel(), stream_ptr()), "gts_colsum")
    _count(2)
    return out


def transpose(W):
    require_cuda(W)
    lib = _lib.load()
    W = _row_major_2d(W)
    r, c = W.shape
    out = torch.empty((c, r), dtype=torch.float32, device=W.device)
    check(lib.gts_transpose(ptr(W), _ld(W), r, c, ptr(out), r, stream_ptr()), "gts_transpose")
    _count()
    return out


def mask_pos(grad, ref):
    require_cuda(grad, ref)
    lib = _lib.load()
    grad = _f32c(grad)
    ref = _f32c(ref)
    out = torch.empty_like(grad)
    check(lib.gts_mask_pos(ptr(grad), ptr(ref), grad.numel(), ptr(out), stream_ptr()), "gts_mask_pos")
    _count()
    return out


# ---------------------------------------------------------------------------
# K2 aggregation
# ---------------------------------------------------------------------------
def segmax_fwd(P, indptr, indices, want_argmax=True):
    require_cuda(P, indptr, indices)
    lib = _lib.load()
    P = _row_major_2d(P)
    N = indptr.numel() - 1
    D = P.shape[1]
    neigh = torch.empty((N, D), dtype=torch.float32, device=P.device)
    arg = torch.empty((N, D), dtype=torch.int32, device=P.device) if want_argmax else None
    check(lib.gts_segmax_fwd(ptr(P), _ld(P), ptr(indptr), ptr(indices), N, D, ptr(neigh), D,
                             ptr(arg), D, stream_ptr()), "gts_segmax_fwd")
    _count()
    return neigh, arg


def segmax_bwd(dNeigh, arg, n_src_rows, csc=None):
    """Scatter through the saved arg-max.  ``csc=(indptr, indices)`` selects the
    deterministic transposed-gather form."""
    require_cuda(dNeigh, arg)
    lib = _lib.load()
    dNeigh = _row_major_2d(dNeigh)
    N, D = dNeigh.shape
    dP = torch.empty((n_src_rows, D), dtype=torch.float32, device=dNeigh.device)
    if csc is None:
        check(lib.gts_segmax_bwd(ptr(dNeigh), _ld(dNeigh), ptr(arg), D, N, D, ptr(dP), D, n_src_rows,
                                 stream_ptr()), "gts_segmax_bwd")
        _count(2)
    else:
        assert n_src_rows == N
        check(lib.gts_segmax_bwd_det(ptr(dNeigh), _ld(dNeigh), ptr(arg), D, ptr(csc[0]), ptr(csc[1]), N, D,
                                     ptr(dP), D, stream_ptr()), "gts_segmax_bwd_det")
        _count()
    return dP


SEGSUM_MODES = {"sum": 0, "mean": 1, "gcn": 2}


def segsum_fwd(P, indptr, indices, mode):
    require_cuda(P, indptr, indices)
    lib = _lib.load()
    P = _row_major_2d(P)
    N = indptr.numel() - 1
    D = P.shape[1]
    out = torch.empty((N, D), dtype=torch.float32, device=P.device)
    check(lib.gts_segsum_fwd(ptr(P), _ld(P), ptr(indptr), ptr(indices), N, D, SEGSUM_MODES[mode], ptr(out), D,
                             stream_ptr()), "gts_segsum_fwd")
    _count()
    return out


def segsum_bwd(dOut, csc_indptr, csc_indices, in_indptr, mode):
    require_cuda(dOut, csc_indptr, csc_indices, in_indptr)
    lib = _lib.load()
    dOut = _row_major_2d(dOut)
    N, D = dOut.shape
    dP = torch.empty((N, D), dtype=torch.float32, device=dOut.device)
    check(lib.gts_segsum_bwd(ptr(dOut), _ld(dOut), ptr(csc_indptr), ptr(csc_indices), ptr(in_indptr), N, D,
                             SEGSUM_MODES[mode], ptr(dP), D, stream_ptr()), "gts_segsum_bwd")
    _count()
    return dP


# ---------------------------------------------------------------------------
# K8 loss
# ---------------------------------------------------------------------------
def ce_weighted(logits, labels, class_w, want_grad=True):
    """Returns (sums[2] = [sum w*nll, sum w], dlogits_unnormalised | None)."""
    require_cuda(logits, labels, class_w)
    lib = _lib.load()
    logits = _row_major_2d(logits)
    labels = labels.contiguous()
    if labels.dtype != torch.int64:
        labels = labels.long()
    class_w = _f32c(class_w)
    N, Cn = logits.shape
    sums = torch.zeros(2, dtype=torch.float32, device=logits.device)
    dl = torch.empty((N, Cn), dtype=torch.float32, device=logits.device) if want_grad else None
    check(lib.gts_ce_weighted(ptr(logits), _ld(logits), ptr(labels), ptr(class_w), N, Cn, ptr(sums), ptr(dl), Cn,
                              stream_ptr()), "gts_ce_weighted")
    _count()
    return sums, dl


def scale_by_inv_(x, alpha, denom):
    require_cuda(x, denom)
    lib = _lib.load()
    assert x.is_contiguous() and x.dtype == torch.float32
    check(lib.gts_scale_by_inv(ptr(x), x.numel(), float(alpha), ptr(denom), stream_ptr()), "gts_scale_by_inv")
    _count()
    return x


class _WeightedCE(torch.autograd.Function):
    """Weighted-mean CE as one fused kernel (+ one scale): drop-in for
    torch.nn.CrossEntropyLoss(weight=w)(logits, labels) (model/gnn_model.py:30,42)."""

    @staticmethod
    def forward(ctx, logits, labels, class_w):
        sums, dl = ce_weighted(logits, labels, class_w, want_grad=True)
        scale_by_inv_(dl, 1.0, sums[1:2])
        ctx.save_for_backward(dl)
        return sums[0] / sums[1]

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None


def weighted_cross_entropy(logits, labels, class_w):
    return _WeightedCE.apply(logits, labels, class_w)


# ---------------------------------------------------------------------------
# SAGEConv('pool') layer: forward + backward as kernel sequences
# ---------------------------------------------------------------------------
class SagePoolLayerFn(torch.autograd.Function):
    """One DGL SAGEConv(in,out,'pool') layer.

    forward : P = relu(h Wp^T + bp) -> (neigh, arg) = segmax(P) ->
              out = act(h Ws^T + neigh Wn^T + b)          [3 kernels]
    backward: see _backward below.

    ``input_is_relu``: h is the ReLU output of the previous layer, so the
    gradient returned for h may be pre-masked by (h > 0) inside the GEMM
    epilogue; ``grad_premasked``: the incoming gradient already carries this
    layer's own ReLU mask (the consumer pre-masked it).  Both are set only by
    GraphSage.forward for its strictly sequential stack.
    """

    @staticmethod
    def forward(ctx, h, Wp, bp, Ws, Wn, b, graph, relu_out, input_is_relu, grad_premasked, deterministic):
        require_cuda(h, Wp, bp, Ws, Wn, b)
        h = _row_major_2d(h)
        indptr, indices = graph.csr
        P = gemm_nt(h, Wp, bias=bp, act=ACT_RELU)
        need_grad = any(ctx.needs_input_grad[:6])
        neigh, arg = segmax_fwd(P, indptr, indices, want_argmax=need_grad)
        del P
        out = gemm_nt(h, Ws, neigh, Wn, bias=b, act=ACT_RELU if relu_out else ACT_NONE)
        if need_grad:
            ctx.save_for_backward(h, neigh, arg, out if relu_out else None, Wp, Ws, Wn)
            ctx.graph = graph
            ctx.flags = (relu_out, input_is_relu, grad_premasked, deterministic)
        return out

    @staticmethod
    def backward(ctx, dOut):
        h, neigh, arg, out, Wp, Ws, Wn = ctx.saved_tensors
        relu_out, input_is_relu, grad_premasked, deterministic = ctx.flags
        dOut = _row_major_2d(dOut)
        dZ = mask_pos(dOut, out) if (relu_out and not grad_premasked) else dOut
        dWs, dWn, db = gemm_tn2_colsum(dZ, h, neigh)        # one pass over dZ
        # dNeigh' = (dZ Wn) * (neigh > 0): ReLU mask of fc_pool folded here, since
        # neigh[v,k] = P[arg[v,k],k] (Appendix A.1)
        dNeigh = gemm_nt(dZ, transpose(Wn), act=ACT_MASK_POS, aux=neigh)
        csc = ctx.graph.csc[:2] if deterministic else None
        dP = segmax_bwd(dNeigh, arg, h.shape[0], csc=csc)     # (the fused gemm_nt_scatter form measured slower)
        del dNeigh
        dWp, dbp = gemm_tn_colsum(dP, h)
        dh = None
        if ctx.needs_input_grad[0]:
            dh = gemm_nt(dZ, transpose(Ws), dP, transpose(Wp),
                         act=ACT_MASK_POS if input_is_relu else ACT_NONE, aux=h if input_is_relu else None)
        return dh, dWp, dbp, dWs, dWn, db, None, None, None, None, None


def sage_pool_layer(h, Wp, bp, Ws, Wn, b, graph, relu_out, input_is_relu=False, grad_premasked=False,
                    deterministic=False):
    return SagePoolLayerFn.apply(h, Wp, bp, Ws, Wn, b, graph, relu_out, input_is_relu, grad_premasked, deterministic)


# the flat buffer behind the parameter gradients of the most recent SageStackFn.backward
# (data-parallel trainer all-reduces it in place: no per-parameter copies)
last_flat_grads = {"flat": None, "n_grad": 0}


class SageStackFn(torch.autograd.Function):
    """GraphSage.forward over ALL SAGEConv('pool') layers as one libgts call, and the
    whole backward as another (gts_sage_forward / gts_sage_backward): no Python between
    the ~20 kernels of a layer.  ``flat`` = (Wp, bp, Ws, bs, Wn, bn) per layer (bs / bn =
    fc_self.bias / fc_neigh.bias, either may be None); ``relus`` = per-layer ReLU flags.
    All parameter gradients are views of ONE flat buffer (``last_flat_grads``)."""

    @staticmethod
    def forward(ctx, graph, feats, relus, deterministic, *flat):
        require_cuda(feats, *flat)
        lib = _lib.load()
        L = len(relus)
        assert len(flat) == 6 * L
        feats = _row_major_2d(feats)
        N = feats.shape[0]
        dev = feats.device
        flat = tuple(None if t is None else _f32c(t) for t in flat)
        layers = (_lib.SageLayer * L)()
        for l in range(L):
            Wp, bp, Ws, bs, Wn, bn = flat[6 * l:6 * l + 6]
            layers[l].din, layers[l].dout, layers[l].relu = Ws.shape[1], Ws.shape[0], int(bool(relus[l]))
            layers[l].Wp, layers[l].bp, layers[l].Ws, layers[l].Wn = ptr(Wp), ptr(bp), ptr(Ws), ptr(Wn)
            layers[l].b, layers[l].b2 = ptr(bs), ptr(bn)
        training = any(ctx.needs_input_grad)
        mode = _gemm_mode
        ws = _workspace(lib.gts_sage_workspace_bytes(layers, L, N, int(training), mode), dev)
        indptr, indices = graph.csr
        logits = torch.empty((N, flat[6 * (L - 1) + 2].shape[0]), dtype=torch.float32, device=dev)
        check(lib.gts_sage_forward(layers, L, ptr(indptr), ptr(indices), N, ptr(feats), _ld(feats), ptr(logits),
                                   logits.shape[1], ptr(ws), ws.numel(), int(training), mode, stream_ptr()),
              "gts_sage_forward")
        _count(3 * L)
        if training:
            ctx.save_for_backward(feats, ws, *[t for t in flat if t is not None])
            ctx.present = [t is not None for t in flat]
            ctx.layers, ctx.graph, ctx.mode, ctx.deterministic = layers, graph, mode, deterministic
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = _lib.load()
        feats, ws, *saved = ctx.saved_tensors
        it = iter(saved)
        flat = [next(it) if pres else None for pres in ctx.present]
        layers, graph = ctx.layers, ctx.graph
        L = len(layers)
        N = feats.shape[0]
        dlogits = _row_major_2d(dlogits)
        # one flat buffer: [dWp | dbp | dWs | dWn | db | db2] per layer, then 2 spare floats (loss sums in DP).
        # fc_self.bias and fc_neigh.bias get the same values in SEPARATE slices: one tensor returned for two inputs
        # is cloned by autograd's AccumulateGrad, which would take that .grad out of the flat buffer.
        sizes = []
        for l in range(L):
            Wp, bp, Ws, bs, Wn, bn = flat[6 * l:6 * l + 6]
            sizes += [Wp.numel(), bp.numel(), Ws.numel(), Wn.numel(), Ws.shape[0], Ws.shape[0]]
        offs = [0]
        for sz in sizes:
            offs.append(offs[-1] + (sz + 3) // 4 * 4)             # keep every slice 16-byte aligned
        flat_g = torch.zeros(offs[-1] + 4, dtype=torch.float32, device=feats.device)
        grads = (_lib.SageLayerGrads * L)()
        outs = []
        for l in range(L):
            Wp, bp, Ws, bs, Wn, bn = flat[6 * l:6 * l + 6]
            v = [flat_g[offs[6 * l + k]:offs[6 * l + k] + sizes[6 * l + k]] for k in range(6)]
            gWp, gbp, gWs, gWn, gb, gb2 = v[0].view_as(Wp), v[1].view_as(bp), v[2].view_as(Ws), v[3].view_as(Wn), v[4], v[5]
            both = bs is not None and bn is not None
            grads[l].dWp, grads[l].dbp, grads[l].dWs, grads[l].dWn, grads[l].db = (ptr(t) for t in (gWp, gbp, gWs, gWn, gb))
            grads[l].db2 = ptr(gb2) if both else None
            outs += [gWp, gbp, gWs, gb if bs is not None else None, gWn,
                     (gb2 if both else gb) if bn is not None else None]
        dfeats = torch.empty_like(feats) if ctx.needs_input_grad[1] else None
        cptr = cidx = None
        if ctx.deterministic:
            cptr, cidx, _ = graph.csc
        check(lib.gts_sage_backward(layers, grads, L, ptr(cptr), ptr(cidx), N, ptr(feats), _ld(feats),
                                    ptr(dlogits), _ld(dlogits), ptr(dfeats), (feats.shape[1] if dfeats is not None else 0),
                                    ptr(ws), ws.numel(), ctx.mode, stream_ptr()), "gts_sage_backward")
        _count(20 * L)
        last_flat_grads["flat"], last_flat_grads["n_grad"] = flat_g, offs[-1]
        return (None, dfeats, None, None, *outs)


class SageSumLayerFn(torch.autograd.Function):
    """SAGEConv 'mean' / 'gcn' (reference model/networks.py:72-75).

    mean: out = act(h Ws^T + mean_in(h) Wn^T + b);  gcn: out = act(((sum_in(h)+h)/(deg+1)) Wn^T + b)
    (DGL applies fc_neigh before aggregating when in > out; the result is the
    same up to fp32 rounding because the aggregation is linear).
    """

    @staticmethod
    def forward(ctx, h, Ws, Wn, b, graph, agg, relu_out):
        require_cuda(h, Wn, b)
        h = _row_major_2d(h)
        indptr, indices = graph.csr
        neigh = segsum_fwd(h, indptr, indices, agg)
        act = ACT_RELU if relu_out else ACT_NONE
        if agg == "gcn":
            out = gemm_nt(neigh, Wn, bias=b, act=act)
        else:
            out = gemm_nt(h, Ws, neigh, Wn, bias=b, act=act)
        ctx.save_for_backward(h, neigh, out if relu_out else None, Ws, Wn)
        ctx.graph, ctx.agg, ctx.relu_out = graph, agg, relu_out
        return out

    @staticmethod
    def backward(ctx, dOut):
        h, neigh, out, Ws, Wn = ctx.saved_tensors
        graph, agg = ctx.graph, ctx.agg
        dZ = mask_pos(dOut, out) if ctx.relu_out else _row_major_2d(dOut)
        db = colsum(dZ)
        dWn = gemm_tn(dZ, neigh)
        dNeigh = gemm_nt(dZ, transpose(Wn))
        cptr, cidx, _ = graph.csc
        dh = segsum_bwd(dNeigh, cptr, cidx, graph.csr[0], agg)
        dWs = None
        if agg != "gcn":
            dWs = gemm_tn(dZ, h)
            dh = dh + gemm_nt(dZ, transpose(Ws))
        return dh, dWs, dWn, db, None, None, None


# ---------------------------------------------------------------------------
# GATConv layer
# ---------------------------------------------------------------------------
class GatLayerFn(torch.autograd.Function):
    """One DGL GATConv layer: fc GEMM -> scores ->
    fused edge-softmax aggregation (+residual +bias +ELU); backward = two
    deterministic edge passes + GEMMs."""

    @staticmethod
    def forward(ctx, x, W, attn_l, attn_r, bias, Wres, graph, H, F, slope, residual_identity, elu):
        require_cuda(x, W, attn_l, attn_r)
        lib = _lib.load()
        x = _row_major_2d(x)
        N = x.shape[0]
        dev = x.device
        indptr, indices = graph.csr
        Z = gemm_nt(x, W)                                   # [N, H*F]
        al = _f32c(attn_l).view(H, F)
        ar = _f32c(attn_r).view(H, F)
        el = torch.empty((N, H), dtype=torch.float32, device=dev)
        er = torch.empty((N, H), dtype=torch.float32, device=dev)
        st = stream_ptr()
        check(lib.gts_gat_scores(ptr(Z), H * F, ptr(al), ptr(ar), N, H, F, ptr(el), ptr(er), st), "gts_gat_scores")
        res = None
        if Wres is not None:
            res = gemm_nt(x, Wres)
        elif residual_identity:
            res = x
        out = torch.empty((N, H * F), dtype=torch.float32, device=dev)
        rowmax = torch.empty((N, H), dtype=torch.float32, device=dev)
        rowsum = torch.empty((N, H), dtype=torch.float32, device=dev)
        err = graph.err_flag
        bias_c = _f32c(bias) if bias is not None else None
        check(lib.gts_gat_fwd(ptr(Z), H * F, ptr(el), ptr(er), ptr(indptr), ptr(indices), N, H, F, float(slope),
                              ptr(res), (_ld(res) if res is not None else 0), ptr(bias_c), 1 if elu else 0,
                              ptr(out), H * F, ptr(rowmax), ptr(rowsum), ptr(err), st), "gts_gat_fwd")
        _count(2)
        ctx.save_for_backward(x, W, al, ar, Wres, Z, el, er, rowmax, rowsum, out if elu else None)
        ctx.graph = graph
        ctx.cfg = (H, F, float(slope), residual_identity, elu, bias is not None)
        return out.view(N, H, F)

    @staticmethod
    def backward(ctx, dOut):
        lib = _lib.load()
        x, W, al, ar, Wres, Z, el, er, rowmax, rowsum, out = ctx.saved_tensors
        H, F, slope, residual_identity, elu, has_bias = ctx.cfg
        graph = ctx.graph
        N = x.shape[0]
        dev = x.device
        st = stream_ptr()
        dOut = _f32c(dOut).view(N, H * F)
        if elu:
            dR = torch.empty_like(dOut)
            check(lib.gts_gat_act_bwd(ptr(dOut), ptr(out), dOut.numel(), 1, ptr(dR), st), "gts_gat_act_bwd")
            _count()
        else:
            dR = dOut
        indptr, indices = graph.csr
        cptr, cidx, c2r = graph.csc
        E = indices.numel()
        dt = torch.empty((E, H), dtype=torch.float32, device=dev)
        der = torch.empty((N, H), dtype=torch.float32, device=dev)
        check(lib.gts_gat_bwd_dst(ptr(Z), H * F, ptr(el), ptr(er), ptr(rowmax), ptr(rowsum), ptr(indptr), ptr(indices),
                                  ptr(dR), H * F, N, H, F, slope, ptr(dt), ptr(der), st), "gts_gat_bwd_dst")
        dZ = torch.empty((N, H * F), dtype=torch.float32, device=dev)
        del_ = torch.empty((N, H), dtype=torch.float32, device=dev)
        check(lib.gts_gat_bwd_src(ptr(el), ptr(er), ptr(rowmax), ptr(rowsum), ptr(cptr), ptr(cidx), ptr(c2r),
                                  ptr(dR), H * F, ptr(dt), ptr(der), ptr(al), ptr(ar), N, H, F, slope,
                                  ptr(dZ), H * F, ptr(del_), st), "gts_gat_bwd_src")
        ws = _workspace(2 * lib.gts_gat_attn_grad_workspace_bytes(N, H, F), dev)
        dal = torch.empty((1, H, F), dtype=torch.float32, device=dev)
        dar = torch.empty((1, H, F), dtype=torch.float32, device=dev)
        check(lib.gts_gat_attn_grad2(ptr(Z), H * F, ptr(del_), ptr(der), N, H, F, ptr(dal), ptr(dar), ptr(ws), ws.numel(), st),
              "gts_gat_attn_grad2")                       # both attention-vector gradients in one pass over Z
        _count(5)
        dW = gemm_tn(dZ, x)
        dbias = colsum(dR) if has_bias else None
        dWres = gemm_tn(dR, x) if Wres is not None else None
        dx = None
        if ctx.needs_input_grad[0]:
            if Wres is not None:
                dx = gemm_nt(dZ, transpose(W), dR, transpose(Wres))
            else:
                dx = gemm_nt(dZ, transpose(W))
                if residual_identity:
                    dx = dx + dR
        return dx, dW, dal, dar, dbias, dWres, None, None, None, None, None, None
