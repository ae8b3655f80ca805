"""Out-of-bounds write detection without compute-sanitizer (closed on this pool: `gpurun` answers "compute-sanitizer is
closed on this pool and stays closed", recorded in profiles/r02_sanitizer.md): every output of the smallest SAGE / GAT /
GEMM-edge-shape cases is allocated INSIDE a larger canary-filled buffer; after the call the canaries on both sides must
be untouched and the payload fully written (no canary value left inside it)."""
import numpy as np
import pytest
import torch

from gnn_tumor_seg_b200 import _lib, graph as G, ops, synth
from gnn_tumor_seg_b200._lib import check, ptr, stream_ptr

pytestmark = pytest.mark.gpu
CANARY = -7.0e30
PAD = 4096          # floats on each side


def _guarded(shape, dev, dtype=torch.float32):
    n = int(np.prod(shape))
    buf = torch.full((n + 2 * PAD,), CANARY if dtype == torch.float32 else -123456789, dtype=dtype, device=dev)
    return buf, buf[PAD:PAD + n].view(*shape)


def _check(buf, view, what, must_fill=True):
    canary = CANARY if buf.dtype == torch.float32 else -123456789
    n = view.numel()
    assert bool((buf[:PAD] == canary).all()) and bool((buf[PAD + n:] == canary).all()), f"{what}: wrote outside its output"
    if must_fill:
        assert not bool((view == canary).any()), f"{what}: left part of its output unwritten"


@pytest.mark.parametrize("M,N,K1,K2", [(256, 256, 256, 0), (257, 256, 32, 32), (300, 256, 520, 0), (385, 192, 256, 256),
                                       (90, 256, 20, 20), (1000, 512, 64, 0), (333, 4, 256, 256), (511, 128, 40, 0)])
@pytest.mark.parametrize("act", ["relu", "mask"])
def test_gemm_nt_writes_only_its_output(cuda_dev, M, N, K1, K2, act):
    torch.manual_seed(M + N)
    A1, B1 = torch.randn(M, K1, device=cuda_dev), torch.randn(N, K1, device=cuda_dev)
    A2 = torch.randn(M, K2, device=cuda_dev) if K2 else None
    B2 = torch.randn(N, K2, device=cuda_dev) if K2 else None
    buf, out = _guarded((M, N), cuda_dev)
    kw = dict(bias=torch.randn(N, device=cuda_dev), act=ops.ACT_RELU) if act == "relu" else \
        dict(act=ops.ACT_MASK_POS, aux=torch.randn(M, N, device=cuda_dev))
    ops.gemm_nt(A1, B1, A2, B2, out=out, mode="tf32x3", **kw)
    torch.cuda.synchronize()
    _check(buf, out, f"gemm_nt {M}x{N}x{K1}+{K2} {act}")
    ref = A1.double() @ B1.double().T + (A2.double() @ B2.double().T if K2 else 0)
    ref = torch.relu(ref + kw["bias"].double()) if act == "relu" else ref * (kw["aux"] > 0)
    assert (out.double() - ref).abs().max() <= 1e-4 * ref.abs().max().clamp_min(1e-6)


@pytest.mark.parametrize("K,Mo,No", [(300, 256, 256), (5000, 256, 128), (1234, 4, 256), (777, 256, 20), (33, 20, 20)])
def test_gemm_tn_colsum_writes_only_its_outputs(cuda_dev, K, Mo, No):
    lib = _lib.load()
    torch.manual_seed(K)
    A, B = torch.randn(K, Mo, device=cuda_dev), torch.randn(K, No, device=cuda_dev)
    cbuf, C = _guarded((Mo, No), cuda_dev)
    sbuf, cs = _guarded((Mo,), cuda_dev)
    mode = _lib.GEMM_TF32X3
    nws = lib.gts_gemm_tn_colsum_workspace_bytes(Mo, No, K, mode)
    wbuf = torch.full((nws // 4 + 2 * PAD,), CANARY, dtype=torch.float32, device=cuda_dev)
    ws = wbuf[PAD:PAD + nws // 4]
    check(lib.gts_gemm_tn_colsum(ptr(A), Mo, ptr(B), No, ptr(C), No, Mo, No, K, mode, ptr(cs), ptr(ws), nws, stream_ptr()),
          "gts_gemm_tn_colsum")
    torch.cuda.synchronize()
    _check(cbuf, C, "gemm_tn product")
    _check(sbuf, cs, "gemm_tn column sums")
    _check(wbuf, ws, "gemm_tn workspace", must_fill=False)
    ref = A.double().T @ B.double()
    assert (C.double() - ref).abs().max() <= 1e-4 * ref.abs().max()
    assert (cs.double() - A.double().sum(0)).abs().max() <= 1e-4 * A.double().sum(0).abs().max().clamp_min(1.0)


@pytest.mark.parametrize("D", [4, 20, 256, 260])
def test_segmax_fwd_bwd_write_only_their_outputs(cuda_dev, D):
    g = synth.make_small_graph(3, n_nodes=257, avg_deg=9, isolated=3)
    dg = G.from_edge_list(g.src, g.dst, g.n_nodes).to(cuda_dev)
    indptr, indices = dg.csr
    lib = _lib.load()
    N = g.n_nodes
    torch.manual_seed(D)
    P = torch.randn(N, D, device=cuda_dev)
    nbuf, neigh = _guarded((N, D), cuda_dev)
    abuf, arg = _guarded((N, D), cuda_dev, torch.int32)
    check(lib.gts_segmax_fwd(ptr(P), D, ptr(indptr), ptr(indices), N, D, ptr(neigh), D, ptr(arg), D, stream_ptr()), "segmax_fwd")
    dbuf, dP = _guarded((N, D), cuda_dev)
    dN = torch.randn(N, D, device=cuda_dev)
    check(lib.gts_segmax_bwd(ptr(dN), D, ptr(arg), D, N, D, ptr(dP), D, N, stream_ptr()), "segmax_bwd")
    torch.cuda.synchronize()
    _check(nbuf, neigh, "segmax_fwd neigh")
    _check(abuf, arg, "segmax_fwd argmax")
    _check(dbuf, dP, "segmax_bwd dP")


def test_csr_build_and_projection_write_only_their_outputs(cuda_dev):
    g = synth.make_small_graph(9, n_nodes=300, avg_deg=7)
    lib = _lib.load()
    src = torch.as_tensor(g.src).to(cuda_dev)
    dst = torch.as_tensor(g.dst).to(cuda_dev)
    E, N = src.numel(), g.n_nodes
    pbuf, indptr = _guarded((N + 1,), cuda_dev, torch.int32)
    ibuf, indices = _guarded((E,), cuda_dev, torch.int32)
    ebuf, eid = _guarded((E,), cuda_dev, torch.int32)
    nws = lib.gts_csr_build_workspace_bytes(E, N)
    ws = torch.empty(nws, dtype=torch.uint8, device=cuda_dev)
    check(lib.gts_csr_build(ptr(dst), ptr(src), E, N, ptr(indptr), ptr(indices), ptr(eid), ptr(ws), nws, stream_ptr()), "csr_build")
    torch.cuda.synchronize()
    _check(pbuf, indptr, "csr indptr"); _check(ibuf, indices, "csr indices"); _check(ebuf, eid, "csr eid")
