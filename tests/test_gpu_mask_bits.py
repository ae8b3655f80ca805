"""ReLU masks as bit matrices (GTS_ACT_MASK_BITS, include/gts.h): the forward concat GEMM's epilogue and the seg-max
forward write them, the backward's data-gradient GEMMs consume them instead of re-reading the 4-byte activations.
Bit-exact against the float forms."""
import numpy as np
import pytest
import torch

from gnn_tumor_seg_b200 import _lib, graph as G, ops, synth
from gnn_tumor_seg_b200._lib import check, ptr, stream_ptr

pytestmark = pytest.mark.gpu


def unpack_bits(bits, n_cols):
    """bool [M, n_cols] from the int32 bit matrix: word (m, n/32), bit b <-> column 32*(n/32) + 4*(b & 7) + (b >> 3)."""
    M = bits.shape[0]
    b = torch.arange(32, device=bits.device)
    col_in_word = 4 * (b & 7) + (b >> 3)
    words = bits.view(M, n_cols // 32, 1)
    on = ((words >> b.view(1, 1, 32)) & 1).bool()                      # [M, W, 32] indexed by bit
    out = torch.empty((M, n_cols // 32, 32), dtype=torch.bool, device=bits.device)
    out[:, :, col_in_word] = on
    return out.view(M, n_cols)


@pytest.mark.parametrize("M,N,K1,K2", [(1000, 256, 256, 256), (257, 256, 64, 0), (4000, 512, 128, 0), (90, 256, 32, 0)])
def test_relu_bits_out_and_mask_bits_equal_float_forms(cuda_dev, M, N, K1, K2):
    lib = _lib.load()
    if not lib.gts_gemm_nt_bits_supported(M, N, _lib.GEMM_TF32X3):
        assert M < 256                                   # only shapes outside the CTA-pair kernel may decline
        return
    torch.manual_seed(M)
    A1, B1 = torch.randn(M, K1, device=cuda_dev), torch.randn(N, K1, device=cuda_dev)
    A2 = torch.randn(M, K2, device=cuda_dev) if K2 else None
    B2 = torch.randn(N, K2, device=cuda_dev) if K2 else None
    bias = torch.randn(N, device=cuda_dev)
    bits = torch.full((M, N // 32), -1, dtype=torch.int32, device=cuda_dev)
    out = ops.gemm_nt(A1, B1, A2, B2, bias=bias, act=ops.ACT_RELU, mode="tf32x3", relu_bits_out=bits)
    ref = ops.gemm_nt(A1, B1, A2, B2, bias=bias, act=ops.ACT_RELU, mode="tf32x3")
    assert torch.equal(out, ref)
    assert torch.equal(unpack_bits(bits, N), out > 0)
    # consumer: mask by bits == mask by the float activation
    g = torch.randn(M, K1, device=cuda_dev)
    W = torch.randn(N, K1, device=cuda_dev)
    by_float = ops.gemm_nt(g, W, act=ops.ACT_MASK_POS, aux=out, mode="tf32x3")
    by_bits = ops.gemm_nt(g, W, act=ops.ACT_MASK_BITS, aux_bits=bits, mode="tf32x3")
    assert torch.equal(by_float, by_bits)
    assert bool((by_bits[out <= 0] == 0).all())


def test_segmax_fwd_bits_match_neigh_positive(cuda_dev):
    lib = _lib.load()
    g = synth.make_small_graph(21, n_nodes=1500, avg_deg=11, isolated=5)
    dg = G.from_edge_list(g.src, g.dst, g.n_nodes).to(cuda_dev)
    indptr, indices = dg.csr
    N, D = g.n_nodes, 256
    assert lib.gts_segmax_fwd_bits_supported(N, D, D)
    torch.manual_seed(1)
    P = torch.relu(torch.randn(N, D, device=cuda_dev) - 0.8)              # plenty of exact zeros
    neigh = torch.empty(N, D, device=cuda_dev)
    arg = torch.empty(N, D, dtype=torch.int32, device=cuda_dev)
    bits = torch.full((N, D // 32), -1, dtype=torch.int32, device=cuda_dev)
    check(lib.gts_segmax_fwd_bits(ptr(P), D, ptr(indptr), ptr(indices), N, D, ptr(neigh), D, ptr(arg), D, ptr(bits), D // 32,
                                  stream_ptr()), "gts_segmax_fwd_bits")
    n2, a2 = ops.segmax_fwd(P, indptr, indices, want_argmax=True)
    assert torch.equal(neigh, n2) and torch.equal(arg, a2)
    assert torch.equal(unpack_bits(bits, D), neigh > 0)
    assert not lib.gts_segmax_fwd_bits_supported(N, 20, 20)


def test_stack_with_bit_masks_equals_stack_with_float_masks(cuda_dev, monkeypatch):
    """The whole-stack backward (gts_sage_forward/backward) gives the same gradients with the bit masks as the per-layer
    autograd path, which uses the float masks."""
    from gnn_tumor_seg_b200 import networks
    graphs = [synth.make_small_graph(s, n_nodes=700 + 40 * s, avg_deg=9) for s in range(3)]
    bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in graphs]).to(cuda_dev)
    x = torch.as_tensor(np.concatenate([g.features for g in graphs])).to(cuda_dev)
    y = torch.as_tensor(np.concatenate([g.labels for g in graphs])).to(cuda_dev)
    w = torch.tensor([0.1, 1.0, 2.0, 2.0], device=cuda_dev)
    torch.manual_seed(0)
    net = networks.GraphSage(20, [256, 256, 256], 4, "pool", 0).to(cuda_dev)
    ops.set_deterministic_backward(True)
    try:
        grads = {}
        for stack in (True, False):
            ops.set_stack_path(stack)
            net.zero_grad()
            ops.weighted_cross_entropy(net(bg, x), y, w).backward()
            grads[stack] = {n: p.grad.clone() for n, p in net.named_parameters()}
        for n in grads[True]:
            assert (grads[True][n] - grads[False][n]).abs().max() <= 1e-6 * grads[False][n].abs().max(), n
    finally:
        ops.set_stack_path(True)
        ops.set_deterministic_backward(False)


@pytest.mark.parametrize("M,K", [(1000, 256), (90000, 256), (333, 64)])
def test_fused_scatter_epilogue_equals_gemm_then_scatter(cuda_dev, M, K):
    """GTS_ACT_MASK_BITS_SCATTER: dP[arg[v,k],k] += ((A W^T) * mask)[v,k] out of the GEMM epilogue == masked GEMM followed
    by gts_segmax_bwd (up to the order of the fp32 additions)."""
    import ctypes as C
    lib = _lib.load()
    N = 256
    if not lib.gts_gemm_nt_bits_supported(M, N, _lib.GEMM_TF32X3):
        pytest.skip("shape outside the 256-wide kernel")
    torch.manual_seed(M)
    A = torch.randn(M, K, device=cuda_dev)
    W = torch.randn(N, K, device=cuda_dev)
    act = torch.randn(M, N, device=cuda_dev)
    # bits of (act > 0) through the product's own producer
    bits = torch.zeros((M, N // 32), dtype=torch.int32, device=cuda_dev)
    # producer: a ReLU GEMM whose output is relu(act) (B = identity)
    I = torch.eye(N, device=cuda_dev)
    out = ops.gemm_nt(act, I, act=ops.ACT_RELU, mode="tf32x3", relu_bits_out=bits)
    assert torch.equal(out > 0, act > 0)
    R = M
    idx = torch.randint(-1, R, (M, N), device=cuda_dev, dtype=torch.int32)
    # two-kernel form
    dN = ops.gemm_nt(A, W, act=ops.ACT_MASK_BITS, aux_bits=bits, mode="tf32x3")
    ref = ops.segmax_bwd(dN, idx, R)
    # fused form
    dst = torch.zeros(R, N, device=cuda_dev)
    a = _lib.GemmNtArgs()
    a.A1, a.lda1, a.K1 = ptr(A), K, K
    a.B1, a.ldb1 = ptr(W), K
    a.M, a.N, a.act, a.mode = M, N, _lib.ACT_MASK_BITS_SCATTER, _lib.GEMM_TF32X3
    a.aux_bits, a.ld_aux_bits = ptr(bits), N // 32
    a.scatter_idx, a.ld_idx, a.scatter_out, a.ld_out = ptr(idx), N, ptr(dst), N
    check(lib.gts_gemm_nt(C.byref(a), stream_ptr()), "gts_gemm_nt scatter")
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert (dst - ref).abs().max().item() <= 2e-5 * scale


@pytest.mark.parametrize("M,N,K,mode", [(3000, 256, 256, "tf32x3"), (3000, 256, 256, "fp32"), (100, 64, 32, "tf32x3"),
                                        (90000, 256, 512, "tf32x3")])
def test_gemm_zero_fill_side_job(cuda_dev, M, N, K, mode):
    """gts_gemm_nt_args.zero_fill: the buffer is zero after the call whichever kernel ran (spare warps of the 256-wide
    kernel, memset in front of the others), the product is unchanged, neighbours of the buffer are untouched."""
    torch.manual_seed(7)
    A, B = torch.randn(M, K, device=cuda_dev), torch.randn(N, K, device=cuda_dev)
    ref = ops.gemm_nt(A, B, mode=mode)
    guard = 1024
    n = M * N + 12                                        # not a multiple of the kernel's 4-way unrolled stride
    buf = torch.full((n + 2 * guard,), 3.5, device=cuda_dev)
    out = ops.gemm_nt(A, B, mode=mode, zero_fill=buf[guard:guard + n])
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    assert int(torch.count_nonzero(buf[guard:guard + n])) == 0
    assert bool((buf[:guard] == 3.5).all()) and bool((buf[guard + n:] == 3.5).all())


def test_segmax_bwd_add_accumulates(cuda_dev):
    """gts_segmax_bwd_add = the scatter half of gts_segmax_bwd: adds into what the caller put there."""
    lib = _lib.load()
    g = synth.make_small_graph(5, n_nodes=2000, avg_deg=9, isolated=3)
    dg = G.from_edge_list(g.src, g.dst, g.n_nodes).to(cuda_dev)
    N, D = g.n_nodes, 128
    torch.manual_seed(2)
    P = torch.randn(N, D, device=cuda_dev)
    indptr, indices = dg.csr
    neigh, arg = ops.segmax_fwd(P, indptr, indices)
    dN = torch.randn(N, D, device=cuda_dev)
    dP = ops.segmax_bwd(dN, arg, N)
    base = torch.randn(N, D, device=cuda_dev)
    acc = base.clone()
    check(lib.gts_segmax_bwd_add(ptr(dN), D, ptr(arg), D, N, D, ptr(acc), D, stream_ptr()), "gts_segmax_bwd_add")
    torch.testing.assert_close(acc - base, dP, rtol=0, atol=2e-5)
