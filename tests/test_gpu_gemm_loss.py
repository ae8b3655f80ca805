"""GPU parity: dense contractions (all arithmetic modes), column sums,
transposes, weighted CE — against fp64 PyTorch on the CPU.
Tolerances (relative to the largest reference magnitude): fp32 1e-5,
tf32x3 2e-5 (fp32-accurate 3xTF32), tf32 2e-3 (10-bit mantissa inputs)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gnn_tumor_seg_b200 import ops
from gnn_tumor_seg_b200._lib import ACT_MASK_POS, ACT_NONE, ACT_RELU

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-5, "tf32x3": 2e-5, "tf32": 2e-3}


def _rel(a, b):
    return (a.double() - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("mode", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("M,N,K1,K2", [(1000, 256, 256, 256), (333, 256, 20, 20), (700, 4, 256, 256),
                                       (129, 256, 256, 0), (4097, 128, 64, 32), (5, 16, 8, 0), (2048, 256, 4, 256)])
def test_gemm_nt(cuda_dev, mode, M, N, K1, K2):
    g = torch.Generator().manual_seed(M + N + K1)
    A1 = torch.randn(M, K1, generator=g); B1 = torch.randn(N, K1, generator=g)
    A2 = torch.randn(M, K2, generator=g) if K2 else None
    B2 = torch.randn(N, K2, generator=g) if K2 else None
    bias = torch.randn(N, generator=g); aux = torch.randn(M, N, generator=g)
    ref = A1.double() @ B1.double().T + bias.double()
    if K2:
        ref = ref + A2.double() @ B2.double().T
    dv = lambda t: None if t is None else t.to(cuda_dev)
    out = ops.gemm_nt(dv(A1), dv(B1), dv(A2), dv(B2), bias=dv(bias), act=ACT_NONE, mode=mode)
    assert _rel(out.cpu(), ref) < TOL[mode]
    out = ops.gemm_nt(dv(A1), dv(B1), dv(A2), dv(B2), bias=dv(bias), act=ACT_RELU, mode=mode)
    assert _rel(out.cpu(), torch.relu(ref)) < TOL[mode]
    out = ops.gemm_nt(dv(A1), dv(B1), dv(A2), dv(B2), bias=None, act=ACT_MASK_POS, aux=dv(aux), mode=mode)
    assert _rel(out.cpu(), (ref - bias.double()) * (aux > 0)) < TOL[mode]


@pytest.mark.parametrize("M,N,K1,K2", [(256, 256, 256, 0), (257, 128, 32, 0), (300, 256, 520, 0), (385, 192, 256, 256),
                                       (511, 1024, 1024, 0), (90000, 256, 20, 0), (640, 320, 96, 64)])
def test_gemm_nt_pair_kernel_edges(cuda_dev, M, N, K1, K2):
    """Shapes that take the CTA-pair kernel (M >= 256, N >= 128, N % 64 == 0) at its edges: a peer CTA whose 128 rows
    are partly or wholly outside C, K tails, N = 192 / 320 (a partly empty second N tile), N = 1024 (8 N tiles)."""
    g = torch.Generator().manual_seed(7 * M + N + K1)
    A1 = torch.randn(M, K1, generator=g); B1 = torch.randn(N, K1, generator=g)
    A2 = torch.randn(M, K2, generator=g) if K2 else None
    B2 = torch.randn(N, K2, generator=g) if K2 else None
    bias = torch.randn(N, generator=g); aux = torch.randn(M, N, generator=g)
    ref = A1.double() @ B1.double().T + bias.double()
    if K2:
        ref = ref + A2.double() @ B2.double().T
    dv = lambda t: None if t is None else t.to(cuda_dev)
    out = ops.gemm_nt(dv(A1), dv(B1), dv(A2), dv(B2), bias=dv(bias), act=ACT_RELU, mode="tf32x3")
    assert _rel(out.cpu(), torch.relu(ref)) < TOL["tf32x3"]
    out = ops.gemm_nt(dv(A1), dv(B1), dv(A2), dv(B2), bias=None, act=ACT_MASK_POS, aux=dv(aux), mode="tf32x3")
    assert _rel(out.cpu(), (ref - bias.double()) * (aux > 0)) < TOL["tf32x3"]
    # output rows beyond M must stay untouched: write into the top of a larger buffer
    big = torch.full((M + 130, N), 7.0, device=cuda_dev)
    ops.gemm_nt(dv(A1), dv(B1), dv(A2), dv(B2), bias=dv(bias), act=ACT_NONE, mode="tf32x3", out=big[:M])
    assert torch.all(big[M:] == 7.0) and _rel(big[:M].cpu(), ref) < TOL["tf32x3"]


@pytest.mark.parametrize("mode", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("K,Mo,No", [(5000, 256, 256), (90000, 256, 256), (1234, 4, 256), (777, 256, 20), (300, 20, 20),
                                     (40000, 128, 64)])
def test_gemm_tn(cuda_dev, mode, K, Mo, No):
    g = torch.Generator().manual_seed(K + Mo)
    A = torch.randn(K, Mo, generator=g); B = torch.randn(K, No, generator=g)
    ref = A.double().T @ B.double()
    out = ops.gemm_tn(A.to(cuda_dev), B.to(cuda_dev), mode=mode)
    assert _rel(out.cpu(), ref) < TOL[mode]


@pytest.mark.parametrize("mode", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("K,Mo,No", [(90000, 256, 256), (1234, 4, 256), (777, 256, 20), (33, 20, 20), (40001, 128, 64)])
def test_gemm_tn_colsum(cuda_dev, mode, K, Mo, No):
    """Weight gradient + bias gradient from one entry point (fused in the tf32x3 mode: the column
    sums come out of the warps that stage A through tensor memory)."""
    g = torch.Generator().manual_seed(K + 7 * Mo)
    A = torch.randn(K, Mo, generator=g); B = torch.randn(K, No, generator=g)
    out, cs = ops.gemm_tn_colsum(A.to(cuda_dev), B.to(cuda_dev), mode=mode)
    assert _rel(out.cpu(), A.double().T @ B.double()) < TOL[mode]
    assert _rel(cs.cpu(), A.double().sum(0)) < 1e-5          # column sums are plain fp32 adds in every mode


@pytest.mark.parametrize("mode", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("K,Mo,No", [(90000, 256, 256), (5000, 256, 128), (1234, 4, 256), (777, 256, 20), (40001, 192, 64),
                                     (31, 256, 256), (33, 512, 192), (100000, 1024, 1024)])
def test_gemm_tn2_colsum(cuda_dev, mode, K, Mo, No):
    """Two weight gradients sharing A + its column sums from one entry point (one pass over A in tf32x3)."""
    g = torch.Generator().manual_seed(K + 3 * Mo + No)
    A = torch.randn(K, Mo, generator=g); B1 = torch.randn(K, No, generator=g); B2 = torch.randn(K, No, generator=g)
    c1, c2, cs = ops.gemm_tn2_colsum(A.to(cuda_dev), B1.to(cuda_dev), B2.to(cuda_dev), mode=mode)
    # max-norm over a million outputs of 100 000-term sums: 2.9e-5 measured for tf32x3 (the tensor core's fp32
    # accumulation; the split plan caps one accumulation chain at 4096 rows — 3.7e-4 without the cap)
    tol = max(TOL[mode], 5e-5) if Mo * No >= (1 << 20) else TOL[mode]
    assert _rel(c1.cpu(), A.double().T @ B1.double()) < tol
    assert _rel(c2.cpu(), A.double().T @ B2.double()) < tol
    assert _rel(cs.cpu(), A.double().sum(0)) < 1e-5
    Ai = torch.randint(-3, 4, (3000, 256), generator=g).float(); Bi = torch.randint(-3, 4, (3000, 256), generator=g).float()
    c1, c2, cs = ops.gemm_tn2_colsum(Ai.to(cuda_dev), Bi.to(cuda_dev), Ai.to(cuda_dev), mode=mode)
    assert torch.equal(c1.cpu(), Ai.T @ Bi) and torch.equal(c2.cpu(), Ai.T @ Ai) and torch.equal(cs.cpu(), Ai.sum(0))


def test_gemm_tn_colsum_integer_exact(cuda_dev):
    """Integer-valued operands: products and sums are exact in every arithmetic mode -> bit-exact."""
    g = torch.Generator().manual_seed(3)
    A = torch.randint(-3, 4, (5000, 256), generator=g).float(); B = torch.randint(-3, 4, (5000, 256), generator=g).float()
    for mode in ("fp32", "tf32x3", "tf32"):
        out, cs = ops.gemm_tn_colsum(A.to(cuda_dev), B.to(cuda_dev), mode=mode)
        assert torch.equal(out.cpu(), A.T @ B)
        assert torch.equal(cs.cpu(), A.sum(0))


@pytest.mark.parametrize("mode", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("M,N,K,R", [(90000, 256, 256, 90000), (5000, 256, 256, 777), (1000, 20, 256, 1000), (300, 4, 256, 300)])
def test_gemm_nt_scatter(cuda_dev, mode, M, N, K, R):
    """GTS_ACT_MASK_POS_SCATTER == masked GEMM followed by the arg-max scatter (segmax_bwd), also vs the fp64 oracle."""
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g); B = torch.randn(N, K, generator=g)
    aux = torch.randn(M, N, generator=g)
    idx = torch.randint(-1, R, (M, N), generator=g, dtype=torch.int64).to(torch.int32)
    ref = torch.zeros(R, N, dtype=torch.float64)
    v = (A.double() @ B.double().T) * (aux > 0)
    ok = idx >= 0
    cols = torch.arange(N).expand(M, N)
    ref.index_put_((idx[ok].long(), cols[ok]), v[ok], accumulate=True)
    dv = lambda t: t.to(cuda_dev)
    out = ops.gemm_nt_scatter(dv(A), dv(B), dv(aux), dv(idx), R, mode=mode)
    assert _rel(out.cpu(), ref) < TOL[mode]
    two = ops.segmax_bwd(ops.gemm_nt(dv(A), dv(B), act=ACT_MASK_POS, aux=dv(aux), mode=mode), dv(idx), R)
    assert _rel(out.cpu(), two.cpu().double()) < max(1e-5, TOL[mode])      # the fused mode always runs the exact fp32 kernel


def test_gemm_strided_operands(cuda_dev):
    big = torch.randn(300, 512)
    A = big[:, 128:384]                       # ld 512
    W = torch.randn(256, 256)
    ref = A.double() @ W.double().T
    for mode in ("fp32", "tf32x3"):
        out = ops.gemm_nt(big.to(cuda_dev)[:, 128:384], W.to(cuda_dev), mode=mode)
        assert _rel(out.cpu(), ref) < TOL[mode]


def test_colsum_transpose_maskpos(cuda_dev):
    for rows, cols in [(90000, 256), (1000, 4), (1, 20), (4097, 33)]:
        A = torch.randn(rows, cols)
        assert _rel(ops.colsum(A.to(cuda_dev)).cpu(), A.double().sum(0)) < 1e-5
    W = torch.randn(37, 256)
    assert torch.equal(ops.transpose(W.to(cuda_dev)).cpu(), W.T.contiguous())
    g, r = torch.randn(1000, 7), torch.randn(1000, 7)
    assert torch.equal(ops.mask_pos(g.to(cuda_dev), r.to(cuda_dev)).cpu(), g * (r > 0))


def test_weighted_ce_forward_backward(cuda_dev):
    torch.manual_seed(0)
    z = (3 * torch.randn(5000, 4)).requires_grad_(True)
    y = torch.randint(0, 4, (5000,)); w = torch.tensor([0.1, 1., 2., 2.])
    ref = F.cross_entropy(z.double(), y, weight=w.double())
    ref.backward()
    zd = z.detach().to(cuda_dev).requires_grad_(True)
    loss = ops.weighted_cross_entropy(zd, y.to(cuda_dev), w.to(cuda_dev))
    loss.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert _rel(zd.grad.cpu(), z.grad.double()) < 1e-5


def test_adamw_step_matches_torch(cuda_dev):
    from gnn_tumor_seg_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(1)
    p0 = torch.randn(10000); grads = [torch.randn(10000) for _ in range(3)]
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([p_ref], lr=1e-2, weight_decay=1e-2)
    p = p0.clone().to(cuda_dev); m = torch.zeros_like(p); v = torch.zeros_like(p)
    for t, g in enumerate(grads, 1):
        p_ref.grad = g.clone(); opt.step()
        gd = g.to(cuda_dev)
        _lib.check(lib.gts_adamw_step(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                                      1e-2, 0.9, 0.999, 1e-8, 1e-2, t, 1.0, None, _lib.stream_ptr()))
    assert torch.allclose(p.cpu(), p_ref.detach(), atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("bf", ["0", "1", "2"])
def test_gemm_nt_x3_cross_term_variants_against_fp64(cuda_dev, bf):
    """GTS_X3_BF16 picks the form of the 3xTF32 cross terms in the CTA-pair NT kernel (0: all TF32, 1: bf16 MMAs -
    the default, 2: bf16 + hi*hi from shared memory with an 8-slot TMEM ring).  The switch is read once per process,
    so each variant runs in its own interpreter; all must stay fp32-accurate against an fp64 product."""
    import os, subprocess, sys
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from gnn_tumor_seg_b200 import ops
torch.manual_seed(0)
dev = torch.device("cuda:0")
M, K, N = 1000, 256, 256                       # M not a multiple of the 256-row pair tile
A1, A2 = torch.randn(M, K, device=dev), torch.randn(M, K, device=dev) * 3
W1, W2 = torch.randn(N, K, device=dev) / 16, torch.randn(N, K, device=dev) / 16
b = torch.randn(N, device=dev)
out = ops.gemm_nt(A1, W1, A2, W2, bias=b, act=0, mode="tf32x3")
ref = A1.double() @ W1.double().t() + A2.double() @ W2.double().t() + b.double()
err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
print("ERR", err)
assert err < 2e-5, err
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GTS_X3_BF16=bf)
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
