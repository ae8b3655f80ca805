"""GPU tests of the host-side overlap helpers: DevicePrefetcher (H2D + CSR build one batch ahead on a side stream)
and VolumeDownloader (label volumes to pinned host memory on a copy stream).  Both must be invisible in the results:
same CSR, same logits, same volumes as the plain in-stream path."""
import numpy as np
import pytest
import torch

from gnn_tumor_seg_b200 import graph as G, networks, ops, project, synth
from gnn_tumor_seg_b200.data_loader import DevicePrefetcher
from oracle import project_ref

pytestmark = pytest.mark.gpu


def _host_batches(n, pin):
    out = []
    for b in range(n):
        gs = [synth.make_small_graph(10 * b + s, n_nodes=300 + 17 * s + 5 * b, avg_deg=9) for s in range(3)]
        bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in gs], pin=pin)
        f = torch.as_tensor(np.concatenate([g.features for g in gs]))
        l = torch.as_tensor(np.concatenate([g.labels for g in gs]))
        if pin:
            f, l = f.pin_memory(), l.pin_memory()
        out.append((["mri%d" % b], bg, f, l))
    return out


@pytest.mark.parametrize("depth", [1, 2, 5])
@pytest.mark.parametrize("pin", [True, False])
def test_prefetcher_yields_what_plain_to_device_yields(cuda_dev, depth, pin):
    batches = _host_batches(4, pin)
    torch.manual_seed(0)
    net = networks.GraphSage(20, [64, 64], 4, "pool", 0).to(cuda_dev).eval()
    ops.set_gemm_mode("fp32")
    ref = []
    for ids, bg, f, l in batches:
        dg = bg.to(cuda_dev)
        with torch.no_grad():
            ref.append((ids, dg.csr[0].cpu(), dg.csr[1].cpu(), net(dg, f.to(cuda_dev)).cpu(), l.clone()))
    seen = 0
    for (ids, dg, f, l), r in zip(DevicePrefetcher(batches, cuda_dev, depth=depth), ref):
        assert ids == r[0] and f.is_cuda and l.is_cuda and dg.device.type == "cuda"
        assert torch.equal(dg.csr[0].cpu(), r[1]) and torch.equal(dg.csr[1].cpu(), r[2])
        with torch.no_grad():
            assert torch.equal(net(dg, f).cpu(), r[3])
        assert torch.equal(l.cpu(), r[4])
        seen += 1
    assert seen == len(batches)
    ops.set_gemm_mode("tf32x3")


def test_prefetcher_empty_and_generator_input(cuda_dev):
    assert list(DevicePrefetcher([], cuda_dev)) == []
    batches = _host_batches(3, True)
    got = [ids for ids, *_ in DevicePrefetcher((b for b in batches), cuda_dev, depth=2)]
    assert got == [b[0] for b in batches]


def test_prefetched_training_matches_in_stream_training(cuda_dev):
    from gnn_tumor_seg_b200 import dp
    batches = _host_batches(3, True)
    w = torch.tensor([0.1, 1.0, 2.0, 2.0], device=cuda_dev)
    ops.set_gemm_mode("fp32")
    ops.set_deterministic_backward(True)
    try:
        grads = []
        for use_pf in (False, True):
            torch.manual_seed(1)
            net = networks.GraphSage(20, [64, 64], 4, "pool", 0).to(cuda_dev)
            tr = dp.DataParallelTrainer(net, w)
            it = DevicePrefetcher(batches, cuda_dev) if use_pf else ((i, g.to(cuda_dev), f.to(cuda_dev), l.to(cuda_dev)) for i, g, f, l in batches)
            losses = [float(tr.forward_backward(dg, f, l)) for _, dg, f, l in it]
            grads.append((losses, [p.grad.clone() for p in net.parameters()]))
        # the loss sums are float atomics (last-bit noise run to run); a hand-over race would be orders larger
        assert np.allclose(grads[0][0], grads[1][0], rtol=1e-5)
        for a, b in zip(grads[0][1], grads[1][1]):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-6 * float(a.abs().max()) + 1e-12)
    finally:
        ops.set_deterministic_backward(False)
        ops.set_gemm_mode("tf32x3")


def test_volume_downloader_round_trip(cuda_dev):
    rng = np.random.default_rng(0)
    dl = project.VolumeDownloader(depth=2, device=cuda_dev)
    crop = np.ix_(np.arange(5, 17), np.arange(100, 111), np.arange(20, 30))
    inv = project.crop_inverse_maps(crop, device=cuda_dev)
    want, slots = [], []
    for i in range(5):                      # more volumes than slots: a slot is re-acquired after its copy
        svs = rng.integers(-1, 300, size=(12, 11, 10)).astype(np.int16)
        lg = rng.normal(size=(300, 4)).astype(np.float32)
        want.append(project_ref.save_voxel_preds_ref(lg, svs, crop))
        slot, vol = dl.acquire()
        project.project_labels_to_brats(torch.as_tensor(lg).to(cuda_dev), torch.as_tensor(svs).to(cuda_dev), None, out=vol, inv_maps=inv)
        dl.submit(slot)
        host = dl.wait(slot)
        assert host.is_pinned() and np.array_equal(host.numpy(), want[-1])
        slots.append(slot)
    assert slots == [0, 1, 0, 1, 0]
    dl.drain()


def test_prefetcher_explicit_stage_next(cuda_dev):
    """The consumer-driven form: stage_next() after the step's kernels are enqueued; same batches, same order."""
    batches = _host_batches(5, True)
    pf = DevicePrefetcher(batches, cuda_dev)
    got = []
    for ids, dg, f, l in pf:
        got.append((ids, int(dg.number_of_nodes()), float(f.sum())))
        pf.stage_next()
    assert [g[0] for g in got] == [b[0] for b in batches]
    for g, b in zip(got, batches):
        assert g[1] == b[1].number_of_nodes() and abs(g[2] - float(b[2].sum())) < 1e-2 * max(1.0, abs(float(b[2].sum())))


def test_staged_batch_equals_plain_to_device(cuda_dev):
    """data_loader.StagedBatch: static device buffers + copy-stream staging give the same CSR / tensors as .to(device),
    also when the buffers are re-filled while the previous contents are still being consumed."""
    from gnn_tumor_seg_b200.data_loader import StagedBatch
    from gnn_tumor_seg_b200 import graph as G2, networks as nets
    gs = [synth.make_small_graph(80 + s, n_nodes=300, avg_deg=8) for s in range(2)]
    hg = G2.batch([G2.from_edge_list(g.src, g.dst, g.n_nodes) for g in gs], pin=True)
    variants = [torch.randn(600, 20, generator=torch.Generator().manual_seed(k)).pin_memory() for k in range(3)]
    sb = StagedBatch(hg, [variants[0]], cuda_dev)
    copy = torch.cuda.Stream(device=cuda_dev)
    torch.manual_seed(0)
    net = nets.GraphSage(20, [64], 4, "pool", 0).to(cuda_dev).eval()
    ref_g = hg.to(cuda_dev)
    outs = []
    sb.load_async(hg, [variants[0]], copy)
    for k in range(3):
        g, (x,) = sb.take()
        assert torch.equal(g.csr[0], ref_g.csr[0]) and torch.equal(g.csr[1], ref_g.csr[1])
        with torch.no_grad():
            outs.append(net(g, x).clone())
        sb.release()
        if k + 1 < 3:
            sb.load_async(hg, [variants[k + 1]], copy)          # must wait for the release point above
    torch.cuda.synchronize()
    for k in range(3):
        with torch.no_grad():
            assert torch.equal(outs[k], net(ref_g, variants[k].to(cuda_dev)))
    with pytest.raises(Exception):
        sb.load_async(hg, [torch.zeros(5, 20).pin_memory()], copy)
