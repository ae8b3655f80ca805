"""Data-parallel host logic on CPU: world_size-2 gloo run of the trainer
(shard by whole graphs, un-normalised loss, one flat all-reduce, divide by the
global weight sum) must equal the single-process step on the union batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_tumor_seg_b200 import dp, synth
from oracle import graph_ref, sage_ref

W = torch.tensor([0.1, 1.0, 2.0, 2.0])


def _cpu_loss_sums(logits, labels, w):
    logp = torch.log_softmax(logits, 1)
    wy = w[labels]
    nll = -logp.gather(1, labels.view(-1, 1)).squeeze(1)
    return torch.stack([(wy * nll).sum(), wy.sum()])


def _make(seeds):
    gs = [synth.make_small_graph(s, n_nodes=60 + 5 * s, avg_deg=5) for s in seeds]
    s, d, n, _, _ = graph_ref.batch_graphs_ref([(g.src, g.dst, g.n_nodes) for g in gs])
    csr = graph_ref.csr_by_dst_ref(s, d, n)[:2]
    feats = torch.as_tensor(np.concatenate([g.features for g in gs]))
    labels = torch.as_tensor(np.concatenate([g.labels for g in gs]))
    return csr, feats, labels


def _net():
    torch.manual_seed(0)
    return sage_ref.GraphSageRef(20, [16, 16], 4)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _net()
        seeds = list(range(6))
        mine = [seeds[i] for i in dp.shard_indices(len(seeds), rank, world)]
        csr, feats, labels = _make(mine)
        tr = dp.DataParallelTrainer(net, W, loss_sums_fn=_cpu_loss_sums)
        loss = tr.forward_backward(csr, feats, labels)
        # the peer-memory exchange is an NVLink / CUDA-IPC mechanism: under gloo every rank gets None before any
        # collective of its own is issued (no hang, no half-built exchange) and the all-reduce path above is what runs
        from gnn_tumor_seg_b200.peer import PeerExchange
        peer_none = PeerExchange.try_create(1024) is None
        torch.save({"loss": loss, "grads": tr.grads.clone(), "extra": tr.extra.clone(), "peer_none": peer_none},
                   os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_indices():
    assert dp.shard_indices(7, 0, 2) == [0, 2, 4, 6] and dp.shard_indices(7, 1, 2) == [1, 3, 5]
    assert dp.shard_indices(1251, 7, 8)[:2] == [7, 15]
    assert sum(len(dp.shard_indices(1251, r, 8)) for r in range(8)) == 1251
    assert dp.shard_indices(3, 5, 8) == []


def test_grad_arena_views():
    net = _net()
    arena = dp.GradArena(net.parameters())
    assert arena.flat.numel() == sum(p.numel() for p in net.parameters()) + 2
    p0 = next(net.parameters())
    p0.grad.fill_(3.0)
    assert float(arena.flat[0]) == 3.0
    arena.zero_()
    assert float(p0.grad.abs().sum()) == 0.0


def test_world2_equals_single_process_union_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "r0.pt"); r1 = torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["grads"], r1["grads"]) and torch.equal(r0["loss"], r1["loss"])     # bitwise equal across ranks
    assert r0["peer_none"] and r1["peer_none"]
    # single process, union batch in any graph order: weighted mean over ALL nodes
    net = _net()
    csr, feats, labels = _make(list(range(6)))
    logits = net(csr, feats)
    loss = torch.nn.functional.cross_entropy(logits, labels, weight=W)
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert abs(float(r0["loss"]) - float(loss)) < 1e-5 * max(1.0, abs(float(loss)))
    assert torch.allclose(r0["grads"], ref, atol=1e-6, rtol=1e-4)
    # naive averaging of per-rank mean losses would NOT match (data-dependent denominators)
    assert float(r0["extra"][1]) == pytest.approx(float(W[labels].sum()), rel=1e-6)


def test_trainer_hands_back_views_of_the_reduced_buffer():
    net = _net()
    csr, feats, labels = _make([0, 1])
    tr = dp.DataParallelTrainer(net, W, loss_sums_fn=_cpu_loss_sums)
    loss = tr.forward_backward(csr, feats, labels)
    ref = _net()
    l2 = torch.nn.functional.cross_entropy(ref(csr, feats), labels, weight=W)
    l2.backward()
    assert abs(float(loss) - float(l2)) < 1e-5
    for p, q in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6, rtol=1e-4)
        assert p.grad.untyped_storage().data_ptr() == tr.flat.untyped_storage().data_ptr()
