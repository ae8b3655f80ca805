"""Host logic of the one-call step (no GPU): bucket plan of the data-parallel backward split."""
from gnn_tumor_seg_b200.trainer import plan_buckets


def test_plan_buckets_covers_every_layer_once_top_first():
    off = [0, 10, 20, 30, 40, 50, 60, 70, 80]
    b = plan_buckets(off, 8, 2)
    assert b == [(8, 4, 40, 80), (4, 0, 0, 40)]
    for n_layers, n_buckets in [(8, 1), (8, 3), (2, 4), (1, 2), (5, 5)]:
        o = list(range(0, 10 * (n_layers + 1), 10))
        bs = plan_buckets(o, n_layers, n_buckets)
        layers = [l for hi, lo, _, _ in bs for l in range(lo, hi)]
        assert sorted(layers) == list(range(n_layers)) and len(layers) == n_layers
        assert bs[0][0] == n_layers and bs[-1][1] == 0
        assert all(bs[i][1] == bs[i + 1][0] for i in range(len(bs) - 1))          # contiguous, descending
        assert all(g_lo == o[lo] and g_hi == o[hi] for hi, lo, g_lo, g_hi in bs)
