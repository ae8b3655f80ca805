"""bench.py's host-side pieces that run without a GPU: the reference arm's contract (the task's
measurement contract) and the ncu-table look-ups behind `roofline.traffic`."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          env=e, cwd=ROOT, timeout=600)


def test_reference_arm_line_carries_the_contract_keys():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--batch", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "graphs_per_s" and d["unit"] == "graphs/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["value"] > 0
    assert abs(d["value"] - 1.0 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]          # batch of 1 graph per step
    cb, e2e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "fwd+CE+bwd+AdamW" in cb["sample"]
    assert e2e == {"value": d["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("GraphSAGE-pool 7x256 training") and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29871"})
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_own_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-extras"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_ncu_table_lookups():
    b = _bench()
    table = {"gat_fwd_kernel<1, 1, 1>": {"dram_bytes": 10.0}, "gat_fwd_kernel<2, 1, 1>": {"dram_bytes": 1000.0},
             "segmax_fwd_pipe_kernel<2, 1, 32, 3, 1>": {"dram_bytes": 7.0}, "no_bytes": {"duration_us": 1.0}}
    assert b.traffic_of(table, "segmax_fwd_pipe") == 7.0
    assert b.traffic_of(table, "no_bytes") is None and b.traffic_of(table, "absent") is None
    assert b.gat_traffic(table, "gat_fwd_kernel") == len(b.GAT_LAYER_SIZES) * 1000.0 + 10.0     # 4 wide layers + the output layer
    assert b.gat_traffic(table, "gat_bwd_dst_kernel") is None
    # the committed captures answer the look-ups bench.py makes
    real = b.load_ncu_traffic()
    for needle in ("segmax_fwd_pipe", "segmax_bwd_vec", "gemm_x3ntw_kernel"):
        assert b.traffic_of(real, needle) > 0
    for k in ("gat_fwd_kernel", "gat_bwd_dst_kernel", "gat_bwd_src_kernel"):
        assert b.gat_traffic(real, k) > 1e9


def test_algorithmic_work_of_the_headline_config():
    """7x256 stack on 90 000 nodes — the flop count the roofline line divides by."""
    b = _bench()
    n, e = 90000, 1340236
    flops = b.model_flops_bytes(n, e)
    flops = flops[0] if isinstance(flops, tuple) else flops
    fwd = 0
    dims = [b.IN_FEATS] + list(b.LAYER_SIZES) + [b.N_CLASSES]
    for din, dout in zip(dims[:-1], dims[1:]):
        fwd += 2 * n * (din * din + 2 * din * dout)                # fc_pool (din x din), fc_self + fc_neigh (din x dout)
    assert 2.5 * fwd <= flops <= 3.0 * fwd + 1                      # backward = 2 GEMMs per forward GEMM, minus layer 0's dh
