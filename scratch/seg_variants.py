"""A/B of the seg-max forward kernels on the bench shape (B=6 x 15k-node RAGs, D=256): each variant runs in its own
process (the dispatch env vars are read once), is checked bit-exact against the grouped kernel's outputs and timed
with CUDA events over 3 rotating 92 MB inputs (> L2).  usage: python scratch/seg_variants.py [out.json]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def _v(name, **env):
    return (name, {k: str(v) for k, v in env.items()})


VARIANTS = [_v("grouped32", GTS_SEGMAX_PIPE=0), _v("default"),
            _v("pipe32_s1_t64_nomaxl1", GTS_SEGMAX_MAXL1=0),
            _v("pipe32_s1_t128", GTS_SEGMAX_CHUNK=128), _v("pipe32_s1_t304", GTS_SEGMAX_CHUNK=304),
            _v("pipe28_s1_t64", GTS_SEGMAX_PIPE=28), _v("pipe24_s2_t64", GTS_SEGMAX_PIPE=24, GTS_SEGMAX_STAGE=2)]
if os.environ.get("SEG_VARIANTS"):
    VARIANTS = [v for v in VARIANTS if v[0] == "grouped32" or v[0] in os.environ["SEG_VARIANTS"].split(",")]


def child(name):
    import torch
    from gnn_tumor_seg_b200 import graph as G, ops, synth
    dev = torch.device("cuda:0")
    graphs = [synth.make_graph(s) for s in range(6)]
    bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in graphs]).to(dev)
    N = bg.number_of_nodes()
    gen = torch.Generator(device=dev).manual_seed(0)
    Ps = [torch.relu(torch.randn(N, 256, device=dev, generator=gen)) for _ in range(3)]
    indptr, indices = bg.csr
    n, a = ops.segmax_fwd(Ps[0], indptr, indices)
    n2, _ = ops.segmax_fwd(Ps[0], indptr, indices, want_argmax=False)
    ref = "/tmp/seg_ref.pt"
    if name == "grouped32":
        torch.save({"n": n.cpu(), "a": a.cpu()}, ref)
        exact = True
    else:
        r = torch.load(ref)
        exact = bool(torch.equal(n.cpu(), r["n"]) and torch.equal(a.cpu(), r["a"]) and torch.equal(n2.cpu(), r["n"]))
    res = {"name": name, "bit_exact": exact}
    for key, want in (("train_ms", True), ("infer_ms", False)):
        for i in range(30):
            ops.segmax_fwd(Ps[i % 3], indptr, indices, want_argmax=want)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 300
        e0.record()
        for i in range(iters):
            ops.segmax_fwd(Ps[i % 3], indptr, indices, want_argmax=want)
        e1.record()
        torch.cuda.synchronize()
        res[key] = e0.elapsed_time(e1) / iters
    E = int(indices.numel())
    res["train_gbs"] = 4 * (3 * N * 256 + N + 1 + E) / (res["train_ms"] * 1e-3) / 1e9
    res["infer_gbs"] = 4 * (2 * N * 256 + N + 1 + E) / (res["infer_ms"] * 1e-3) / 1e9
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    out = []
    for name, env in VARIANTS:
        e = dict(os.environ); e.update(env)
        p = subprocess.run([sys.executable, __file__, "--child", name], env=e, capture_output=True, text=True, timeout=600)
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        if line:
            out.append(json.loads(line[0][7:]))
            print(out[-1], flush=True)
        else:
            print(name, "FAILED", p.stderr[-2000:], flush=True)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
