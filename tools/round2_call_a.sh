#!/bin/bash
# gpurun --gpus 2 -- bash tools/round2_call_a.sh : the peer-memory gradient exchange on two B200s — its tests, then the
# scaling A/B on ONE box: N = 1, N = 2 over peer memory, N = 2 over NCCL (GTS_DP_PEER=0).
out=gpurun_out; mkdir -p $out
timeout 400 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -x -q -k "peer" > $out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r02a_pytest.log
timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-extras > $out/r02a_n1.json 2> $out/r02a_n1.err; echo "n1 rc=$?"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $T --master-port 29612 bench.py --gpus 2 --steps 30 --warmup 5 > $out/r02a_n2_peer.json 2> $out/r02a_n2_peer.err; echo "n2 peer rc=$?"
GTS_DP_PEER=0 timeout 240 $T --master-port 29613 bench.py --gpus 2 --steps 30 --warmup 5 > $out/r02a_n2_nccl.json 2> $out/r02a_n2_nccl.err; echo "n2 nccl rc=$?"
python - <<'PY'
import json
base = None
for nm in ("n1", "n2_peer", "n2_nccl"):
    try:
        for ln in open(f"gpurun_out/r02a_{nm}.json"):
            if ln.startswith("{"):
                d = json.loads(ln)
                if nm == "n1": base = d
                n = d["n_gpus"]
                print(nm, "ms %.4f value %.1f eff %.3f | e2e ms %.4f eff %.3f | eager %.4f | %s %s" % (
                    d["ms_per_step"], d["value"], d["value"] / (n * base["value"]), d["e2e"]["ms_per_step"],
                    d["e2e"]["value"] / (n * base["e2e"]["value"]), d.get("eager_ms_per_step", 0),
                    (d.get("dp_exchange") or "")[:40], d.get("dp_peer_status")))
    except Exception as e:
        print(nm, "unreadable", e)
PY
tail -5 $out/r02a_n2_peer.err
