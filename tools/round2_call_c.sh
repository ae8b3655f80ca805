#!/bin/bash
# gpurun --gpus 4 -- bash tools/round2_call_c.sh : the driver's own scaling command at N = 4 (peer-memory exchange with
# more than one peer per rank: IPC mapping of 3 buffers, flag fan-out, rank-order sum over 4 arenas).
out=gpurun_out; mkdir -p $out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 150 $T --master-port 29614 bench.py --gpus 4 --steps 20 --warmup 5 > $out/r02d_n4_peer.json 2> $out/r02d_n4_peer.err; echo "n4 peer rc=$?"
python - <<'PY'
import json
for ln in open("gpurun_out/r02d_n4_peer.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("n4 ms %.4f value %.1f | e2e ms %.4f | eager %.4f | %s %s" % (
            d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d.get("eager_ms_per_step", 0),
            (d.get("dp_exchange") or "")[:40], d.get("dp_peer_status")))
PY
tail -3 $out/r02d_n4_peer.err
