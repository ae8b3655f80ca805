#!/bin/bash
# gpurun --gpus 8 -- bash tools/scale_run.sh <tag>: the driver's scaling sequence (N = 1, 2, 4, 8 back to back).
tag=${1:-r02}
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/${tag}_scale_n$n.json 2> gpurun_out/${tag}_scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_scale_n$n.json 2> gpurun_out/${tag}_scale_n$n.err
  fi
done
python - "$tag" <<'PY'
import json, glob, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
base = None
for n in (1, 2, 4, 8):
    for ln in open(f"gpurun_out/{tag}_scale_n{n}.json"):
        if ln.startswith("{"):
            d = json.loads(ln)
            if n == 1: base = d
            print(n, "value %.0f graphs/s  %.3f ms/step  eff %.3f | e2e %.0f graphs/s %.3f ms eff %.3f (%s)" % (
                d["value"], d["ms_per_step"], d["value"] / (n * base["value"]), d["e2e"]["value"], d["e2e"]["ms_per_step"],
                d["e2e"]["value"] / (n * base["e2e"]["value"]), d["e2e"]["mode"]))
PY
