#!/bin/bash
# gpurun -- bash tools/round2_call_d.sh : what the driver runs at round end, on the final tree — the GPU test suite,
# smoke(), the default bench line (riders and CPU baseline included).
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/r02e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02e_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/r02e_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/r02e_smoke.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/r02e_bench.json 2> $out/r02e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for ln in open("gpurun_out/r02e_bench.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("ms %.4f value %.1f e2e_ms %.4f eager %.4f frac %.3f cpu %s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"],
              d.get("eager_ms_per_step", 0), d["roofline"]["frac"], d.get("cpu_baseline", {}).get("value")))
        for k in ("config2_gat", "config4_bulk_inference"):
            print(k, json.dumps(d.get(k))[:300])
PY
