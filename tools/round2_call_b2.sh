#!/bin/bash
# gpurun -- bash tools/round2_call_b2.sh : full GPU test suite (no -x: every failure listed), quick bench, launch list.
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r02c_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $out/r02c_bench.json 2> $out/r02c_bench.err; echo "bench rc=$?"
timeout 200 python tools/one_step.py 2 > $out/r02c_one_step_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r02c_launches.csv python tools/one_step.py 1 > $out/r02c_ncu_launches.log 2>&1
python - <<'PY'
import json
for ln in open("gpurun_out/r02c_bench.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("ms %.4f value %.1f e2e_ms %.4f eager %.4f frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d.get("eager_ms_per_step", 0), d["roofline"]["frac"]))
PY
