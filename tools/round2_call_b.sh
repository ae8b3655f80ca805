#!/bin/bash
# gpurun -- bash tools/round2_call_b.sh : 1-GPU verification of the round-2 tree — full GPU test suite, the default
# bench line (riders + CPU baseline), the A/B of the deferred split-K reduction, the launch list of one eager step.
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02b_pytest.log
timeout 600 python bench.py > $out/r02b_bench_full.json 2> $out/r02b_bench_full.err; echo "bench rc=$?"
GTS_DEFER_REDUCE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $out/r02b_bench_defer0.json 2> $out/r02b_bench_defer0.err; echo "bench defer0 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $out/r02b_bench_defer1.json 2> $out/r02b_bench_defer1.err; echo "bench defer1 rc=$?"
timeout 200 python tools/one_step.py 2 > $out/r02b_one_step_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r02b_launches.csv python tools/one_step.py 1 > $out/r02b_ncu_launches.log 2>&1
python - <<'PY'
import json
for nm in ("full", "defer0", "defer1"):
    try:
        for ln in open(f"gpurun_out/r02b_bench_{nm}.json"):
            if ln.startswith("{"):
                d = json.loads(ln)
                print(nm, "ms %.4f value %.1f e2e_ms %.4f eager %.4f frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d.get("eager_ms_per_step", 0), d["roofline"]["frac"]))
    except Exception as e:
        print(nm, "unreadable", e)
PY
