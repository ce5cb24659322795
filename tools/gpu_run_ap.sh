#!/bin/bash
# GPU call AP: full parity suite on the final build, smoke, final bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# per-branch device times (eager, CUDA events on each branch's stream) with and without the SM partition
timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/ap_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ap_tests.log; tail -5 gpurun_out/ap_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/ap_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/ap_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/ap_bench_ref.json 2> gpurun_out/ap_bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ap_bench_256_bf16.json 2> gpurun_out/ap_bench_256_bf16.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ap_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, round(d["ms_per_step"], 3), round(d["value"]), d.get("clocks"), {k: (round(r[k], 3) if r[k] is not None else None) for k in ("frac", "step_frac", "lm_step_frac", "lm_step_ms", "traffic") if k in r}, (d.get("latency") or {}).get("p99_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
