#!/bin/bash
# GPU call K: full parity suite on the current build; large-batch LM after the split / rope fixes; bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/k_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/k_tests.log; tail -5 gpurun_out/k_tests.log
timeout 300 python tools/profile_steps.py --workload lm --steps 5 --streams 2000 --kv int8 > gpurun_out/k_lm2000.txt 2>&1; grep -h alone gpurun_out/k_lm2000.txt
timeout 300 python tools/profile_steps.py --workload lm --steps 5 --kv int8 > gpurun_out/k_lm256_int8.txt 2>&1; grep -h alone gpurun_out/k_lm256_int8.txt
timeout 900 python bench.py --steps 20 --warmup 5 --kv int8 --streams 2000 --latency-steps 200 --no-cpu-baseline > gpurun_out/k_bench_2000_int8.json 2> gpurun_out/k_bench_2000_int8.err
timeout 600 python bench.py --steps 20 --warmup 5 --kv int8 --no-cpu-baseline --latency-steps 100 > gpurun_out/k_bench_256_int8.json 2> gpurun_out/k_bench_256_int8.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/k_bench_256_bf16.json 2> gpurun_out/k_bench_256_bf16.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/k_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, round(d["ms_per_step"], 3), round(d["value"]), d["clocks"], {k: round(r[k], 3) for k in ("frac", "step_frac", "lm_step_frac", "lm_step_ms") if k in r}, d["latency"]["p99_ms"])
    except Exception as e:
        print(f, "ERR", e)
PY
